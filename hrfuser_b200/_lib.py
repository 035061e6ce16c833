"""ctypes binding of libhrfuser_b200.so (the C-ABI of include/hrfuser_b200.h).

The library is built in-tree by `hrfuser_b200.build` (nvcc, sm_100a).  Loading
fails loudly: there is no Python/CPU fallback for any op.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# HRF_LIB: alternative build of the same ABI (debug / instrumented), tools only
LIB_PATH = os.environ.get('HRF_LIB') or os.path.join(HERE, 'libhrfuser_b200.so')
ABI_VERSION = 9

HRF_F32, HRF_BF16, HRF_U8 = 0, 1, 2
MAX_FUSE_TERMS = 4


class HrfError(RuntimeError):
    pass


class AttnDesc(C.Structure):
    _fields_ = [('B', C.c_int32), ('H', C.c_int32), ('W', C.c_int32), ('C', C.c_int32),
                ('heads', C.c_int32), ('win', C.c_int32), ('n_kv', C.c_int32),
                ('dtype', C.c_int32), ('with_pad_mask', C.c_int32), ('ln_eps', C.c_float)]


class FfnDesc(C.Structure):
    _fields_ = [('B', C.c_int32), ('H', C.c_int32), ('W', C.c_int32), ('C', C.c_int32),
                ('hidden', C.c_int32), ('dtype', C.c_int32), ('ln_eps', C.c_float)]


class PwDesc(C.Structure):
    _fields_ = [('B', C.c_int32), ('H', C.c_int32), ('W', C.c_int32), ('Cin', C.c_int32),
                ('Cout', C.c_int32), ('dtype', C.c_int32), ('relu', C.c_int32)]


class DwPwDesc(PwDesc):
    pass


class ConvDesc(C.Structure):
    _fields_ = [('B', C.c_int32), ('H', C.c_int32), ('W', C.c_int32), ('Cin', C.c_int32),
                ('Cout', C.c_int32), ('stride', C.c_int32), ('dtype', C.c_int32), ('relu', C.c_int32)]


class ConvGemmDesc(C.Structure):
    _fields_ = [('B', C.c_int32), ('H', C.c_int32), ('W', C.c_int32), ('Cin', C.c_int32),
                ('Cout', C.c_int32), ('ksize', C.c_int32), ('stride', C.c_int32), ('relu', C.c_int32)]


class StemDesc(C.Structure):
    _fields_ = [('B', C.c_int32), ('Cin', C.c_int32), ('H', C.c_int32), ('W', C.c_int32),
                ('Cout', C.c_int32), ('relu', C.c_int32)]


class BnDesc(C.Structure):
    _fields_ = [('B', C.c_int32), ('C', C.c_int32), ('HW', C.c_int32), ('dtype', C.c_int32)]


class InputDesc(C.Structure):
    _fields_ = [('B', C.c_int32), ('H', C.c_int32), ('W', C.c_int32), ('C', C.c_int32),
                ('Hp', C.c_int32), ('Wp', C.c_int32), ('src_dtype', C.c_int32),
                ('to_rgb', C.c_int32), ('pad_val', C.c_float)]


class FuseDesc(C.Structure):
    _fields_ = [('B', C.c_int32), ('H', C.c_int32), ('W', C.c_int32), ('C', C.c_int32),
                ('dtype', C.c_int32), ('n_up', C.c_int32),
                ('up_H', C.c_int32 * MAX_FUSE_TERMS), ('up_W', C.c_int32 * MAX_FUSE_TERMS),
                ('n_same', C.c_int32), ('relu', C.c_int32)]


_F = C.POINTER(C.c_float)
_FP4 = C.POINTER(_F)          # const float* const bn[4]
_VPP = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); every symbol include/hrfuser_b200.h declares
SIGNATURES = {
    'hrf_abi_version': (C.c_int, []),
    'hrf_last_error': (C.c_char_p, []),
    'hrf_device_check': (C.c_int, []),
    'hrf_launch_count': (C.c_ulonglong, []),
    'hrf_set_pdl': (C.c_int, [C.c_int32]),
    'hrf_attn_blob_floats': (C.c_size_t, [C.POINTER(AttnDesc)]),
    'hrf_attn_pack': (C.c_int, [C.POINTER(AttnDesc)] + [_F] * 13 + [_F]),
    'hrf_attn_workspace_bytes': (C.c_size_t, [C.POINTER(AttnDesc)]),
    'hrf_window_attn_fwd': (C.c_int, [C.POINTER(AttnDesc), C.c_void_p, _VPP, _VPP, C.c_void_p,
                                      C.c_void_p, C.c_size_t, C.c_void_p]),
    'hrf_ffn_blob_floats': (C.c_size_t, [C.POINTER(FfnDesc)]),
    'hrf_ffn_pack': (C.c_int, [C.POINTER(FfnDesc), _F, _F, _F, _F, _FP4, _F, _F, _FP4, _F, _F,
                               _FP4, C.c_float, _F]),
    'hrf_ffn_workspace_bytes': (C.c_size_t, [C.POINTER(FfnDesc)]),
    'hrf_mixffn_fwd': (C.c_int, [C.POINTER(FfnDesc), C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_size_t, C.c_void_p]),
    'hrf_pw_blob_floats': (C.c_size_t, [C.POINTER(PwDesc)]),
    'hrf_pw_pack': (C.c_int, [C.POINTER(PwDesc), _F, _F, _FP4, C.c_float, _F]),
    'hrf_pw_fwd': (C.c_int, [C.POINTER(PwDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'hrf_conv3x3_blob_floats': (C.c_size_t, [C.POINTER(ConvDesc)]),
    'hrf_conv3x3_pack': (C.c_int, [C.POINTER(ConvDesc), _F, _F, _FP4, C.c_float, _F]),
    'hrf_conv3x3_fwd': (C.c_int, [C.POINTER(ConvDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'hrf_convgemm_supported': (C.c_int, [C.POINTER(ConvGemmDesc)]),
    'hrf_convgemm_blob_floats': (C.c_size_t, [C.POINTER(ConvGemmDesc)]),
    'hrf_convgemm_pack': (C.c_int, [C.POINTER(ConvGemmDesc), _F, _F, _FP4, C.c_float, _F, _F]),
    'hrf_convgemm_fwd': (C.c_int, [C.POINTER(ConvGemmDesc), C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
    'hrf_convgemm_grouped_fwd': (C.c_int, [C.POINTER(ConvGemmDesc), C.c_int32, _VPP, _VPP, _VPP, _VPP, C.c_void_p]),
    'hrf_convgemm_grouped_cat_fwd': (C.c_int, [C.POINTER(ConvGemmDesc), C.c_int32, _VPP, _VPP, C.c_int32, _VPP, _VPP,
                                              C.c_void_p]),
    'hrf_dwpw_blob_floats': (C.c_size_t, [C.POINTER(DwPwDesc)]),
    'hrf_dwpw_pack': (C.c_int, [C.POINTER(DwPwDesc), _F, _FP4, _F, _FP4, C.c_float, _F]),
    'hrf_dwpw_fwd': (C.c_int, [C.POINTER(DwPwDesc), C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_void_p]),
    'hrf_fuse_sum_fwd': (C.c_int, [C.POINTER(FuseDesc), C.c_void_p, _VPP, _VPP, C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
    'hrf_stem_blob_floats': (C.c_size_t, [C.POINTER(StemDesc)]),
    'hrf_stem_pack': (C.c_int, [C.POINTER(StemDesc), _F, _FP4, C.c_float, _F]),
    'hrf_stem_conv_fwd': (C.c_int, [C.POINTER(StemDesc), C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p]),
    'hrf_bias_act_fwd': (C.c_int, [C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    'hrf_bn_workspace_bytes': (C.c_size_t, [C.POINTER(BnDesc)]),
    'hrf_bn_stats': (C.c_int, [C.POINTER(BnDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                               C.c_void_p]),
    'hrf_bn_normalize': (C.c_int, [C.POINTER(BnDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    'hrf_bn_bwd_stats': (C.c_int, [C.POINTER(BnDesc), C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    'hrf_bn_bwd_dx': (C.c_int, [C.POINTER(BnDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p]),
    'hrf_bn_affine': (C.c_int, [C.POINTER(BnDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    'hrf_selftest_umma': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                    C.c_int32, C.c_void_p]),
    'hrf_input_prologue_fwd': (C.c_int, [C.POINTER(InputDesc), C.c_void_p, _F, _F, C.c_void_p,
                                         C.c_void_p]),
    'hrf_attn_core_train_fwd': (C.c_int, [C.c_int32] * 4 + [C.c_float] + [C.c_void_p] * 8),
    'hrf_attn_core_train_ws_floats': (C.c_size_t, [C.c_int32] * 3),
    'hrf_attn_core_train_bwd': (C.c_int, [C.c_int32] * 4 + [C.c_float] + [C.c_void_p] * 9 + [C.c_int32] +
                                [C.c_void_p] * 2 + [C.c_size_t, C.c_void_p]),
    'hrf_dwconv_train_ws_floats': (C.c_size_t, [C.c_int32] * 5),
    'hrf_dwconv_train_fwd': (C.c_int, [C.c_int32] * 5 + [C.c_void_p] * 5),
    'hrf_dwconv_train_dgrad': (C.c_int, [C.c_int32] * 5 + [C.c_void_p] * 4),
    'hrf_dwconv_train_wgrad': (C.c_int, [C.c_int32] * 5 + [C.c_void_p] * 5 + [C.c_size_t, C.c_void_p]),
    'hrf_mixffn_grouped_fwd': (C.c_int, [C.POINTER(FfnDesc), C.c_int32, _VPP, _VPP, _VPP, C.c_void_p, C.c_size_t,
                                         C.c_void_p]),
    'hrf_window_attn_grouped_fwd': (C.c_int, [C.POINTER(AttnDesc), C.c_int32, _VPP, _VPP, _VPP, C.c_void_p,
                                              C.c_size_t, C.c_void_p]),
    'hrf_ln_bwd_workspace_floats': (C.c_size_t, [C.c_int32, C.c_int32]),
    'hrf_ln_fwd': (C.c_int, [C.c_int32, C.c_int32, C.c_float] + [C.c_void_p] * 7),
    'hrf_ln_bwd': (C.c_int, [C.c_int32, C.c_int32] + [C.c_void_p] * 9 + [C.c_size_t, C.c_void_p]),
    'hrf_pool_fwd': (C.c_int, [C.c_int32] * 6 + [C.c_void_p, C.c_void_p, C.c_void_p]),
    'hrf_nchw_to_nhwc': (C.c_int, [C.c_int32] * 5 + [C.c_void_p, C.c_int32, C.c_void_p,
                                                     C.c_void_p]),
    'hrf_nhwc_to_nchw': (C.c_int, [C.c_int32] * 5 + [C.c_void_p, C.c_int32, C.c_void_p,
                                                     C.c_void_p]),
}

_lib = None


def load():
    """Load the shared library (once) and bind every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise HrfError(
            f'{LIB_PATH} is not built. Run `python -m hrfuser_b200.build` (needs nvcc); '
            'hrfuser_b200 has no CPU or eager-PyTorch fallback for inference.')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype, fn.argtypes = res, args
    if lib.hrf_abi_version() != ABI_VERSION:
        raise HrfError(f'ABI mismatch: library {lib.hrf_abi_version()} != binding {ABI_VERSION}; '
                       'rebuild with `python -m hrfuser_b200.build`')
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise HrfError(f'hrfuser_b200 error {rc}: {load().hrf_last_error().decode()}')
