/*
 * hrfuser_b200 -- C-ABI of the B200 (sm_100a) HRFuser fusion-backbone hot path.
 *
 * The reference (timbroed/HRFuser) is pure Python/PyTorch and has no native
 * layer; the entry points below are what a native binding for its hot path
 * binds instead of the PyTorch op sequences cited at each function
 * (file:line relative to the reference checkout).  INTEGRATION.md shows the
 * ctypes stub a reference maintainer would add.
 *
 * Conventions
 *   - plain C, no torch types; all activation pointers are DEVICE pointers to
 *     channels-last token tensors  [B][H][W][C]  (== the reference's (B, H*W, C)
 *     "NLC" layout, hrformer.py:368 / models/utils/transformer.py:49-59) of
 *     `dtype` HRF_F32 or HRF_BF16; arithmetic is fp32 unless stated.
 *   - weights are fp32 blobs packed by the hrf_*_pack() helpers below (host
 *     pointers in, host pointer out; copy the blob to the device yourself).
 *     The DEVICE copy of a blob must be 16-byte aligned: the tensor-core
 *     kernels stage its bf16 operand tiles with bulk async copies
 *     (HRF_EINVAL otherwise).
 *   - every *_fwd call is stream-ordered, non-allocating and non-blocking;
 *     the caller owns all memory.  `stream` is a cudaStream_t passed as void*.
 *   - return 0 on success, a negative HRF_E* code otherwise;
 *     hrf_last_error() returns a thread-local message.  Nothing throws or exits.
 */
#ifndef HRFUSER_B200_H_
#define HRFUSER_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HRF_ABI_VERSION 9

enum { HRF_F32 = 0, HRF_BF16 = 1, HRF_U8 = 2 /* hrf_input_prologue_fwd source only */ };
enum {
  HRF_OK = 0,
  HRF_EINVAL = -1,      /* bad descriptor / null pointer            */
  HRF_EUNSUPPORTED = -2,/* shape outside what the kernels cover     */
  HRF_ECUDA = -3,       /* CUDA runtime error (see hrf_last_error)  */
  HRF_EDEVICE = -4      /* not an sm_100 device                     */
};

int hrf_abi_version(void);
const char* hrf_last_error(void);
/* 0 when the current device can run the sm_100a kernels. */
int hrf_device_check(void);
/* Number of kernels this library has launched since load (all threads). */
unsigned long long hrf_launch_count(void);
/* Programmatic dependent launch (a kernel's prologue overlaps its predecessor's tail).  Off by
 * default (HRF_PDL=1 in the environment enables it): it helps a single-stream chain of kernels
 * and costs ~2 % in the engine's multi-stream graph.  Returns the previous setting.  Affects
 * launches (and graph captures) made after the call. */
int hrf_set_pdl(int32_t enable);

/* ------------------------------------------------------------------------
 * Window attention: LocalWindowSelfAttention + WindowMSA
 * (hrformer.py:96-131,184-236) and MultiWindowCrossAttention + WindowMCA
 * (hrfuser_hrformer_based.py:106-151,189-248), fused with the LayerNorms in
 * front of them and the residual adds behind them
 * (hrformer.py:369; hrfuser_hrformer_based.py:305-313).
 *
 *   n_kv == 0 :  out = x + Attn(LN(x))                                 (LSA)
 *   n_kv >= 1 :  out = x + sum_k [ kv_k + Attn_k(LN1_k(x), LN2_k(kv_k)) ] (MWCA)
 *
 * Centre padding to window multiples happens after the LayerNorm with zeros;
 * padded slots are live keys unless with_pad_mask (and both pads > 0).
 * No window tensor is materialised: pad / partition / merge / crop are
 * address arithmetic inside the kernel.
 * ---------------------------------------------------------------------- */
typedef struct HrfAttnDesc {
  int32_t B, H, W, C;
  int32_t heads;
  int32_t win;            /* window edge; 7 in every shipped config */
  int32_t n_kv;           /* 0 = self-attention, else number of key/value modalities */
  int32_t dtype;          /* HRF_F32 | HRF_BF16 (activation storage) */
  int32_t with_pad_mask;
  float   ln_eps;         /* 1e-6 */
} HrfAttnDesc;

/* floats in one packed blob (one per modality for MWCA, one for LSA) */
size_t hrf_attn_blob_floats(const HrfAttnDesc* d);
/* Host-side packing.  wq/wk/wv/wo are row-major (C_out, C_in) nn.Linear
 * weights; for LSA pass the three (C, C) slices of qkv.weight.  ln_kv_* may
 * equal ln_q_* (LSA).  rpb_table is ((2*win-1)^2, heads) or NULL (with_rpe=False). */
int hrf_attn_pack(const HrfAttnDesc* d,
                  const float* ln_q_w, const float* ln_q_b,
                  const float* ln_kv_w, const float* ln_kv_b,
                  const float* wq, const float* bq, const float* wk, const float* bk,
                  const float* wv, const float* bv, const float* wo, const float* bo,
                  const float* rpb_table, float* blob_out);
/* Scratch the call needs (device bytes; 0 for most shapes).  Wide low-resolution
 * branches (C = 72, 144 in bf16) split the heads of a window pair over several
 * CTAs and sum their fp32 partial out-projections from this workspace. */
size_t hrf_attn_workspace_bytes(const HrfAttnDesc* d);
/* kv: array of n_kv device pointers (ignored when n_kv == 0);
 * blobs: array of max(1, n_kv) device pointers to packed blobs;
 * workspace: >= hrf_attn_workspace_bytes(d) bytes (may be NULL when that is 0).
 * out must not alias x or any kv. */
int hrf_window_attn_fwd(const HrfAttnDesc* d, const void* x, const void* const* kv,
                        const float* const* blobs, void* out, void* workspace,
                        size_t workspace_bytes, void* stream);
/* Self-attention (n_kv == 0) of n tensors of ONE shape, each with its own packed blob (the
 * camera's branch 0 and the modality streams walk the same HRFormer blocks in the same stages):
 * ONE launch where the kernel covers it (C = 18), one per tensor otherwise.  outs[q] must not
 * alias xs[q].  The workspace rule of hrf_window_attn_fwd applies (shared by the tensors). */
int hrf_window_attn_grouped_fwd(const HrfAttnDesc* d, int32_t n, const void* const* xs,
                                const float* const* blobs, void* const* outs, void* workspace,
                                size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------
 * MixFFN (CrossFFN, hrformer.py:267-295) fused with the LayerNorm in front and
 * the residual behind (hrformer.py:371; hrfuser_hrformer_based.py:315):
 *   out = x + GELU(BN3(W2 * GELU(BN2(dw3x3(GELU(BN1(W1 * LN(x) + b1))) + bd)) + b2))
 * BatchNorms are eval-mode affines folded into the packed weights.
 * ---------------------------------------------------------------------- */
typedef struct HrfFfnDesc {
  int32_t B, H, W, C;
  int32_t hidden;         /* mlp_ratio * C */
  int32_t dtype;
  float   ln_eps;
} HrfFfnDesc;

size_t hrf_ffn_blob_floats(const HrfFfnDesc* d);
/* bnX = {weight, bias, running_mean, running_var} each of the layer's width. */
int hrf_ffn_pack(const HrfFfnDesc* d, const float* ln_w, const float* ln_b,
                 const float* w1, const float* b1, const float* const bn1[4],
                 const float* wd, const float* bd, const float* const bn2[4],
                 const float* w2, const float* b2, const float* const bn3[4],
                 float bn_eps, float* blob_out);
/* scratch bytes (0 except C = 144 in bf16, whose 8 hidden chunks are split over CTAs) */
size_t hrf_ffn_workspace_bytes(const HrfFfnDesc* d);
int hrf_mixffn_fwd(const HrfFfnDesc* d, const void* x, const float* blob, void* out,
                   void* workspace, size_t workspace_bytes, void* stream);
/* The same block on n tensors of ONE shape, each with its own packed blob (the camera's branch 0 and
 * the modality streams): ONE launch where the kernel covers it (C = 18), one per tensor otherwise. */
int hrf_mixffn_grouped_fwd(const HrfFfnDesc* d, int32_t n, const void* const* xs,
                           const float* const* blobs, void* const* outs, void* workspace,
                           size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------
 * Multi-resolution exchange (HRModule.forward hrnet.py:184-207 with the fuse
 * layers of hrformer.py:498-561).
 * ---------------------------------------------------------------------- */
typedef struct HrfPwDesc {      /* pointwise (1x1) conv + folded BN (+ReLU)  */
  int32_t B, H, W, Cin, Cout;
  int32_t dtype;
  int32_t relu;
} HrfPwDesc;
size_t hrf_pw_blob_floats(const HrfPwDesc* d);
/* w: (Cout, Cin); bn = {weight,bias,mean,var} or NULL; bias may be NULL */
int hrf_pw_pack(const HrfPwDesc* d, const float* w, const float* bias,
                const float* const bn[4], float bn_eps, float* blob_out);
int hrf_pw_fwd(const HrfPwDesc* d, const void* x, const float* blob, void* out, void* stream);

typedef struct HrfConvDesc {    /* dense 3x3 conv, pad 1, stride 1|2, + folded BN (+ReLU): the
                                   transition layers (hrformer.py:562-607 and the modality
                                   transitions of hrfuser_hrformer_based.py) */
  int32_t B, H, W, Cin, Cout;   /* H, W: INPUT size; output is ceil(H/stride) x ceil(W/stride) */
  int32_t stride;
  int32_t dtype;
  int32_t relu;
} HrfConvDesc;
size_t hrf_conv3x3_blob_floats(const HrfConvDesc* d);
/* w: (Cout, Cin, 3, 3); bn = {weight,bias,mean,var} or NULL; bias may be NULL */
int hrf_conv3x3_pack(const HrfConvDesc* d, const float* w, const float* bias,
                     const float* const bn[4], float bn_eps, float* blob_out);
/* x: [B][H][W][Cin] tokens -> out: [B][Ho][Wo][Cout] tokens */
int hrf_conv3x3_fwd(const HrfConvDesc* d, const void* x, const float* blob, void* out, void* stream);

typedef struct HrfDwPwDesc {    /* dw3x3 stride 2 pad 1 + BN + 1x1 + BN (+ReLU): one
                                   down-sampling step, hrformer.py:531-557 */
  int32_t B, H, W, Cin, Cout;   /* H, W: INPUT size; output is ceil(H/2) x ceil(W/2) */
  int32_t dtype;
  int32_t relu;
} HrfDwPwDesc;
size_t hrf_dwpw_blob_floats(const HrfDwPwDesc* d);
int hrf_dwpw_pack(const HrfDwPwDesc* d, const float* wdw /*(Cin,1,3,3)*/,
                  const float* const bn_dw[4], const float* wpw /*(Cout,Cin)*/,
                  const float* const bn_pw[4], float bn_eps, float* blob_out);
int hrf_dwpw_fwd(const HrfDwPwDesc* d, const void* x, const float* blob, void* out,
                 void* stream);

/* Dense 1x1 / 3x3 convolution + folded BN (+ residual) (+ ReLU) on channels-last bf16 tokens:
 * the stem conv2, the Bottlenecks (resnet.py:263-302: conv1 / conv2 / conv3 + identity or
 * downsample, ReLU after the add), the 256-channel transitions (hrnet.py:419-463,
 * hrfuser_hrformer_based.py:375-412) and the HRFPN 3x3 convs (necks/hrfpn.py:87-100) --
 * SURVEY.md 8(f) rank 1 / 2.  One warp-specialised, TMA-fed tcgen05 implicit-GEMM kernel
 * (csrc/conv_gemm_tc.cuh).  ksize 1 (pad 0, stride 1) or 3 (pad 1, stride 1 | 2);
 * Cin a multiple of 64, Cout even and <= 256, dtype HRF_BF16. */
typedef struct HrfConvGemmDesc {
  int32_t B, H, W, Cin, Cout;   /* H, W: INPUT size; output is ceil(H/stride) x ceil(W/stride) */
  int32_t ksize;
  int32_t stride;
  int32_t relu;                 /* applied after the residual add */
} HrfConvGemmDesc;
/* 0 when the kernel covers the problem (else HRF_EUNSUPPORTED: keep the layer on cuDNN) */
int hrf_convgemm_supported(const HrfConvGemmDesc* d);
size_t hrf_convgemm_blob_floats(const HrfConvGemmDesc* d);
/* w: (Cout, Cin, k, k); bias may be NULL; bn = {weight,bias,mean,var} or NULL; extra_bias
 * (Cout) may be NULL -- the folded shift of a bias-free downsample branch that is added to
 * conv3's bias so that the add + ReLU fuse into conv3 */
int hrf_convgemm_pack(const HrfConvGemmDesc* d, const float* w, const float* bias,
                      const float* const bn[4], float bn_eps, const float* extra_bias,
                      float* blob_out);
/* x: [B][H][W][Cin] bf16 tokens; resid: [B][Ho][Wo][Cout] bf16 or NULL; out likewise
 * (must not alias x).  All 16-byte aligned. */
int hrf_convgemm_fwd(const HrfConvGemmDesc* d, const void* x, const void* resid,
                     const float* blob, void* out, void* stream);
/* The same layer shape on n (1..4) independent tensors, each with its own weights, in ONE
 * launch: the camera stream and the modality streams of the backbone run identical stem /
 * Bottleneck / transition layers (hrfuser_hrformer_based.py:533-543).  xs / outs / blobs:
 * arrays of n device pointers; resids: array of n pointers or NULL.  d->relu is a bit mask
 * here: bit q = ReLU on problem q. */
int hrf_convgemm_grouped_fwd(const HrfConvGemmDesc* d, int32_t n, const void* const* xs,
                             const void* const* resids, const float* const* blobs,
                             void* const* outs, void* stream);
/* The 1x1 convolution of the channel concatenation [x | x2] (x: cin1 channels, x2: d->Cin - cin1;
 * both multiples of 64) without materialising it: the K loop takes its first cin1 / 64 chunks
 * from x, the rest from x2.  A Bottleneck's conv3 and its downsample branch
 * (resnet.py:263-302: out = relu(bn3(conv3(y)) + downsample(x))) are ONE GEMM this way, weights
 * [W3 * s3 | Wd * sd] and bias b3 + bd folded by the caller: the downsample map is never written
 * or re-read and the sum stays in the fp32 accumulator.  ksize 1, stride 1, no residual. */
int hrf_convgemm_grouped_cat_fwd(const HrfConvGemmDesc* d, int32_t n, const void* const* xs,
                                 const void* const* xs2, int32_t cin1, const float* const* blobs,
                                 void* const* outs, void* stream);

#define HRF_MAX_FUSE_TERMS 4
typedef struct HrfFuseDesc {    /* out = ReLU(x + sum_j bilinear_up(up_j) + sum_j same_j) */
  int32_t B, H, W, C;           /* output branch size */
  int32_t dtype;
  int32_t n_up;                 /* coarser terms, bilinearly up-sampled (align_corners=False,
                                   scale from sizes: hrnet.py:199-203) */
  int32_t up_H[HRF_MAX_FUSE_TERMS], up_W[HRF_MAX_FUSE_TERMS];
  int32_t n_same;               /* terms already at (H, W) */
  int32_t relu;
} HrfFuseDesc;
/* out_nchw_f32 may be NULL; when given, the result is additionally written as a
 * contiguous fp32 (B, C, H, W) tensor -- the backbone's output contract
 * (hrfuser_hrformer_based.py:627). */
int hrf_fuse_sum_fwd(const HrfFuseDesc* d, const void* x, const void* const* up,
                     const void* const* same, void* out, float* out_nchw_f32, void* stream);

/* Stem convolution (bf16 mode): conv 3x3 stride 2 pad 1 with Cin <= 3 + folded BN + ReLU
 * (hrnet.py:341-348 conv1/bn1, hrfuser_hrformer_based.py:380-388 conv_a/norm_a), reading
 * the caller's fp32 NCHW image and writing bf16 channels-last (B, ceil(H/2), ceil(W/2), Cout).
 * Implicit GEMM on the tensor cores (K = 9*Cin padded to 32).  Cout = 64. */
typedef struct HrfStemDesc {
  int32_t B, Cin, H, W, Cout;
  int32_t relu;
} HrfStemDesc;
size_t hrf_stem_blob_floats(const HrfStemDesc* d);
int hrf_stem_pack(const HrfStemDesc* d, const float* w /*(Cout,Cin,3,3)*/, const float* const bn[4],
                  float bn_eps, float* blob_out);
int hrf_stem_conv_fwd(const HrfStemDesc* d, const float* x_nchw, const float* blob,
                      void* out_nhwc_bf16, void* stream);

/* Epilogue of the cuDNN-side convolutions (stems, Bottlenecks, transitions:
 * hrnet.py:341-371,419-463, resnet.py:263-302 with the BatchNorm folded into the
 * conv): y = act(y + bias[c] (+ residual)) in place, one pass, channels-last.
 * bias is fp32 [C] (device); residual may be NULL. */
int hrf_bias_act_fwd(int64_t n_tokens, int32_t C, int32_t dtype, int32_t relu, void* y,
                     const float* bias, const void* residual, void* stream);

/* Train-mode BatchNorm / SyncBatchNorm statistics and per-channel affine passes on
 * contiguous NCHW planes (x is [B][C][HW]), replacing what nn.BatchNorm2d /
 * nn.SyncBatchNorm compute in the reference's training forward and backward
 * (mmcv build_norm_layer call sites hrnet.py:338-358, hrformer.py:267-282,
 * resnet.py:161-206; cfg norm_cfg=dict(type='SyncBN'),
 * configs/_base_/models/cascade_rcnn_hrfuser_fpn_nus_clr_fusion.py:2).
 *   hrf_bn_stats:      stats[0..C) = sum_x, stats[C..2C) = sum_x^2, stats[2C] = B*HW
 *                      (2C + 1 doubles, device: the SyncBN message)
 *   hrf_bn_bwd_stats:  sums[0..C) = sum_dy, sums[C..2C) = sum_dy*(x-mean)*invstd (2C doubles);
 *                      dbias / dweight (fp32 [C], may be NULL) receive the same numbers: the
 *                      rank-local parameter gradients
 * Both are additive across ranks: for SyncBN the caller all-reduces them over NCCL between
 * the statistics call and the apply call.  Deterministic (fixed-order two-stage reduction,
 * the forward one centred per chunk so that fp32 never sees E[x^2] - mean^2).
 * workspace >= hrf_bn_workspace_bytes(d).
 *   hrf_bn_normalize:  y = (x - mean) * invstd * weight + bias from the (reduced) stats;
 *                      stores mean / invstd (fp32 [C]) for the backward and, when
 *                      running_mean / running_var are given, updates them in place with
 *                      `momentum` and the unbiased variance, as nn.BatchNorm2d does.
 *                      weight / bias may be NULL (1 / 0).
 *   hrf_bn_bwd_dx:     dx = weight*invstd*(dy - sum_dy/n - xhat * sum_dy_xhat/n) from the
 *                      (reduced) sums; count points at the forward's stats[2C].
 * Fused activation: `act` = HRF_BN_ACT_NONE | _RELU | _GELU (exact, erf) is the nn.ReLU /
 * nn.GELU module that follows the norm layer (hrformer.py:267-282 MlpDWBN, hrnet.py:338-358,
 * resnet.py:161-206).  hrf_bn_normalize returns act(BN(x)); the two backward calls take dy
 * with respect to THAT output and multiply it by act'(z), z = weight * xhat + bias recomputed
 * from x and the saved statistics (pass the same weight / bias / act to both); with act = 0
 * weight / bias are not read by hrf_bn_bwd_stats.
 *   hrf_bn_affine:     out = a[c]*x + c0[c]  (dy == NULL)  or  a[c]*dy + b[c]*x + c0[c],
 *                      caller-supplied fp32 [C] coefficients.
 * out / y / dx may alias their inputs. */
typedef struct {
  int32_t B, C, HW;
  int32_t dtype;
} HrfBnDesc;
enum { HRF_BN_ACT_NONE = 0, HRF_BN_ACT_RELU = 1, HRF_BN_ACT_GELU = 2 };
size_t hrf_bn_workspace_bytes(const HrfBnDesc* d);
int hrf_bn_stats(const HrfBnDesc* d, const void* x, double* stats, void* workspace,
                 size_t workspace_bytes, void* stream);
int hrf_bn_normalize(const HrfBnDesc* d, const void* x, const double* stats, const float* weight,
                     const float* bias, float eps, float momentum, float* running_mean,
                     float* running_var, float* save_mean, float* save_invstd, int32_t act,
                     void* y, void* stream);
int hrf_bn_bwd_stats(const HrfBnDesc* d, const void* x, const void* dy, const float* mean,
                     const float* invstd, const float* weight, const float* bias, int32_t act,
                     double* sums, float* dweight, float* dbias, void* workspace,
                     size_t workspace_bytes, void* stream);
int hrf_bn_bwd_dx(const HrfBnDesc* d, const void* x, const void* dy, const double* sums,
                  const double* count, const float* weight, const float* bias, int32_t act,
                  const float* mean, const float* invstd, void* dx, void* stream);
int hrf_bn_affine(const HrfBnDesc* d, const void* x, const void* dy, const float* a,
                  const float* b, const float* c0, int32_t relu, void* out, void* stream);

/* ------------------------------------------------------------------------
 * Train-mode LayerNorm over the channel axis of fp32 token tensors [rows][C]
 * (nn.LayerNorm(C, eps) of hrformer.py:345-371 / hrfuser_hrformer_based.py:262-313 in the
 * training configs).  C <= 1024.
 *   hrf_ln_fwd: y = (x - mean) * rstd * gamma + beta; saves mean / rstd (fp32 [rows]).
 *   hrf_ln_bwd: dx (may be NULL), dgamma, dbeta (fp32 [C]) from x, dy and the saved statistics;
 *               deterministic two-stage reduction through `workspace`
 *               (>= hrf_ln_bwd_workspace_floats(rows, C) floats).
 * ---------------------------------------------------------------------- */
size_t hrf_ln_bwd_workspace_floats(int32_t rows, int32_t C);
int hrf_ln_fwd(int32_t rows, int32_t C, float eps, const float* x, const float* gamma,
               const float* beta, float* y, float* mean, float* rstd, void* stream);
int hrf_ln_bwd(int32_t rows, int32_t C, const float* x, const float* dy, const float* mean,
               const float* rstd, const float* gamma, float* dx, float* dgamma, float* dbeta,
               float* workspace, size_t workspace_floats, void* stream);

/* ------------------------------------------------------------------------
 * Training-mode window-attention core, fp32, forward and backward: the part of WindowMSA /
 * WindowMCA between the q / k / v projections and the output projection
 * (hrformer.py:110-131, hrfuser_hrformer_based.py:127-151):
 *     P = softmax(scale * Q K^T + table[rpi]),   O = P V       per (window, head)
 * q / k / v / o / their gradients: [nWin][N][C] tokens, head h in columns h*hd..(h+1)*hd
 * (C = heads * hd; N = Wh*Ww <= 64, hd <= 64).  table: [T][heads] relative-position bias
 * table with its [N*N] int32 index `rpi`, or both NULL.  P ([nWin][heads][N][N]) is saved by
 * the forward for the backward.  dtable ([T][heads]) is summed in fixed order through
 * `workspace` (>= hrf_attn_core_train_ws_floats floats); pass dtable = NULL to skip it.
 * ---------------------------------------------------------------------- */
int hrf_attn_core_train_fwd(int32_t nWin, int32_t N, int32_t C, int32_t heads, float scale,
                            const float* q, const float* k, const float* v, const float* table,
                            const int32_t* rpi, float* o, float* P, void* stream);
size_t hrf_attn_core_train_ws_floats(int32_t nWin, int32_t N, int32_t heads);
int hrf_attn_core_train_bwd(int32_t nWin, int32_t N, int32_t C, int32_t heads, float scale,
                            const float* q, const float* k, const float* v, const float* P,
                            const float* dout, float* dq, float* dk, float* dv,
                            const int32_t* rpi, int32_t T, float* dtable, float* workspace,
                            size_t workspace_floats, void* stream);

/* ------------------------------------------------------------------------
 * Training-mode depthwise 3x3 convolution (padding 1, stride 1 | 2, groups = C), fp32 NCHW:
 * the depthwise conv of CrossFFN (hrformer.py:273-279) and the stride-2 depthwise convs of the
 * HRModule exchange (hrformer.py:528-548) in the training configs.  x: [B][C][H][W],
 * y / g: [B][C][Ho][Wo] with Ho = (H - 1) / stride + 1; w: [C][3][3]; bias may be NULL.
 * hrf_dwconv_train_wgrad sums per-CTA partials in fixed order through `workspace`
 * (>= hrf_dwconv_train_ws_floats floats); dbias may be NULL.
 * ---------------------------------------------------------------------- */
size_t hrf_dwconv_train_ws_floats(int32_t B, int32_t C, int32_t H, int32_t W, int32_t stride);
int hrf_dwconv_train_fwd(int32_t B, int32_t C, int32_t H, int32_t W, int32_t stride, const float* x,
                         const float* w, const float* bias, float* y, void* stream);
int hrf_dwconv_train_dgrad(int32_t B, int32_t C, int32_t H, int32_t W, int32_t stride, const float* g,
                           const float* w, float* dx, void* stream);
int hrf_dwconv_train_wgrad(int32_t B, int32_t C, int32_t H, int32_t W, int32_t stride, const float* x,
                           const float* g, float* dw, float* dbias, float* workspace,
                           size_t workspace_floats, void* stream);

/* Diagnostic: D[128][N] (fp32) = A[128][K] (bf16) x B on the tcgen05 tensor cores,
 * through the same descriptor helpers the fused kernels use.  B is [N][K]
 * (b_mn_major = 0, K-major operand) or [K][N] (b_mn_major = 1, MN-major operand).
 * N, K multiples of 16, <= 256.  Device pointers. */
int hrf_selftest_umma(const void* A, const void* B, float* D, int32_t N, int32_t K,
                      int32_t b_mn_major, void* stream);

/* ------------------------------------------------------------------------
 * Input prologue: Normalize -> Pad(size_divisor) ->
 * DefaultFormatBundle -> batch for one sensor stream, one pass on the device.
 * Replaces, per step and per sensor key ('img', 'lidar_img', 'radar_img', 'gated_img'):
 *   Normalize.__call__            mmdet/datasets/pipelines/transforms.py:719-744
 *     (mmcv.imnormalize, mmcv-full 1.3.17: astype(float32), optional BGR->RGB,
 *      cv2.subtract(img, mean) in fp32, cv2.multiply(img, 1/float64(std)) in fp64)
 *   Pad._pad_img                  transforms.py:652-667 (impad_to_multiple: pad_val bottom / right,
 *                                 applied after Normalize)
 *   DefaultFormatBundle.__call__  formating.py:211-227  (uint8 -> float32, HWC -> CHW)
 *   and the collate stack of equally sized frames.
 * src: (B, H, W, C) contiguous, HRF_U8 or HRF_F32, C in 1..4 (device pointer).
 * mean / std: C host floats (the config's img_norm_cfg values as float32);
 * stdinv = 1.0 / (double)std[c] is formed here exactly as the reference does.
 * to_rgb reverses the channel order (3-channel images only).
 * dst: (B, C, Hp, Wp) fp32, Hp >= H, Wp >= W, Wp a multiple of 4 (the reference pads to
 * multiples of 32); pixels outside H x W are pad_val.
 * y = fl32((double)fl32(x - mean[c]) * stdinv[c]), OpenCV's roundings: bit-exact to the reference. */
typedef struct {
  int32_t B, H, W, C;     /* source frames                                   */
  int32_t Hp, Wp;         /* padded output size                              */
  int32_t src_dtype;      /* HRF_U8 | HRF_F32                                */
  int32_t to_rgb;         /* 1: BGR -> RGB before normalising                */
  float pad_val;          /* Pad's pad_val['img'] (0 in every shipped config)*/
} HrfInputDesc;
int hrf_input_prologue_fwd(const HrfInputDesc* d, const void* src, const float* mean,
                           const float* std, float* dst, void* stream);

/* k x k pooling with stride k of bf16 tokens [B][H][W][C] -> [B][H/k][W/k][C]: the HRFPN
 * pyramid levels (necks/hrfpn.py:88-92, F.avg_pool2d / F.max_pool2d, kernel_size = stride =
 * 2^i).  C a multiple of 8, H and W multiples of k. */
int hrf_pool_fwd(int32_t B, int32_t H, int32_t W, int32_t C, int32_t k, int32_t is_max,
                 const void* x, void* out, void* stream);

/* Layout converters at the boundary of the path. */
int hrf_nchw_to_nhwc(int32_t B, int32_t C, int32_t H, int32_t W, int32_t src_dtype,
                     const void* src, int32_t dst_dtype, void* dst, void* stream);
int hrf_nhwc_to_nchw(int32_t B, int32_t C, int32_t H, int32_t W, int32_t src_dtype,
                     const void* src, int32_t dst_dtype, void* dst, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HRFUSER_B200_H_ */
