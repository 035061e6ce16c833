// Input prologue: Normalize -> Pad(size_divisor) -> DefaultFormatBundle (HWC -> CHW) -> batch,
// one pass per sensor stream.
//
// Reference (CPU, numpy / OpenCV, one pass per step and per sensor):
//   Normalize.__call__            mmdet/datasets/pipelines/transforms.py:719-744
//     -> mmcv.imnormalize (mmcv-full 1.3.17, not in the tree): astype(float32), optional
//        BGR->RGB, cv2.subtract(img, mean), cv2.multiply(img, 1/float64(std)).  OpenCV (4.13,
//        checked bit for bit, tests/golden/make_golden_input.py) subtracts in fp32 and multiplies
//        in fp64: y = fl32(f64(fl32(x - mean_f32)) * stdinv_f64), stdinv_f64 = 1 / f64(std_f32)
//   Pad._pad_img                  transforms.py:652-667  (impad_to_multiple: zeros bottom/right,
//        applied AFTER Normalize, so the border is pad_val, not a normalised value)
//   DefaultFormatBundle.__call__  formating.py:211-227   (uint8 -> float32, HWC -> CHW)
//   collate                       stack of equally sized frames along dim 0
//
// HBM-bound: reads B*H*W*C source elements (u8 or fp32) once, writes B*C*Hp*Wp fp32 once.
// One thread per 4 consecutive output pixels of a row: the C*4 source elements of those
// pixels are one contiguous run (a warp reads 128 consecutive pixels = one contiguous
// segment), and each channel plane gets one 16-byte store, 512 bytes per warp.  The grid is one
// thread per quad (no grid-stride loop): the kernel is 9-55 us long, so what matters is that all
// loads are in flight at once, not CTA reuse.
#pragma once
#include "common.cuh"

namespace hrf {

constexpr int INPUT_MAX_C = 4;

struct InputNorm {
  float mean[INPUT_MAX_C];
  double stdinv[INPUT_MAX_C];
};

template <typename S>
__device__ __forceinline__ float input_load(const S* p) {
  return (float)__ldg(p);
}

template <typename S, int C>
__global__ void __launch_bounds__(256) input_prologue_kernel(const S* __restrict__ src,
                                                             float* __restrict__ dst, InputNorm nm,
                                                             int H, int W, int Hp, int Wp,
                                                             int to_rgb, float pad_val,
                                                             int n_quads, FastDiv div_wq,
                                                             FastDiv div_hp) {
  // one quad (4 pixels of a row, all channels) per thread, no loop: every load of the launch is
  // issued up front; 32-bit index arithmetic with magic-number division
  {
    const int q = (int)(blockIdx.x * 256u + threadIdx.x);
    if (q >= n_quads) return;
    int r, xq, bi, y;
    div_wq.divmod(q, r, xq);
    div_hp.divmod(r, bi, y);
    const long long b = bi;
    const int x0 = xq << 2;
    float v[C][4];
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int i = 0; i < 4; ++i) v[c][i] = pad_val;
    if (y < H && x0 < W) {
      const S* p = src + ((b * H + y) * (long long)W + x0) * C;
      const int nx = min(4, W - x0);
      float e[4 * C];
      if (nx == 4) {
        if constexpr (sizeof(S) == 4) {
          if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
            for (int i = 0; i < C; ++i) {
              const float4 f = __ldg(reinterpret_cast<const float4*>(p) + i);
              e[4 * i] = f.x, e[4 * i + 1] = f.y, e[4 * i + 2] = f.z, e[4 * i + 3] = f.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 4 * C; ++i) e[i] = input_load(p + i);
          }
        } else {
          if ((reinterpret_cast<uintptr_t>(p) & 3) == 0) {
#pragma unroll
            for (int i = 0; i < C; ++i) {
              const uchar4 u = __ldg(reinterpret_cast<const uchar4*>(p) + i);
              e[4 * i] = (float)u.x, e[4 * i + 1] = (float)u.y, e[4 * i + 2] = (float)u.z,
                    e[4 * i + 3] = (float)u.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 4 * C; ++i) e[i] = input_load(p + i);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4 * C; ++i) e[i] = i < nx * C ? input_load(p + i) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (i < nx) {
#pragma unroll
          for (int c = 0; c < C; ++c) {
            // BGR -> RGB: output channel c is source channel C-1-c (3-channel images only)
            const float a = to_rgb ? e[i * C + (C - 1 - c)] : e[i * C + c];
            // OpenCV's roundings: fp32 subtract, fp64 multiply, one rounding back to fp32
            v[c][i] = __double2float_rn(__dmul_rn((double)__fsub_rn(a, nm.mean[c]), nm.stdinv[c]));
          }
        }
      }
    }
    float* o = dst + ((b * C) * Hp + y) * (long long)Wp + x0;
#pragma unroll
    for (int c = 0; c < C; ++c)
      *reinterpret_cast<float4*>(o + (long long)c * Hp * Wp) = make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
  }
}

template <typename S>
static int launch_input_prologue_t(int B, int H, int W, int C, int Hp, int Wp, const void* src,
                                   const InputNorm& nm, int to_rgb, float pad_val, float* dst,
                                   cudaStream_t stream) {
  const long long n_quads_ll = (long long)B * Hp * (Wp >> 2);
  HRF_REQUIRE(n_quads_ll < (1ll << 31) - 256, HRF_EUNSUPPORTED, "input_prologue: more than 2^33 output pixels");
  const int n_quads = (int)n_quads_ll;
  const int grid = (n_quads + 255) / 256;
  const FastDiv dq(Wp >> 2), dh(Hp);
  const S* s = (const S*)src;
  switch (C) {
    case 1: input_prologue_kernel<S, 1><<<grid, 256, 0, stream>>>(s, dst, nm, H, W, Hp, Wp, to_rgb, pad_val, n_quads, dq, dh); break;
    case 2: input_prologue_kernel<S, 2><<<grid, 256, 0, stream>>>(s, dst, nm, H, W, Hp, Wp, to_rgb, pad_val, n_quads, dq, dh); break;
    case 3: input_prologue_kernel<S, 3><<<grid, 256, 0, stream>>>(s, dst, nm, H, W, Hp, Wp, to_rgb, pad_val, n_quads, dq, dh); break;
    case 4: input_prologue_kernel<S, 4><<<grid, 256, 0, stream>>>(s, dst, nm, H, W, Hp, Wp, to_rgb, pad_val, n_quads, dq, dh); break;
    default: HRF_REQUIRE(false, HRF_EUNSUPPORTED, "input_prologue: 1..4 channels");
  }
  count_launch();
  HRF_CUDA(cudaGetLastError());
  return HRF_OK;
}

}  // namespace hrf
