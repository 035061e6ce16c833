// k x k / stride k pooling of a token tensor [B][H][W][C] (bf16): the HRFPN pyramid levels
// (necks/hrfpn.py:88-92: F.avg_pool2d / F.max_pool2d with kernel_size = stride = 2^i).
// Bandwidth-bound: a thread owns 8 channels (one 16-byte chunk) of an output token, walks the
// k x k window with 128-bit loads, accumulates in fp32.
#pragma once
#include "common.cuh"

namespace hrf {

template <bool MAX>
__global__ void __launch_bounds__(256) pool_tokens_kernel(const __nv_bfloat16* __restrict__ x,
                                                          __nv_bfloat16* __restrict__ out, int B, int H, int W,
                                                          int C, int k, FastDiv d_cv, FastDiv d_wo, FastDiv d_ho) {
  pdl_launch_dependents();
  pdl_wait();
  const int Ho = H / k, Wo = W / k, cv = C / 8;
  const long long n = (long long)B * Ho * Wo * cv;
  const float inv = 1.0f / (float)(k * k);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x) {
    int t, c8, r, wo, b, ho;
    d_cv.divmod((int)idx, t, c8);
    d_wo.divmod(t, r, wo);
    d_ho.divmod(r, b, ho);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = MAX ? -INFINITY : 0.f;
    const __nv_bfloat16* src = x + (((size_t)b * H + (size_t)ho * k) * W + (size_t)wo * k) * C + c8 * 8;
    for (int dy = 0; dy < k; ++dy)
      for (int dx = 0; dx < k; ++dx) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + ((size_t)dy * W + dx) * C));
        const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float a = __uint_as_float(w4[j] << 16), bq = __uint_as_float(w4[j] & 0xffff0000u);
          if (MAX) { acc[2 * j] = fmaxf(acc[2 * j], a); acc[2 * j + 1] = fmaxf(acc[2 * j + 1], bq); }
          else { acc[2 * j] += a; acc[2 * j + 1] += bq; }
        }
      }
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat162 hb = MAX ? __floats2bfloat162_rn(acc[2 * j], acc[2 * j + 1])
                                    : __floats2bfloat162_rn(acc[2 * j] * inv, acc[2 * j + 1] * inv);
      o[j] = *reinterpret_cast<const uint32_t*>(&hb);
    }
    *reinterpret_cast<uint4*>(out + (size_t)t * C + c8 * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

static int launch_pool_tokens(const void* x, void* out, int B, int H, int W, int C, int k, bool is_max,
                              cudaStream_t stream) {
  const int Ho = H / k, Wo = W / k;
  const long long n = (long long)B * Ho * Wo * (C / 8);
  HRF_REQUIRE(n < (1ll << 31), HRF_EUNSUPPORTED, "pool: problem too large");
  const int grid = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  const FastDiv d_cv(C / 8), d_wo(Wo), d_ho(Ho);
  if (is_max)
    HRF_CUDA(launch_pdl(pool_tokens_kernel<true>, dim3(grid), dim3(256), 0, stream, static_cast<const __nv_bfloat16*>(x),
                        static_cast<__nv_bfloat16*>(out), B, H, W, C, k, d_cv, d_wo, d_ho));
  else
    HRF_CUDA(launch_pdl(pool_tokens_kernel<false>, dim3(grid), dim3(256), 0, stream, static_cast<const __nv_bfloat16*>(x),
                        static_cast<__nv_bfloat16*>(out), B, H, W, C, k, d_cv, d_wo, d_ho));
  count_launch();
  HRF_CUDA(cudaGetLastError());
  return HRF_OK;
}

}  // namespace hrf
