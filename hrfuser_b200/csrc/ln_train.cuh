// Train-mode LayerNorm over the channel axis of token tensors [rows][C], fp32 -- forward and
// backward (the reference's nn.LayerNorm of hrformer.py:345-371 and
// hrfuser_hrformer_based.py:262-313 in the training configs).  ATen's kernels spend 145 us per
// call on these shapes (30 720 rows of 78 channels: rows far narrower than their block size);
// here a warp owns a row, keeps its channels in registers (VPT = ceil(C / 32) per lane), and the
// kernels stream at HBM rate.
//   forward : y = (x - mean) * rstd * gamma + beta; mean / rstd saved for the backward
//   backward: dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma;
//             dgamma = sum_rows dy * xhat, dbeta = sum_rows dy  -- per-lane register
//             accumulators over the warp's rows -> shared memory per CTA -> partials [CTA][2][C]
//             -> a fixed-order sum over the CTAs (deterministic: no atomics)
#pragma once
#include "common.cuh"

namespace hrf {

constexpr int kLnThreads = 256, kLnWarps = kLnThreads / 32;

template <int VPT>
__global__ void __launch_bounds__(kLnThreads) ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ y,
                                                            float* __restrict__ mean_o, float* __restrict__ rstd_o, int rows,
                                                            int C, float eps) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float g[VPT], b[VPT];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = lane + 32 * i;
    g[i] = c < C ? __ldg(gamma + c) : 0.f;
    b[i] = c < C ? __ldg(beta + c) : 0.f;
  }
  const float inv_c = 1.0f / (float)C;
  for (int r = blockIdx.x * kLnWarps + warp; r < rows; r += gridDim.x * kLnWarps) {
    const float* xr = x + (size_t)r * C;
    float v[VPT], s = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int c = lane + 32 * i;
      v[i] = c < C ? xr[c] : 0.f;
      s += v[i];
    }
    const float mean = warp_sum(s) * inv_c;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const float d = (lane + 32 * i < C) ? v[i] - mean : 0.f;
      q = fmaf(d, d, q);
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_c + eps);
    float* yr = y + (size_t)r * C;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int c = lane + 32 * i;
      if (c < C) yr[c] = fmaf((v[i] - mean) * rstd, g[i], b[i]);
    }
    if (lane == 0) {
      mean_o[r] = mean;
      rstd_o[r] = rstd;
    }
  }
}

// partial: [gridDim.x][2][C]
template <int VPT>
__global__ void __launch_bounds__(kLnThreads) ln_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                            const float* __restrict__ mean_i, const float* __restrict__ rstd_i,
                                                            const float* __restrict__ gamma, float* __restrict__ dx,
                                                            float* __restrict__ partial, int rows, int C) {
  extern __shared__ float sred[];                       // [kLnWarps][2][C]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float g[VPT], dg[VPT], db[VPT];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = lane + 32 * i;
    g[i] = c < C ? __ldg(gamma + c) : 0.f;
    dg[i] = 0.f;
    db[i] = 0.f;
  }
  const float inv_c = 1.0f / (float)C;
  for (int r = blockIdx.x * kLnWarps + warp; r < rows; r += gridDim.x * kLnWarps) {
    const float* xr = x + (size_t)r * C;
    const float* dr = dy + (size_t)r * C;
    const float mean = mean_i[r], rstd = rstd_i[r];
    float xh[VPT], gy[VPT], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int c = lane + 32 * i;
      const bool in = c < C;
      const float d = in ? dr[c] : 0.f;
      xh[i] = in ? (xr[c] - mean) * rstd : 0.f;
      gy[i] = d * g[i];
      s1 += gy[i];
      s2 = fmaf(gy[i], xh[i], s2);
      dg[i] = fmaf(d, xh[i], dg[i]);
      db[i] += d;
    }
    s1 = warp_sum(s1) * inv_c;
    s2 = warp_sum(s2) * inv_c;
    if (dx != nullptr) {
      float* or_ = dx + (size_t)r * C;
#pragma unroll
      for (int i = 0; i < VPT; ++i) {
        const int c = lane + 32 * i;
        if (c < C) or_[c] = rstd * (gy[i] - s1 - xh[i] * s2);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = lane + 32 * i;
    if (c < C) {
      sred[(warp * 2 + 0) * C + c] = dg[i];
      sred[(warp * 2 + 1) * C + c] = db[i];
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 2 * C; e += kLnThreads) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < kLnWarps; ++w) a += sred[w * 2 * C + e];
    partial[(size_t)blockIdx.x * 2 * C + e] = a;
  }
}

__global__ void __launch_bounds__(256) ln_bwd_finalize_kernel(const float* __restrict__ partial, int n_part, int C,
                                                              float* __restrict__ dgamma, float* __restrict__ dbeta) {
  // a warp per output element: lanes stride over the CTA partials, fixed-order shuffle sum
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (e >= 2 * C) return;
  float a = 0.f;
  for (int p = lane; p < n_part; p += 32) a += partial[(size_t)p * 2 * C + e];
  a = warp_sum(a);
  if (lane == 0) {
    if (e < C) dgamma[e] = a;
    else dbeta[e - C] = a;
  }
}

static int ln_grid(int rows) {
  const int need = ceil_div(rows, kLnWarps);
  return need < 148 * 4 ? need : 148 * 4;
}
static size_t ln_bwd_workspace_floats(int rows, int C) { return (size_t)ln_grid(rows) * 2 * C; }

template <int VPT>
static int launch_ln_fwd_v(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                           int rows, int C, float eps, cudaStream_t st) {
  ln_fwd_kernel<VPT><<<ln_grid(rows), kLnThreads, 0, st>>>(x, gamma, beta, y, mean, rstd, rows, C, eps);
  count_launch();
  HRF_CUDA(cudaGetLastError());
  return HRF_OK;
}
template <int VPT>
static int launch_ln_bwd_v(const float* x, const float* dy, const float* mean, const float* rstd, const float* gamma,
                           float* dx, float* partial, float* dgamma, float* dbeta, int rows, int C, cudaStream_t st) {
  const int grid = ln_grid(rows);
  const size_t smem = (size_t)kLnWarps * 2 * C * sizeof(float);
  HRF_CUDA(ensure_smem((const void*)ln_bwd_kernel<VPT>, smem));
  ln_bwd_kernel<VPT><<<grid, kLnThreads, smem, st>>>(x, dy, mean, rstd, gamma, dx, partial, rows, C);
  count_launch();
  HRF_CUDA(cudaGetLastError());
  ln_bwd_finalize_kernel<<<ceil_div(2 * C * 32, 256), 256, 0, st>>>(partial, grid, C, dgamma, dbeta);
  count_launch();
  HRF_CUDA(cudaGetLastError());
  return HRF_OK;
}

#define HRF_LN_DISPATCH(fn, ...)                                      \
  do {                                                                \
    const int vpt_ = ceil_div(C, 32);                                 \
    if (vpt_ <= 1) return fn<1>(__VA_ARGS__);                         \
    if (vpt_ <= 2) return fn<2>(__VA_ARGS__);                         \
    if (vpt_ <= 3) return fn<3>(__VA_ARGS__);                         \
    if (vpt_ <= 5) return fn<5>(__VA_ARGS__);                         \
    if (vpt_ <= 10) return fn<10>(__VA_ARGS__);                       \
    if (vpt_ <= 20) return fn<20>(__VA_ARGS__);                       \
    return fn<32>(__VA_ARGS__);                                       \
  } while (0)

static int launch_ln_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                         int rows, int C, float eps, cudaStream_t st) {
  HRF_LN_DISPATCH(launch_ln_fwd_v, x, gamma, beta, y, mean, rstd, rows, C, eps, st);
}
static int launch_ln_bwd(const float* x, const float* dy, const float* mean, const float* rstd, const float* gamma,
                         float* dx, float* partial, float* dgamma, float* dbeta, int rows, int C, cudaStream_t st) {
  HRF_LN_DISPATCH(launch_ln_bwd_v, x, dy, mean, rstd, gamma, dx, partial, dgamma, dbeta, rows, C, st);
}

}  // namespace hrf
