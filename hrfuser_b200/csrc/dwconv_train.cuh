// Training-mode depthwise 3x3 convolution (padding 1, stride 1 | 2), NCHW fp32: forward, input
// gradient, weight / bias gradient (the depthwise conv of CrossFFN,
// hrformer.py:273-279, and the stride-2 depthwise convs of the HRModule exchange,
// hrformer.py:528-548, in the training configs).  ATen's kernels take 39 / 65 / 64 us per call on
// these shapes (18.5 ms of the 123 ms HRFuser-B training step); the three passes are streaming
// stencils: a thread owns four consecutive outputs of a row, so the 3 x (3 s + 3) inputs it needs
// are loaded once and neighbouring threads read contiguous addresses.
//   forward : y = bias + sum_tap w[c][tap] x[.., oh s + dy - 1, ow s + dx - 1]
//   dgrad   : dx[ih][iw] = sum over the taps whose output position is integral of w[c][tap] g[..]
//   wgrad   : dw[c][tap] = sum_{b, oh, ow} g x[..],  dbias[c] = sum g  -- per-CTA partials per plane
//             chunk, summed in fixed order by a warp per output (deterministic, no atomics)
#pragma once
#include "common.cuh"

namespace hrf {

struct DwTrainParams {
  const float* x;      // forward / wgrad: input [B][C][H][W]; dgrad: unused
  const float* g;      // dgrad / wgrad: output gradient [B][C][Ho][Wo]
  const float* w;      // [C][9]
  const float* bias;   // [C] or nullptr
  float* out;          // forward: y [B][C][Ho][Wo]; dgrad: dx [B][C][H][W]
  float* part;         // wgrad partials [C][chunks][10]
  int B, C, H, W, Ho, Wo, chunks;
  FastDiv d_q, d_rows;   // quads per row, rows per plane (of the tensor the threads walk)
};

constexpr int kDwThreads = 256;

template <int S>
__global__ void __launch_bounds__(kDwThreads) dwconv_train_fwd_kernel(DwTrainParams p) {
  const int nq = (p.Wo + 3) / 4;
  const long long total = (long long)p.B * p.C * p.Ho * nq;
  for (long long e = (long long)blockIdx.x * kDwThreads + threadIdx.x; e < total; e += (long long)gridDim.x * kDwThreads) {
    int row, qx, plane, oh;
    p.d_q.divmod((int)e, row, qx);
    p.d_rows.divmod(row, plane, oh);
    const int c = plane % p.C, ow0 = qx * 4;
    float wt[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) wt[k] = __ldg(p.w + c * 9 + k);
    const float b0 = p.bias ? __ldg(p.bias + c) : 0.f;
    float acc[4] = {b0, b0, b0, b0};
    const float* xp = p.x + (size_t)plane * p.H * p.W;
    constexpr int NC = 3 * S + 3;                       // input columns of four outputs
    const int iw0 = ow0 * S - 1;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int ih = oh * S + dy - 1;
      if (ih < 0 || ih >= p.H) continue;
      float v[NC];
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        const int iw = iw0 + j;
        v[j] = (iw >= 0 && iw < p.W) ? __ldg(xp + (size_t)ih * p.W + iw) : 0.f;
      }
#pragma unroll
      for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) acc[o] = fmaf(v[o * S + dx], wt[dy * 3 + dx], acc[o]);
    }
    float* yp = p.out + ((size_t)plane * p.Ho + oh) * p.Wo + ow0;
    if (ow0 + 3 < p.Wo && (p.Wo & 3) == 0) {
      *reinterpret_cast<float4*>(yp) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    } else {
#pragma unroll
      for (int o = 0; o < 4; ++o)
        if (ow0 + o < p.Wo) yp[o] = acc[o];
    }
  }
}

// input gradient: threads walk the INPUT tensor [B][C][H][W], four consecutive iw per thread
template <int S>
__global__ void __launch_bounds__(kDwThreads) dwconv_train_dgrad_kernel(DwTrainParams p) {
  const int nq = (p.W + 3) / 4;
  const long long total = (long long)p.B * p.C * p.H * nq;
  for (long long e = (long long)blockIdx.x * kDwThreads + threadIdx.x; e < total; e += (long long)gridDim.x * kDwThreads) {
    int row, qx, plane, ih;
    p.d_q.divmod((int)e, row, qx);
    p.d_rows.divmod(row, plane, ih);
    const int c = plane % p.C, iw0 = qx * 4;
    float wt[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) wt[k] = __ldg(p.w + c * 9 + k);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float* gp = p.g + (size_t)plane * p.Ho * p.Wo;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int t = ih + 1 - dy;                       // oh * S
      if (t < 0 || (S == 2 && (t & 1))) continue;
      const int oh = t / S;
      if (oh >= p.Ho) continue;
      const float* gr = gp + (size_t)oh * p.Wo;
#pragma unroll
      for (int o = 0; o < 4; ++o) {
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int u = iw0 + o + 1 - dx;              // ow * S
          if (u < 0 || (S == 2 && (u & 1))) continue;
          const int ow = u / S;
          if (ow < p.Wo) acc[o] = fmaf(__ldg(gr + ow), wt[dy * 3 + dx], acc[o]);
        }
      }
    }
    float* op = p.out + ((size_t)plane * p.H + ih) * p.W + iw0;
    if (iw0 + 3 < p.W && (p.W & 3) == 0) {
      *reinterpret_cast<float4*>(op) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    } else {
#pragma unroll
      for (int o = 0; o < 4; ++o)
        if (iw0 + o < p.W) op[o] = acc[o];
    }
  }
}

// weight / bias gradient: CTA (c, chunk) walks the output quads of channel c that fall to its chunk
// over all images; per-thread accumulators -> warp shuffles -> shared memory -> part[c][chunk][10]
template <int S>
__global__ void __launch_bounds__(kDwThreads) dwconv_train_wgrad_kernel(DwTrainParams p) {
  __shared__ float red[kDwThreads / 32][10];
  const int c = blockIdx.x / p.chunks, chunk = blockIdx.x - c * p.chunks;
  const int nq = (p.Wo + 3) / 4;
  const int per_img = p.Ho * nq, total = p.B * per_img;
  float acc[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) acc[k] = 0.f;
  constexpr int NC = 3 * S + 3;
  for (int e = chunk * kDwThreads + threadIdx.x; e < total; e += p.chunks * kDwThreads) {
    const int b = e / per_img, r = e - b * per_img;
    int oh, qx;
    p.d_q.divmod(r, oh, qx);
    const int ow0 = qx * 4, plane = b * p.C + c;
    const float* gp = p.g + ((size_t)plane * p.Ho + oh) * p.Wo + ow0;
    float gv[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) gv[o] = ow0 + o < p.Wo ? __ldg(gp + o) : 0.f;
    acc[9] += (gv[0] + gv[1]) + (gv[2] + gv[3]);
    const float* xp = p.x + (size_t)plane * p.H * p.W;
    const int iw0 = ow0 * S - 1;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int ih = oh * S + dy - 1;
      if (ih < 0 || ih >= p.H) continue;
      float v[NC];
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        const int iw = iw0 + j;
        v[j] = (iw >= 0 && iw < p.W) ? __ldg(xp + (size_t)ih * p.W + iw) : 0.f;
      }
#pragma unroll
      for (int dx = 0; dx < 3; ++dx)
#pragma unroll
        for (int o = 0; o < 4; ++o) acc[dy * 3 + dx] = fmaf(gv[o], v[o * S + dx], acc[dy * 3 + dx]);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    const float s = warp_sum(acc[k]);
    if (lane == 0) red[warp][k] = s;
  }
  __syncthreads();
  if (threadIdx.x < 10) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kDwThreads / 32; ++w) s += red[w][threadIdx.x];
    p.part[((size_t)c * p.chunks + chunk) * 10 + threadIdx.x] = s;
  }
}

// dw[c][k] (k < 9), dbias[c] (k == 9): a warp per output, fixed-order sum over the chunks
__global__ void __launch_bounds__(256) dwconv_train_wgrad_finalize_kernel(const float* __restrict__ part, int C, int chunks,
                                                                           float* __restrict__ dw, float* __restrict__ dbias) {
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (e >= C * 10) return;
  const int c = e / 10, k = e - c * 10;
  float a = 0.f;
  for (int q = lane; q < chunks; q += 32) a += part[((size_t)c * chunks + q) * 10 + k];
  a = warp_sum(a);
  if (lane == 0) {
    if (k < 9) dw[c * 9 + k] = a;
    else if (dbias) dbias[c] = a;
  }
}

static int dw_train_chunks(int B, int C, int Ho, int Wo) {
  const int work = ceil_div(B * Ho * ((Wo + 3) / 4), kDwThreads);        // CTA-sized pieces per channel
  const int want = ceil_div(148 * 8, C);                                   // enough CTAs to fill the GPU
  const int n = work < want ? work : want;
  return n < 1 ? 1 : n;
}
static size_t dw_train_ws_floats(int B, int C, int Ho, int Wo) { return (size_t)C * dw_train_chunks(B, C, Ho, Wo) * 10; }

static int dw_grid(long long total) {
  const long long need = (total + kDwThreads - 1) / kDwThreads;
  return (int)(need < 148 * 16 ? need : 148 * 16);
}

template <int S>
static int launch_dw_train(int pass, DwTrainParams p, float* dw, float* dbias, cudaStream_t st) {
  if (pass == 0) {
    p.d_q = FastDiv((p.Wo + 3) / 4);
    p.d_rows = FastDiv(p.Ho);
    dwconv_train_fwd_kernel<S><<<dw_grid((long long)p.B * p.C * p.Ho * ((p.Wo + 3) / 4)), kDwThreads, 0, st>>>(p);
  } else if (pass == 1) {
    p.d_q = FastDiv((p.W + 3) / 4);
    p.d_rows = FastDiv(p.H);
    dwconv_train_dgrad_kernel<S><<<dw_grid((long long)p.B * p.C * p.H * ((p.W + 3) / 4)), kDwThreads, 0, st>>>(p);
  } else {
    p.d_q = FastDiv((p.Wo + 3) / 4);
    p.chunks = dw_train_chunks(p.B, p.C, p.Ho, p.Wo);
    dwconv_train_wgrad_kernel<S><<<p.C * p.chunks, kDwThreads, 0, st>>>(p);
    count_launch();
    HRF_CUDA(cudaGetLastError());
    dwconv_train_wgrad_finalize_kernel<<<ceil_div(p.C * 10 * 32, 256), 256, 0, st>>>(p.part, p.C, p.chunks, dw, dbias);
  }
  count_launch();
  HRF_CUDA(cudaGetLastError());
  return HRF_OK;
}

}  // namespace hrf
