// Training-mode window-attention core, fp32, forward and backward (SURVEY.md 8 a4 / a5 in the
// training configs, 8f rank 3): the part of WindowMSA / WindowMCA between the q / k / v projections
// and the output projection (reference hrformer.py:110-131, hrfuser_hrformer_based.py:127-151):
//
//     P = softmax(scale * Q K^T + rel_pos_bias)        O = P V
//
// per (window, head), with q / k / v / o as [nWin][N][C] token tensors whose head h lives in
// columns h*hd .. (h+1)*hd (what the Linear layers produce / consume: no head transposes).
// torch runs this as 2 batched SIMT GEMMs + bias add + softmax forward and 4 batched GEMMs + a
// softmax backward + an index_add backward (23 ms of fp32 SIMT GEMM alone in the 132 ms
// HRFuser-B training step, tools/train_profile.py).  Here one CTA owns a (window, head): every
// operand of it fits shared memory (N = 49, hd = 18 | 39).
//   forward : S -> softmax -> P (saved) -> O
//   backward: dV = P^T dO, dP = dO V^T, dS = P * (dP - rowsum(dP * P)), dQ = scale dS K,
//             dK = scale dS^T Q, dBias[h][i][j] = sum over windows of dS -- accumulated in
//             registers over the windows a CTA walks, written as per-CTA partials and summed in
//             fixed order (deterministic), then folded onto the (2Wh-1)(2Ww-1) table entries.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace hrf {

constexpr int kAtThreads = 128, kAtMaxN = 64, kAtMaxHd = 64;

struct AttnCoreTrain {
  const float* q; const float* k; const float* v;   // [nWin][N][C]
  const float* table;                               // [T][heads] or nullptr
  const int* rpi;                                   // [N*N] or nullptr
  float* o;                                         // [nWin][N][C]
  float* P;                                         // [nWin][heads][N][N]
  int nWin, N, C, heads, hd;
  float scale;
};

__host__ __device__ __forceinline__ int at_ld(int hd) { return hd | 1; }     // odd row stride: conflict-free column walks

// rows of one head's slice [N][hd] (row stride C in global) -> shared [N][ld]
__device__ __forceinline__ void at_load_tile(const float* __restrict__ g, int C, int N, int hd, int ld, float* s,
                                             float mul) {
  for (int e = threadIdx.x; e < N * hd; e += kAtThreads) {
    const int i = e / hd, d = e - i * hd;
    s[i * ld + d] = g[(size_t)i * C + d] * mul;
  }
}

// C(i, j) = sum_k A(i, k) B(k, j) on shared-memory operands with arbitrary strides
// (A(i,k) = A[i*sa_i + k*sa_k], B(k,j) = B[k*sb_k + j*sb_j]); a thread owns a TM x TN tile of C,
// so every k step is TM + TN shared-memory loads for TM * TN FMAs (the element-per-thread form
// needed two loads per FMA and was bound by the LSU: 80 / 153 us per call at 644 windows x 2 heads).
// Out-of-range rows / columns read a clamped address and are dropped by the epilogue.
template <int TM, int TN, typename Epi>
__device__ __forceinline__ void at_block_mm(const float* __restrict__ A, int sa_i, int sa_k, const float* __restrict__ B,
                                            int sb_k, int sb_j, int M, int N, int K, Epi epi) {
  const int tn = (N + TN - 1) / TN, tm = (M + TM - 1) / TM;
  for (int t = threadIdx.x; t < tm * tn; t += kAtThreads) {
    const int i0 = (t / tn) * TM, j0 = (t - (t / tn) * tn) * TN;
    int ia[TM], jb[TN];
#pragma unroll
    for (int r = 0; r < TM; ++r) ia[r] = (i0 + r < M ? i0 + r : M - 1) * sa_i;
#pragma unroll
    for (int c = 0; c < TN; ++c) jb[c] = (j0 + c < N ? j0 + c : N - 1) * sb_j;
    float acc[TM][TN];
#pragma unroll
    for (int r = 0; r < TM; ++r)
#pragma unroll
      for (int c = 0; c < TN; ++c) acc[r][c] = 0.f;
    for (int k = 0; k < K; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int r = 0; r < TM; ++r) a[r] = A[ia[r] + k * sa_k];
#pragma unroll
      for (int c = 0; c < TN; ++c) b[c] = B[jb[c] + k * sb_k];
#pragma unroll
      for (int r = 0; r < TM; ++r)
#pragma unroll
        for (int c = 0; c < TN; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
    }
#pragma unroll
    for (int r = 0; r < TM; ++r)
#pragma unroll
      for (int c = 0; c < TN; ++c)
        if (i0 + r < M && j0 + c < N) epi(i0 + r, j0 + c, acc[r][c]);
  }
}

__global__ void __launch_bounds__(kAtThreads) attn_core_train_fwd_kernel(AttnCoreTrain p) {
  extern __shared__ float at_sm[];
  const int N = p.N, hd = p.hd, ld = at_ld(hd), ldp = N | 1;
  float* sq = at_sm;
  float* sk = sq + N * ld;
  float* sv = sk + N * ld;
  float* sp = sv + N * ld;                        // [N][ldp]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int wh = blockIdx.x; wh < p.nWin * p.heads; wh += gridDim.x) {
    const int w = wh / p.heads, h = wh - w * p.heads;
    const size_t base = (size_t)w * N * p.C + h * hd;
    __syncthreads();
    at_load_tile(p.q + base, p.C, N, hd, ld, sq, p.scale);
    at_load_tile(p.k + base, p.C, N, hd, ld, sk, 1.f);
    at_load_tile(p.v + base, p.C, N, hd, ld, sv, 1.f);
    __syncthreads();
    // S = (scale Q) K^T + bias
    at_block_mm<7, 4>(sq, ld, 1, sk, 1, ld, N, N, hd, [&](int i, int j, float s) {
      if (p.table) s += __ldg(p.table + (size_t)__ldg(p.rpi + i * N + j) * p.heads + h);
      sp[i * ldp + j] = s;
    });
    __syncthreads();
    float* Pg = p.P + (size_t)wh * N * N;
    for (int i = warp; i < N; i += kAtThreads / 32) {       // softmax: a warp per row (N <= 64)
      const float a0 = lane < N ? sp[i * ldp + lane] : -INFINITY;
      const float a1 = lane + 32 < N ? sp[i * ldp + lane + 32] : -INFINITY;
      const float mx = warp_max(fmaxf(a0, a1));
      const float e0 = lane < N ? expf(a0 - mx) : 0.f, e1 = lane + 32 < N ? expf(a1 - mx) : 0.f;
      const float inv = 1.0f / warp_sum(e0 + e1);
      if (lane < N) { sp[i * ldp + lane] = e0 * inv; Pg[i * N + lane] = e0 * inv; }
      if (lane + 32 < N) { sp[i * ldp + lane + 32] = e1 * inv; Pg[i * N + lane + 32] = e1 * inv; }
    }
    __syncthreads();
    // O = P V
    float* og = p.o + base;
    const int Cc = p.C;
    if (hd > 24)
      at_block_mm<7, 4>(sp, ldp, 1, sv, ld, 1, N, hd, N, [&](int i, int d, float a) { og[(size_t)i * Cc + d] = a; });
    else
      at_block_mm<7, 2>(sp, ldp, 1, sv, ld, 1, N, hd, N, [&](int i, int d, float a) { og[(size_t)i * Cc + d] = a; });
  }
}

struct AttnCoreTrainBwd {
  const float* q; const float* k; const float* v; const float* P; const float* dout;
  float* dq; float* dk; float* dv;
  float* dbias_part;                                // [gridDim.x][N*N] per head-chunk partials or nullptr
  int nWin, N, C, heads, hd, chunks;                // grid = chunks * heads; CTA (c, h) walks windows c, c + chunks, ...
  float scale;
};

constexpr int kAtAcc = (kAtMaxN * kAtMaxN + kAtThreads - 1) / kAtThreads;   // dS elements per thread

__global__ void __launch_bounds__(kAtThreads) attn_core_train_bwd_kernel(AttnCoreTrainBwd p) {
  extern __shared__ float at_sm[];
  const int N = p.N, hd = p.hd, ld = at_ld(hd), ldp = N | 1;
  float* sq = at_sm;
  float* sk = sq + N * ld;
  float* sv = sk + N * ld;
  float* sdo = sv + N * ld;
  float* sp = sdo + N * ld;                       // P   [N][ldp]
  float* sds = sp + N * ldp;                      // dP -> dS [N][ldp]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x / p.heads, h = blockIdx.x - c * p.heads;
  float acc[kAtAcc];
#pragma unroll
  for (int r = 0; r < kAtAcc; ++r) acc[r] = 0.f;
  for (int w = c; w < p.nWin; w += p.chunks) {
    const size_t base = (size_t)w * N * p.C + h * hd;
    const float* Pg = p.P + ((size_t)w * p.heads + h) * N * N;
    __syncthreads();
    at_load_tile(p.q + base, p.C, N, hd, ld, sq, 1.f);
    at_load_tile(p.k + base, p.C, N, hd, ld, sk, 1.f);
    at_load_tile(p.v + base, p.C, N, hd, ld, sv, 1.f);
    at_load_tile(p.dout + base, p.C, N, hd, ld, sdo, 1.f);
    for (int e = threadIdx.x; e < N * N; e += kAtThreads) sp[(e / N) * ldp + (e % N)] = Pg[e];
    __syncthreads();
    // dP = dO V^T
    at_block_mm<7, 4>(sdo, ld, 1, sv, 1, ld, N, N, hd, [&](int i, int j, float s) { sds[i * ldp + j] = s; });
    __syncthreads();
    for (int i = warp; i < N; i += kAtThreads / 32) {               // dS = P (dP - sum_j dP P)
      const float p0 = lane < N ? sp[i * ldp + lane] : 0.f, p1 = lane + 32 < N ? sp[i * ldp + lane + 32] : 0.f;
      const float g0 = lane < N ? sds[i * ldp + lane] : 0.f, g1 = lane + 32 < N ? sds[i * ldp + lane + 32] : 0.f;
      const float r = warp_sum(fmaf(p0, g0, p1 * g1));
      if (lane < N) sds[i * ldp + lane] = p0 * (g0 - r);
      if (lane + 32 < N) sds[i * ldp + lane + 32] = p1 * (g1 - r);
    }
    __syncthreads();
    if (p.dbias_part) {
#pragma unroll
      for (int r = 0; r < kAtAcc; ++r) {
        const int e = threadIdx.x + r * kAtThreads;
        if (e < N * N) acc[r] += sds[(e / N) * ldp + (e % N)];
      }
    }
    {
      float* gq = p.dq + base;
      float* gk = p.dk + base;
      float* gv = p.dv + base;
      const int Cc = p.C;
      const float sc = p.scale;
      auto run = [&](auto tile) {
        constexpr int TN = decltype(tile)::value;
        // dQ = scale dS K ; dK = scale dS^T Q ; dV = P^T dO
        at_block_mm<7, TN>(sds, ldp, 1, sk, ld, 1, N, hd, N, [&](int i, int d, float a) { gq[(size_t)i * Cc + d] = a * sc; });
        at_block_mm<7, TN>(sds, 1, ldp, sq, ld, 1, N, hd, N, [&](int i, int d, float a) { gk[(size_t)i * Cc + d] = a * sc; });
        at_block_mm<7, TN>(sp, 1, ldp, sdo, ld, 1, N, hd, N, [&](int i, int d, float a) { gv[(size_t)i * Cc + d] = a; });
      };
      if (hd > 24) run(std::integral_constant<int, 4>{});
      else run(std::integral_constant<int, 2>{});
    }
  }
  if (p.dbias_part) {
    float* part = p.dbias_part + (size_t)blockIdx.x * N * N;
#pragma unroll
    for (int r = 0; r < kAtAcc; ++r) {
      const int e = threadIdx.x + r * kAtThreads;
      if (e < N * N) part[e] = acc[r];
    }
  }
}

// dS summed over the windows, then folded onto the table entries, both in fixed order:
//   (1) full[h][ij] = sum over chunks of the per-CTA partials (a thread per (h, ij): coalesced)
//   (2) dtable[t][h] = sum over the (i, j) with rpi[i][j] == t of full[h][ij]
// `full` reuses the first heads * N * N floats of the workspace's tail.
__global__ void __launch_bounds__(256) attn_core_train_dsum_kernel(const float* __restrict__ part, int chunks, int HNN,
                                                                    float* __restrict__ full) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= HNN) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int c = 0;
  for (; c + 4 <= chunks; c += 4) {                 // four independent loads in flight, fixed order
    a0 += part[(size_t)c * HNN + e];
    a1 += part[(size_t)(c + 1) * HNN + e];
    a2 += part[(size_t)(c + 2) * HNN + e];
    a3 += part[(size_t)(c + 3) * HNN + e];
  }
  for (; c < chunks; ++c) a0 += part[(size_t)c * HNN + e];
  full[e] = (a0 + a1) + (a2 + a3);
}
__global__ void __launch_bounds__(128) attn_core_train_dtable_kernel(const float* __restrict__ full, const int* __restrict__ rpi,
                                                                      int heads, int NN, int T, float* __restrict__ dtable) {
  // a warp per table entry (t, h): lanes stride over the N*N pairs, fixed-order shuffle sum
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (e >= T * heads) return;
  const int t = e / heads, h = e - t * heads;
  float a = 0.f;
  for (int ij = lane; ij < NN; ij += 32)
    if (__ldg(rpi + ij) == t) a += full[(size_t)h * NN + ij];
  a = warp_sum(a);
  if (lane == 0) dtable[e] = a;
}

static size_t attn_core_fwd_smem(int N, int hd) { return sizeof(float) * (3 * N * at_ld(hd) + N * (N | 1)); }
static size_t attn_core_bwd_smem(int N, int hd) { return sizeof(float) * (4 * N * at_ld(hd) + 2 * N * (N | 1)); }
static int attn_core_chunks(int nWin, int heads) {
  const int want = (148 * 4 + heads - 1) / heads;
  return nWin < want ? nWin : want;
}

static int launch_attn_core_train_fwd(const AttnCoreTrain& p, cudaStream_t st) {
  const size_t smem = attn_core_fwd_smem(p.N, p.hd);
  HRF_CUDA(ensure_smem((const void*)attn_core_train_fwd_kernel, smem));
  const int total = p.nWin * p.heads;
  attn_core_train_fwd_kernel<<<total < 148 * 8 ? total : 148 * 8, kAtThreads, smem, st>>>(p);
  count_launch();
  HRF_CUDA(cudaGetLastError());
  return HRF_OK;
}
static int launch_attn_core_train_bwd(AttnCoreTrainBwd p, const int* rpi, int T, float* dtable, cudaStream_t st) {
  const size_t smem = attn_core_bwd_smem(p.N, p.hd);
  HRF_CUDA(ensure_smem((const void*)attn_core_train_bwd_kernel, smem));
  p.chunks = attn_core_chunks(p.nWin, p.heads);
  attn_core_train_bwd_kernel<<<p.chunks * p.heads, kAtThreads, smem, st>>>(p);
  count_launch();
  HRF_CUDA(cudaGetLastError());
  if (p.dbias_part) {
    const int HNN = p.heads * p.N * p.N;
    float* full = p.dbias_part + (size_t)p.chunks * HNN;
    attn_core_train_dsum_kernel<<<ceil_div(HNN, 256), 256, 0, st>>>(p.dbias_part, p.chunks, HNN, full);
    count_launch();
    HRF_CUDA(cudaGetLastError());
    attn_core_train_dtable_kernel<<<ceil_div(T * p.heads * 32, 128), 128, 0, st>>>(full, rpi, p.heads, p.N * p.N, T, dtable);
    count_launch();
    HRF_CUDA(cudaGetLastError());
  }
  return HRF_OK;
}

}  // namespace hrf
