// Generic (any width) un-fused path for window attention and MixFFN.
//
// The fused kernels keep a whole window / tile in shared memory, which caps the channel
// count (C <= ~250 for the SIMT kernels, head_dim 18 for the tcgen05 ones).  HRFuser-B's
// low-resolution branches (C = 312, 624; 16 heads of 39) go through this path instead:
//   LayerNorm rows -> tiled fp32 GEMMs (+bias, activation, residuals) -> per-(window, head)
//   attention core that gathers its window slots from the token tensors (pad slots contribute
//   k = b_k, v = b_v exactly like the reference's zero padding after the norm).  In bf16
//   mode the GEMMs and the core (attn_core_tc_kernel: QK^T and PV as UMMAs, any window up to
//   16 x 16) run on tcgen05; this is also the bf16 path of every window size other than 7.
// Same packed fp32 weight sections as the fused SIMT kernels; intermediates live in a
// caller-provided fp32 workspace.  Accuracy-first (fp32 FFMA, erff, expf).
#pragma once
#include <type_traits>

#include "common.cuh"
#include "mixffn.cuh"
#include "umma.cuh"
#include "window_attn.cuh"

namespace hrf {

// ---- LayerNorm of token rows: (n_tok, C) T -> fp32 -------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) ln_rows_kernel(const T* x, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, float eps,
                                                      float* out, int n_tok, int C) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_tok) return;
  const T* src = x + (size_t)warp * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += Elem<T>::ld(src + c);
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
  for (int c = lane; c < C; c += 32) { const float d = Elem<T>::ld(src + c) - mean; q = fmaf(d, d, q); }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
  float* dst = out + (size_t)warp * C;
  for (int c = lane; c < C; c += 32)
    dst[c] = (Elem<T>::ld(src + c) - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
}

// ---- tiled fp32 GEMM: out = act(A[M,K] * Wt[K,N] + bias) (+ r1) (+ r2) ---------------------
// A fp32 row-major with leading dimension lda; Wt k-major with leading dimension N.
// act: 0 none, 1 ReLU, 2 GELU(erf).  out / residuals have leading dimension ldo.
struct GemmParams {
  const float* A; const float* Wt; const float* bias;
  const void* r1; const void* r2; void* out;
  int M, N, K, lda, ldo, act;
};

template <typename TO>
__global__ void __launch_bounds__(256) gemm_bias_act_kernel(GemmParams p) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tr = (tid / 16) * 4, tc = (tid % 16) * 4;      // 4x4 outputs per thread
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < p.K; k0 += BK) {
    for (int e = tid; e < BM * BK; e += 256) {             // A tile, transposed into smem
      const int r = e / BK, k = e % BK;
      As[k][r] = (m0 + r < p.M && k0 + k < p.K) ? __ldg(p.A + (size_t)(m0 + r) * p.lda + k0 + k) : 0.f;
    }
    for (int e = tid; e < BK * BN; e += 256) {
      const int k = e / BN, n = e % BN;
      Bs[k][n] = (k0 + k < p.K && n0 + n < p.N) ? __ldg(p.Wt + (size_t)(k0 + k) * p.N + n0 + n) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tc]);
      const float a[4] = {As[k][tr], As[k][tr + 1], As[k][tr + 2], As[k][tr + 3]};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i][0] = fmaf(a[i], b.x, acc[i][0]); acc[i][1] = fmaf(a[i], b.y, acc[i][1]);
        acc[i][2] = fmaf(a[i], b.z, acc[i][2]); acc[i][3] = fmaf(a[i], b.w, acc[i][3]);
      }
    }
    __syncthreads();
  }
  const TO* r1 = static_cast<const TO*>(p.r1);
  const TO* r2 = static_cast<const TO*>(p.r2);
  TO* out = static_cast<TO*>(p.out);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + tr + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tc + j;
      if (n >= p.N) continue;
      float v = acc[i][j] + (p.bias ? __ldg(p.bias + n) : 0.f);
      if (p.act == 1) v = fmaxf(v, 0.f);
      else if (p.act == 2) v = gelu_erf(v);
      const size_t off = (size_t)m * p.ldo + n;
      if (r1) v += Elem<TO>::ld(r1 + off);
      if (r2) v += Elem<TO>::ld(r2 + off);
      Elem<TO>::st(out + off, v);
    }
  }
}

// ---- the same GEMM on the tcgen05 tensor cores (bf16 mode) ------------------------------------
// out[128 x 128 tile] over K blocks of 64: A (fp32 activations) and Wt (fp32 k-major weights,
// the sections the SIMT kernels use) are converted to bf16 operand tiles on the way into
// shared memory -- no extra packed copies of the wide branches' weights -- two stages, so the
// next block's global loads and conversion overlap the MMAs of the current one.  fp32
// accumulation in TMEM; bias / activation / residual epilogue as above.
template <typename TO>
__global__ void __launch_bounds__(256) gemm_tc_kernel(GemmParams p) {
  using namespace umma;
  constexpr int NT = 128, KB = 64, A_B = 128 * KB * 2, B_B = NT * KB * 2;
  extern __shared__ __align__(128) unsigned char sm[];          // 2 x (A tile | B tile)
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = warp_idx_uniform(), lane = tid & 31;
  const int m0 = blockIdx.y * 128, n0 = blockIdx.x * NT;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, NT);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const int nkb = ceil_div(p.K, KB);
  uint32_t ph[2] = {0u, 0u};
  // A: thread -> (row, 32 consecutive k); B: thread -> (n, 32 consecutive k)
  const int a_row = tid >> 1, a_k = (tid & 1) * 32;
  const int b_n = tid & 127, b_k = (tid >> 7) * 32;
  const bool a_ok = m0 + a_row < p.M, b_ok = n0 + b_n < p.N;
  const float* a_src = p.A + (size_t)(a_ok ? m0 + a_row : 0) * p.lda;
  const float* b_src = p.Wt + (b_ok ? n0 + b_n : 0);
  // the global loads of block kb + 1 are requested before block kb is converted / issued, so their
  // latency hides behind the conversion, the barrier and the MMAs (it was exposed once per block)
  float a[32], b[32];
  auto load_block = [&](int kb) {
    const int k0 = kb * KB;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {                // K % 4 == 0 (checked by the launcher)
      const int k = k0 + a_k + j;
      const float4 v = (a_ok && k < p.K) ? __ldg(reinterpret_cast<const float4*>(a_src + k))
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
      a[j] = v.x; a[j + 1] = v.y; a[j + 2] = v.z; a[j + 3] = v.w;
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int k = k0 + b_k + j;
      b[j] = (b_ok && k < p.K) ? __ldg(b_src + (size_t)k * p.N) : 0.f;
    }
  };
  load_block(0);
#pragma unroll 1
  for (int kb = 0; kb < nkb; ++kb) {
    const int buf = kb & 1;
    if (kb >= 2) {                                   // the MMAs of block kb-2 read this stage
      cta_wait(&bars[buf], ph[buf]);
      ph[buf] ^= 1;
    }
    unsigned char* sA = sm + buf * (A_B + B_B);
    unsigned char* sB = sA + A_B;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      st_chunk(sA, a_row, a_k / 8 + c, 128, a + 8 * c);
      st_chunk(sB, b_n, b_k / 8 + c, NT, b + 8 * c);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (warp == 0 && elect_one()) {
      tc_fence_after();
      constexpr uint32_t id = idesc_bf16(128, NT, false, false);
      const uint32_t aA = smem_u32(sA), aB = smem_u32(sB);
#pragma unroll
      for (int s2 = 0; s2 < KB / 16; ++s2)
        mma_bf16(tmem, desc_kmajor(aA, 128, s2), desc_kmajor(aB, NT, s2), id, kb > 0 || s2 > 0);
      mma_commit(&bars[buf]);
    }
    // (behind the fence: fence.proxy.async's MEMBAR would wait for loads requested in front of it)
    if (kb + 1 < nkb) load_block(kb + 1);
  }
  {                                                  // the last commit covers every MMA
    const int buf = (nkb - 1) & 1;
    cta_wait(&bars[buf], ph[buf]);
  }
  tc_fence_after();
  // ---- epilogue: warp w -> TMEM quadrant w % 4, column group w / 4 --------------------------
  const int q = warp & 3, gq = warp >> 2;
  const int m = m0 + q * 32 + lane;
  const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
  const TO* r1 = static_cast<const TO*>(p.r1);
  const TO* r2 = static_cast<const TO*>(p.r2);
  TO* out = static_cast<TO*>(p.out);
#pragma unroll 1
  for (int cc = gq; cc < NT / 8; cc += 2) {
    float y[8];
    tmem_ld8(trow + cc * 8, y);
    tmem_ld_wait();
    if (m < p.M) {
      const int nb = n0 + cc * 8;
      if (nb + 8 <= p.N && p.ldo % 8 == 0 && p.N % 8 == 0) {
        // eight consecutive outputs of a row as 16- / 32-byte vectors (scalar stores 512 bytes apart
        // between lanes cost a sector per element)
        const size_t off = (size_t)m * p.ldo + nb;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[j] = y[j] + (p.bias ? __ldg(p.bias + nb + j) : 0.f);
          if (p.act == 1) v[j] = fmaxf(v[j], 0.f);
          else if (p.act == 2) v[j] = gelu_erf(v[j]);
        }
        if constexpr (std::is_same<TO, float>::value) {
          if (r1) { const float4 a = *reinterpret_cast<const float4*>(r1 + off), b = *reinterpret_cast<const float4*>(r1 + off + 4);
                    v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w; }
          if (r2) { const float4 a = *reinterpret_cast<const float4*>(r2 + off), b = *reinterpret_cast<const float4*>(r2 + off + 4);
                    v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w; }
          *reinterpret_cast<float4*>(out + off) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(out + off + 4) = make_float4(v[4], v[5], v[6], v[7]);
        } else {
          float t[8];
          if (r1) { load_row_bf16<8>(r1 + off, t);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] += t[j]; }
          if (r2) { load_row_bf16<8>(r2 + off, t);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] += t[j]; }
          store_row_bf16<8>(out + off, v);
        }
      } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int n = n0 + cc * 8 + j;
        if (n < p.N) {
          float v = y[j] + (p.bias ? __ldg(p.bias + n) : 0.f);
          if (p.act == 1) v = fmaxf(v, 0.f);
          else if (p.act == 2) v = gelu_erf(v);
          const size_t off = (size_t)m * p.ldo + n;
          if (r1) v += Elem<TO>::ld(r1 + off);
          if (r2) v += Elem<TO>::ld(r2 + off);
          Elem<TO>::st(out + off, v);
        }
      }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, NT);
}

// tc: bf16 mode (the caller's activations are bf16, so bf16 operand precision is the contract)
template <typename TO>
static int launch_gemm(const GemmParams& p, cudaStream_t stream, bool tc = false) {
  if (tc && p.K % 4 == 0 && p.lda % 4 == 0 && (reinterpret_cast<uintptr_t>(p.A) & 15) == 0 && !tc_disabled()) {
    dim3 grid(ceil_div(p.N, 128), ceil_div(p.M, 128));
    HRF_REQUIRE(grid.y <= 65535, HRF_EUNSUPPORTED, "gemm: M=%d too large", p.M);
    constexpr size_t smem = 2 * (128 * 64 * 2 + 128 * 64 * 2);
    HRF_CUDA(ensure_smem((const void*)gemm_tc_kernel<TO>, smem));
    gemm_tc_kernel<TO><<<grid, 256, smem, stream>>>(p);
    count_launch();
    HRF_CUDA(cudaGetLastError());
    return HRF_OK;
  }
  dim3 grid(ceil_div(p.N, 64), ceil_div(p.M, 64));
  HRF_REQUIRE(grid.y <= 65535, HRF_EUNSUPPORTED, "gemm: M=%d too large", p.M);
  gemm_bias_act_kernel<TO><<<grid, 256, 0, stream>>>(p);
  count_launch();
  HRF_CUDA(cudaGetLastError());
  return HRF_OK;
}

// ---- attention core: one CTA (4 warps) per (window, head) ---------------------------------
// q, k, v: fp32 (n_tok, C) already projected (q pre-scaled); o: fp32 (n_tok, KO) head padded.
struct CoreParams {
  const float* q; const float* k; const float* v; float* o;
  const float* bk; const float* bv;      // projection biases: keys / values of pad slots
  const float* rpb;                      // (heads, (2w-1)^2)
  int B, H, W, C, heads, win, KO, hdp, pad_mask;
};

__global__ void __launch_bounds__(128) attn_core_kernel(CoreParams p) {
  extern __shared__ __align__(16) float core_sm[];
  float* sm = core_sm;
  const int win = p.win, S = win * win, hd = p.C / p.heads;
  const int ld = hd | 1;                                   // odd stride: conflict-free row walks
  float* Q = sm; float* Kt = Q + S * ld; float* V = Kt + S * ld;
  float* P = V + S * ld;                                   // [4 warps][S]
  int* tok = reinterpret_cast<int*>(P + 4 * S);
  const int h = blockIdx.y, wdx = blockIdx.x;
  const int nWh = ceil_div(p.H, win), nWw = ceil_div(p.W, win);
  const int pad_h = nWh * win - p.H, pad_w = nWw * win - p.W;
  const bool use_mask = p.pad_mask && pad_h > 0 && pad_w > 0;
  const int b = wdx / (nWh * nWw), wy = (wdx / nWw) % nWh, wx = wdx % nWw;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const int y = wy * win + s / win - pad_h / 2, x = wx * win + s % win - pad_w / 2;
    tok[s] = (y >= 0 && y < p.H && x >= 0 && x < p.W) ? (b * p.H + y) * p.W + x : -1;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < S * hd; e += blockDim.x) {
    const int s = e / hd, d = e - s * hd, t = tok[s], c = h * hd + d;
    Q[s * ld + d] = t >= 0 ? __ldg(p.q + (size_t)t * p.C + c) : 0.f;
    Kt[s * ld + d] = t >= 0 ? __ldg(p.k + (size_t)t * p.C + c) : __ldg(p.bk + c);
    V[s * ld + d] = t >= 0 ? __ldg(p.v + (size_t)t * p.C + c) : __ldg(p.bv + c);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* Pw = P + warp * S;
  const float* rp = p.rpb + (size_t)h * (2 * win - 1) * (2 * win - 1);
  for (int i = warp; i < S; i += 4) {
    if (tok[i] < 0) continue;                              // pad queries are cropped anyway
    const int ih = i / win, iw = i - ih * win;
    float mx = -INFINITY;
    for (int j = lane; j < S; j += 32) {
      float s = 0.f;
      for (int d = 0; d < hd; ++d) s = fmaf(Q[i * ld + d], Kt[j * ld + d], s);
      const int jh = j / win, jw = j - jh * win;
      s += __ldg(rp + (ih - jh + win - 1) * (2 * win - 1) + (iw - jw + win - 1));
      if (use_mask && tok[j] < 0) s = -INFINITY;
      Pw[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < S; j += 32) { const float e = expf(Pw[j] - mx); Pw[j] = e; sum += e; }
    const float inv = 1.0f / warp_sum(sum);
    __syncwarp();
    for (int d = lane; d < hd; d += 32) {
      float o = 0.f;
      for (int j = 0; j < S; ++j) o = fmaf(Pw[j], V[j * ld + d], o);
      p.o[(size_t)tok[i] * p.KO + h * p.hdp + d] = o * inv;
    }
    __syncwarp();
  }
}

// ---- depthwise 3x3 (stride 1, pad 1) + bias + GELU on fp32 (n_tok, CH) ------------------------
// A thread owns four consecutive channels of a token (CH % 4 == 0: float4 loads of the nine taps and
// of their weights), 32-bit index arithmetic.  FAST: the tanh-form GELU of the bf16 mode (the
// fused kernels' `gelu_as`); the fp32 mode keeps erf.  (The element-per-thread version with 64-bit
// `/ %` ran at 0.56 TB/s: 135 us for the 24 x 40 x 8 x 1248 hidden tensor of HRFuser-B.)
template <bool FAST>
__global__ void __launch_bounds__(256) dw3x3_gelu_kernel(const float* x, const float* __restrict__ wd,
                                                         const float* __restrict__ bd, float* out,
                                                         int B, int H, int W, int CH) {
  const int cq = CH >> 2;
  const unsigned total = (unsigned)B * H * W * cq;
  for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const unsigned t = e / cq;
    const int c = (int)(e - t * cq) * 4;
    const unsigned hb = t / W;
    const int w = (int)(t - hb * W), b = (int)(hb / H), h = (int)(hb - (unsigned)b * H);
    float4 s = __ldg(reinterpret_cast<const float4*>(bd + c));
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int y = h + dy;
      if (y < 0 || y >= H) continue;
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int xx = w + dx;
        if (xx < 0 || xx >= W) continue;
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + ((size_t)(b * H + y) * W + xx) * CH + c));
        const float4 k = __ldg(reinterpret_cast<const float4*>(wd + (size_t)((dy + 1) * 3 + dx + 1) * CH + c));
        s.x = fmaf(v.x, k.x, s.x); s.y = fmaf(v.y, k.y, s.y); s.z = fmaf(v.z, k.z, s.z); s.w = fmaf(v.w, k.w, s.w);
      }
    }
    float4 o;
    if (FAST) { o.x = gelu_as(s.x); o.y = gelu_as(s.y); o.z = gelu_as(s.z); o.w = gelu_as(s.w); }
    else { o.x = gelu_erf(s.x); o.y = gelu_erf(s.y); o.z = gelu_erf(s.z); o.w = gelu_erf(s.w); }
    *reinterpret_cast<float4*>(out + (size_t)t * CH + c) = o;
  }
}

// ---- drivers --------------------------------------------------------------------------------
static size_t attn_generic_ws_floats(int B, int H, int W, int C, int heads, int win, bool cross) {
  const AttnLayout L(C, heads, win);
  const size_t n = (size_t)B * H * W;
  return n * C * (cross ? 5 : 4) + n * L.KO;      // xn (zn) q k v + o
}

// ---- attention core on the tensor cores (bf16 mode): one CTA per (window, head) ------------
// Any window with S = win^2 <= 256 slots and any head_dim: S is padded to SP (multiple of 16,
// the UMMA N / K granule), head_dim to HDP.  Per 128-row query tile:
//   UMMA  S = Q K^T   (M 128, N SP, K HDP; Q and K K-major tiles)          -> TMEM
//   rows: + relative position bias (table in shared memory), mask, max, exp, sum; the
//         unnormalised P goes back to shared memory as a bf16 K-major tile
//   UMMA  O = P V     (M 128, N HDP, K SP; V is the K-layout tile described MN-major)
//         into the TMEM columns S occupied (every row has been read by then)
//   rows: x 1/sum -> o (fp32, head padded), only for real tokens
// Pad slots inside the window carry k = b_k, v = b_v (the reference pads after the norm).
struct CoreTcGeom {
  int S, SP, HDP, MT, T;
  uint32_t tmem_cols;
  size_t q_bytes, kv_bytes, p_bytes, smem;
  FastDiv hdiv;            // by head_dim
};
static CoreTcGeom core_tc_geom(int win, int hd) {
  CoreTcGeom g;
  g.S = win * win;
  g.SP = round_up(g.S, 16);
  g.HDP = round_up(hd, 16);
  g.MT = ceil_div(g.S, 128);
  g.T = (2 * win - 1) * (2 * win - 1);
  g.hdiv = FastDiv(hd);
  g.tmem_cols = 32;
  while ((int)g.tmem_cols < (g.SP > g.HDP ? g.SP : g.HDP)) g.tmem_cols <<= 1;
  g.q_bytes = (size_t)g.MT * 128 * g.HDP * 2;
  g.kv_bytes = (size_t)g.SP * g.HDP * 2;
  g.p_bytes = (size_t)128 * g.SP * 2;
  g.smem = g.q_bytes + 2 * g.kv_bytes + g.p_bytes + (size_t)round_up(g.T, 4) * 4 +
           (size_t)g.MT * 128 * 4 + (size_t)g.SP * 4;
  return g;
}

__global__ void __launch_bounds__(128) attn_core_tc_kernel(CoreParams p, CoreTcGeom g) {
  using namespace umma;
  extern __shared__ __align__(128) unsigned char core_tc_sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int win = p.win, S = g.S, SP = g.SP, HDP = g.HDP, MT = g.MT, hd = p.C / p.heads;
  unsigned char* Qt = core_tc_sm;                       // MT x [HDP/8][128][8]
  unsigned char* Kt = Qt + g.q_bytes;                   // [HDP/8][SP][8]
  unsigned char* Vt = Kt + g.kv_bytes;                  // same bytes, described MN-major
  unsigned char* Pt = Vt + g.kv_bytes;                  // [SP/8][128][8]
  float* rpb = reinterpret_cast<float*>(Pt + g.p_bytes);
  int* tok = reinterpret_cast<int*>(rpb + round_up(g.T, 4));       // [MT * 128]
  int* joff = tok + MT * 128;                                      // [SP]
  const int tid = threadIdx.x, warp = warp_idx_uniform();
  const int h = blockIdx.y, wdx = blockIdx.x;
  const int nWh = ceil_div(p.H, win), nWw = ceil_div(p.W, win);
  const int pad_h = nWh * win - p.H, pad_w = nWw * win - p.W;
  const bool use_mask = p.pad_mask && pad_h > 0 && pad_w > 0;
  const int b = wdx / (nWh * nWw), wy = (wdx / nWw) % nWh, wx = wdx % nWw;
  // tok: token index, -1 = zero-padded slot of the window, -2 = row beyond the window
  for (int s = tid; s < MT * 128; s += 128) {
    int t = -2;
    if (s < S) {
      const int y = wy * win + s / win - pad_h / 2, x = wx * win + s % win - pad_w / 2;
      t = (y >= 0 && y < p.H && x >= 0 && x < p.W) ? (b * p.H + y) * p.W + x : -1;
    }
    tok[s] = t;
  }
  for (int e = tid; e < g.T; e += 128) rpb[e] = __ldg(p.rpb + (size_t)h * g.T + e);
  // joff[j]: offset of key slot j in a row of the bias table; keys beyond S: 0 (masked below)
  for (int j = tid; j < SP; j += 128) joff[j] = j < S ? (j / win) * (2 * win - 1) + j % win : 0;
  if (warp == 0) tmem_alloc(&tmem_base_s, g.tmem_cols);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  // operand tiles: zero fill (row / column padding), then every real element by a thread of
  // its own, head_dim fastest: a warp reads runs of hd consecutive floats (a per-row gather
  // costs one L1 wavefront per thread and load instead)
  for (size_t off = (size_t)tid * 16; off < g.q_bytes + 2 * g.kv_bytes; off += 128 * 16)
    *reinterpret_cast<uint4*>(Qt + off) = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  {
    const int cb = h * hd;
    // four elements per trip: their 12 global loads are issued before the first store
    for (int e0 = tid; e0 < S * hd; e0 += 4 * 128) {
      int sl[4], d[4], t[4];
      float gq[4], gk[4], gv[4], pk[4], pv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u * 128 < S * hd ? e0 + u * 128 : e0;
        g.hdiv.divmod(e, sl[u], d[u]);
        t[u] = tok[sl[u]];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const size_t off = (size_t)(t[u] >= 0 ? t[u] : 0) * p.C + cb + d[u];
        gq[u] = __ldg(p.q + off);
        gk[u] = __ldg(p.k + off);
        gv[u] = __ldg(p.v + off);
        pk[u] = __ldg(p.bk + cb + d[u]);
        pv[u] = __ldg(p.bv + cb + d[u]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (e0 + u * 128 < S * hd) {
          const bool real = t[u] >= 0;
          *reinterpret_cast<__nv_bfloat16*>(Qt + (size_t)(sl[u] >> 7) * 128 * HDP * 2 +
                                            tile_off(sl[u] & 127, d[u], 128)) = __float2bfloat16_rn(real ? gq[u] : 0.f);
          *reinterpret_cast<__nv_bfloat16*>(Kt + tile_off(sl[u], d[u], SP)) = __float2bfloat16_rn(real ? gk[u] : pk[u]);
          *reinterpret_cast<__nv_bfloat16*>(Vt + tile_off(sl[u], d[u], SP)) = __float2bfloat16_rn(real ? gv[u] : pv[u]);
        }
      }
    }
  }
  const uint32_t q_addr = smem_u32(Qt), k_addr = smem_u32(Kt), v_addr = smem_u32(Vt), p_addr = smem_u32(Pt);
  const uint32_t idesc_s = idesc_bf16(128, SP, false, false), idesc_o = idesc_bf16(128, HDP, false, true);
  const int tw = 2 * win - 1;
  uint32_t phase = 0;
  for (int mt = 0; mt < MT; ++mt) {
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    const uint32_t tmem = tmem_base_s;
    if (warp == 0 && elect_one()) {
      tc_fence_after();
      for (int ks = 0; ks < HDP / 16; ++ks)
        mma_bf16(tmem, desc_kmajor(q_addr + (uint32_t)mt * 128u * (uint32_t)HDP * 2u, 128, ks),
                 desc_kmajor(k_addr, SP, ks), idesc_s, ks > 0);
      mma_commit(&bar);
    }
    cta_wait(&bar, phase);
    phase ^= 1;
    tc_fence_after();
    const int i = mt * 128 + tid, ic = i < S ? i : S - 1;
    const int ti = tok[i];
    const int ih = ic / win, iw = ic - ih * win;
    const float* rrow = rpb + (ih + win - 1) * tw + (iw + win - 1);      // - joff[j]
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    constexpr float kLog2e = 1.4426950408889634f;
    // pass 1: row maximum of (score + bias) in the log2 domain; branch-free so that the
    // shared-memory lookups of the 16 columns overlap
    float mx = -INFINITY;
    for (int c0 = 0; c0 < SP; c0 += 32) {
      float v[32];
      const bool two = c0 + 16 < SP;                      // SP is a multiple of 16, not of 32
      tmem_ld16(lane_addr + c0, v);
      if (two) tmem_ld16(lane_addr + c0 + 16, v + 16);
      tmem_ld_wait();
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) {
        const int j = (jj < 16 || two) ? c0 + jj : c0;
        float s = (v[jj < 16 || two ? jj : 0] + rrow[-joff[j]]) * kLog2e;
        const bool dead = j >= S || (use_mask && tok[j] < 0);
        s = dead ? -INFINITY : s;
        mx = fmaxf(mx, s);
      }
    }
    // pass 2: p = 2^(s - max) (0 for dead keys), row sum, bf16 P tile
    float sum = 0.f;
    for (int c0 = 0; c0 < SP; c0 += 32) {
      float v[32];
      const bool two = c0 + 16 < SP;
      tmem_ld16(lane_addr + c0, v);
      if (two) tmem_ld16(lane_addr + c0 + 16, v + 16);
      tmem_ld_wait();
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) {
        const bool live = jj < 16 || two;
        const int j = live ? c0 + jj : c0;
        float s = fmaf(v[live ? jj : 0] + rrow[-joff[j]], kLog2e, -mx);
        const bool dead = !live || j >= S || (use_mask && tok[j] < 0);
        s = dead ? -INFINITY : s;
        float e;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(s));
        v[jj] = e;
        sum += e;
      }
      st_chunk(Pt, tid, c0 / 8, 128, v);
      st_chunk(Pt, tid, c0 / 8 + 1, 128, v + 8);
      if (two) {
        st_chunk(Pt, tid, c0 / 8 + 2, 128, v + 16);
        st_chunk(Pt, tid, c0 / 8 + 3, 128, v + 24);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();                       // every row of S has been read: O may overwrite it
    if (warp == 0 && elect_one()) {
      tc_fence_after();
      for (int ks = 0; ks < SP / 16; ++ks)
        mma_bf16(tmem, desc_kmajor(p_addr, 128, ks), desc_mnmajor(v_addr, SP, ks), idesc_o, ks > 0);
      mma_commit(&bar);
    }
    cta_wait(&bar, phase);
    phase ^= 1;
    tc_fence_after();
    const float inv = 1.0f / sum;
    for (int c0 = 0; c0 < HDP; c0 += 16) {
      float v[16];
      tmem_ld16(lane_addr + c0, v);
      tmem_ld_wait();
      if (ti >= 0) {
#pragma unroll
        for (int jj = 0; jj < 16; ++jj)
          if (c0 + jj < hd) p.o[(size_t)ti * p.KO + h * p.hdp + c0 + jj] = v[jj] * inv;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, g.tmem_cols);
}

template <typename T>
static int launch_window_attn_generic(const AttnParams& p, cudaStream_t st) {
  const AttnLayout L(p.C, p.heads, p.win);
  const int n = p.B * p.H * p.W, C = p.C;
  HRF_REQUIRE(p.ws != nullptr, HRF_EINVAL, "attn (generic path): workspace required");
  float* ws = static_cast<float*>(p.ws);
  float* xn = ws; ws += (size_t)n * C;
  float* zn = xn;
  if (p.cross) { zn = ws; ws += (size_t)n * C; }
  float* q = ws; ws += (size_t)n * C;
  float* k = ws; ws += (size_t)n * C;
  float* v = ws; ws += (size_t)n * C;
  float* o = ws;
  const float* blob = p.blob;
  const int lnb = ceil_div(n * 32, 256);
  ln_rows_kernel<T><<<lnb, 256, 0, st>>>(static_cast<const T*>(p.xq), blob + L.o_lnq_w,
                                         blob + L.o_lnq_b, p.eps, xn, n, C);
  count_launch();
  if (p.cross) {
    ln_rows_kernel<T><<<lnb, 256, 0, st>>>(static_cast<const T*>(p.z), blob + L.o_lnkv_w,
                                           blob + L.o_lnkv_b, p.eps, zn, n, C);
    count_launch();
  }
  HRF_CUDA(cudaGetLastError());
  int rc;
  constexpr bool tc = std::is_same<T, __nv_bfloat16>::value;   // bf16 mode: tensor-core GEMMs
  GemmParams g{xn, blob + L.o_wq, blob + L.o_bq, nullptr, nullptr, q, n, C, C, C, C, 0};
  if ((rc = launch_gemm<float>(g, st, tc))) return rc;
  g.A = zn; g.Wt = blob + L.o_wk; g.bias = blob + L.o_bk; g.out = k;
  if ((rc = launch_gemm<float>(g, st, tc))) return rc;
  g.Wt = blob + L.o_wv; g.bias = blob + L.o_bv; g.out = v;
  if ((rc = launch_gemm<float>(g, st, tc))) return rc;
  HRF_CUDA(cudaMemsetAsync(o, 0, (size_t)n * L.KO * sizeof(float), st));   // head-pad lanes
  CoreParams c{q, k, v, o, blob + L.o_bk, blob + L.o_bv, blob + L.o_rpb,
               p.B, p.H, p.W, C, p.heads, p.win, L.KO, L.hdp, p.pad_mask};
  dim3 grid(p.B * ceil_div(p.H, p.win) * ceil_div(p.W, p.win), p.heads);
  const CoreTcGeom tg = core_tc_geom(p.win, L.hd);
  if (tc && !tc_disabled() && tg.smem <= 200 * 1024 && tg.HDP <= 256) {
    HRF_CUDA(ensure_smem((const void*)attn_core_tc_kernel, tg.smem));
    attn_core_tc_kernel<<<grid, 128, tg.smem, st>>>(c, tg);
  } else {
    const int S = p.win * p.win, ld = L.hd | 1;
    const size_t smem = ((size_t)3 * S * ld + 4 * S + S) * sizeof(float);
    HRF_REQUIRE(smem <= 227 * 1024, HRF_EUNSUPPORTED, "attn core: window %d x head_dim %d", p.win, L.hd);
    HRF_CUDA(ensure_smem((const void*)attn_core_kernel, smem));
    attn_core_kernel<<<grid, 128, smem, st>>>(c);
  }
  count_launch();
  HRF_CUDA(cudaGetLastError());
  // out = resid (+ z) + o Wo^T + bo
  GemmParams go{o, blob + L.o_wo, blob + L.o_bo, p.resid, p.cross ? p.z : nullptr, p.out,
                n, C, L.KO, L.KO, C, 0};
  return launch_gemm<T>(go, st, tc);
}

static size_t ffn_generic_ws_floats(int B, int H, int W, int C, int hidden) {
  const size_t n = (size_t)B * H * W;
  return n * C + 2 * n * hidden;
}

template <typename T>
static int launch_mixffn_generic(const FfnParams& p, cudaStream_t st) {
  const FfnLayout L(p.C, p.hidden);
  const int n = p.B * p.H * p.W, C = p.C, Hd = p.hidden;
  HRF_REQUIRE(p.ws != nullptr, HRF_EINVAL, "mixffn (generic path): workspace required");
  float* xn = static_cast<float*>(p.ws);
  float* h1 = xn + (size_t)n * C;
  float* h2 = h1 + (size_t)n * Hd;
  const float* blob = p.blob;
  ln_rows_kernel<T><<<ceil_div(n * 32, 256), 256, 0, st>>>(static_cast<const T*>(p.x), blob + L.o_ln_w,
                                                           blob + L.o_ln_b, p.eps, xn, n, C);
  count_launch();
  HRF_CUDA(cudaGetLastError());
  int rc;
  constexpr bool tc = std::is_same<T, __nv_bfloat16>::value;   // bf16 mode: tensor-core GEMMs
  GemmParams g1{xn, blob + L.o_w1, blob + L.o_b1, nullptr, nullptr, h1, n, Hd, C, C, Hd, 2};
  if ((rc = launch_gemm<float>(g1, st, tc))) return rc;
  HRF_REQUIRE(Hd % 4 == 0 && (size_t)n * (Hd / 4) < (1ull << 31), HRF_EUNSUPPORTED, "mixffn (generic path): hidden=%d", Hd);
  const size_t total = (size_t)n * (Hd / 4);
  const int dgrid = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
  if (tc) dw3x3_gelu_kernel<true><<<dgrid, 256, 0, st>>>(h1, blob + L.o_wd, blob + L.o_bd, h2, p.B, p.H, p.W, Hd);
  else dw3x3_gelu_kernel<false><<<dgrid, 256, 0, st>>>(h1, blob + L.o_wd, blob + L.o_bd, h2, p.B, p.H, p.W, Hd);
  count_launch();
  HRF_CUDA(cudaGetLastError());
  // out = x + GELU(h2 W2^T + b2)
  GemmParams g2{h2, blob + L.o_w2, blob + L.o_b2, p.x, nullptr, p.out, n, C, Hd, Hd, C, 2};
  return launch_gemm<T>(g2, st, tc);
}

}  // namespace hrf
