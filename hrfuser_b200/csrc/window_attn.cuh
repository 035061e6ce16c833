// Fused window attention (LSA / MWCA), fp32-math SIMT kernel.
//
// One CTA owns one window at a time:
//   LN prologue (pad slots -> zero rows, reference pads after the norm)
//   -> q/k/v projections (block_gemm, weights streamed k-major through L1/L2)
//   -> per (query row, head) softmax(QK^T + rpb) V by one warp
//   -> output projection with the residual adds fused into its epilogue.
// Pad / partition / merge / crop (reference hrformer.py:200-209,229-236) are
// pure address arithmetic: slot s of window (wy, wx) is token
// (wy*win + s/win - pad_t, wx*win + s%win - pad_l).
#pragma once
#include "common.cuh"

namespace hrf {

struct AttnLayout {
  int C, heads, hd, hdp, Cp, KO, S, T;  // T = (2*win-1)^2 table rows
  int o_lnq_w, o_lnq_b, o_lnkv_w, o_lnkv_b;
  int o_wq, o_bq, o_wk, o_bk, o_wv, o_bv, o_wo, o_bo, o_rpb;
  // tensor-core (bf16) sections: head-padded fp32 biases, then bf16 operand
  // tiles in the chunk-major layout of umma.cuh (offsets in floats)
  int tc_HDP, tc_KC, tc_NQ, tc_NOUT;
  int o_tc_bias, o_tc_wq, o_tc_wk, o_tc_wv, o_tc_wo;
  // third-generation kernel (C = 18, one head, window 7: window_attn_v3.cuh): self section
  // (4784 bytes) then cross section (5808 bytes); -1 when the shape is not covered
  int o_v3;
  int total;
  // shared-memory strides
  int ldx, ldq;
  __host__ __device__ AttnLayout(int C_, int heads_, int win) {
    C = C_; heads = heads_; hd = C / heads; hdp = round_up(hd, 4);
    Cp = round_up(C, 4); KO = heads * hdp; S = win * win; T = (2 * win - 1) * (2 * win - 1);
    const int c4 = round_up(C, 4);
    int o = 0;
    o_lnq_w = o; o += c4; o_lnq_b = o; o += c4; o_lnkv_w = o; o += c4; o_lnkv_b = o; o += c4;
    o_wq = o; o += Cp * C; o = round_up(o, 4); o_bq = o; o += c4;
    o_wk = o; o += Cp * C; o = round_up(o, 4); o_bk = o; o += c4;
    o_wv = o; o += Cp * C; o = round_up(o, 4); o_bv = o; o += c4;
    o_wo = o; o += KO * C; o = round_up(o, 4); o_bo = o; o += c4;
    o_rpb = o; o += round_up(heads * T, 4);
    tc_HDP = round_up(hd, 16); tc_KC = round_up(C, 16); tc_NQ = heads * tc_HDP; tc_NOUT = tc_KC;
    o_tc_bias = o; o += 3 * tc_NQ + tc_NOUT;                 // bq|bk|bv (head padded), bo
    o_tc_wq = o; o += tc_NQ * tc_KC / 2;                     // bf16 [KC/8][NQ][8]
    o_tc_wk = o; o += tc_NQ * tc_KC / 2;
    o_tc_wv = o; o += tc_NQ * tc_KC / 2;
    o_tc_wo = o; o += tc_NOUT * tc_NQ / 2;                   // bf16 [KO/8][NOUT][8], KO == NQ
    o_v3 = -1;
    if (C == 18 && heads == 1 && win == 7) {
      o = round_up(o, 4);
      o_v3 = o; o += (4784 + 5808) / 4;
    }
    total = o;
    ldx = stride4odd(Cp);
    ldq = stride4odd(KO);
  }
  __host__ __device__ size_t smem_floats(bool cross, int nwarps) const {
    size_t f = (size_t)S * ldx * (cross ? 2 : 1) + (size_t)3 * S * ldq + (size_t)heads * T +
               (size_t)nwarps * round_up(S, 32);
    return f + S /*token index*/ + 4;
  }
};

struct AttnParams {
  const void* xq;     // query source tokens
  const void* resid;  // residual source (== xq, or == out when accumulating)
  const void* z;      // key/value source tokens (== xq for self-attention)
  const float* blob;
  void* out;
  void* ws;           // fp32 workspace (split-head tensor-core variants) or nullptr
  int B, H, W, C, heads, win;
  int cross, pad_mask;
  float eps;
  FastDiv d_win_img, d_win_row;   // tensor-core kernel: window index -> (image, window row, col)
};

constexpr int kAttnThreads = 256;

template <typename T, int CT>
__global__ void __launch_bounds__(kAttnThreads) window_attn_kernel(AttnParams p) {
  extern __shared__ __align__(16) float smem[];
  const AttnLayout L(p.C, p.heads, p.win);
  const int S = L.S, C = L.C, win = p.win;
  const int nwarps = blockDim.x / 32, warp = threadIdx.x / 32, lane = threadIdx.x & 31;
  const int Spad = round_up(S, 32);

  float* XN = smem;
  float* ZN = p.cross ? XN + (size_t)S * L.ldx : XN;
  float* Qs = ZN + (size_t)S * L.ldx;
  float* Ks = Qs + (size_t)S * L.ldq;
  float* Vs = Ks + (size_t)S * L.ldq;
  float* RPB = Vs + (size_t)S * L.ldq;
  float* Pw = RPB + (size_t)L.heads * L.T;
  int* tok = reinterpret_cast<int*>(Pw + (size_t)nwarps * Spad);

  const int nWh = ceil_div(p.H, win), nWw = ceil_div(p.W, win);
  const int pad_h = nWh * win - p.H, pad_w = nWw * win - p.W;
  const int pad_t = pad_h / 2, pad_l = pad_w / 2;
  const bool use_mask = p.pad_mask && pad_h > 0 && pad_w > 0;
  const int n_windows = p.B * nWh * nWw;
  const float* blob = p.blob;
  const T* xq = static_cast<const T*>(p.xq);
  const T* zz = static_cast<const T*>(p.z);
  const T* rs = static_cast<const T*>(p.resid);
  T* out = static_cast<T*>(p.out);

  // one-time: rpb table (already (heads, T) in the blob), zero the q/k/v pad lanes
  for (int i = threadIdx.x; i < L.heads * L.T; i += blockDim.x) RPB[i] = __ldg(blob + L.o_rpb + i);
  for (int i = threadIdx.x; i < 3 * S * L.ldq; i += blockDim.x) Qs[i] = 0.f;

  for (int wdx = blockIdx.x; wdx < n_windows; wdx += gridDim.x) {
    const int b = wdx / (nWh * nWw);
    const int wy = (wdx / nWw) % nWh, wx = wdx % nWw;
    __syncthreads();  // previous window fully consumed (and one-time init visible)

    // ---- LN prologue: one warp per slot ------------------------------------
    for (int s = warp; s < S; s += nwarps) {
      const int h = wy * win + s / win - pad_t, w = wx * win + s % win - pad_l;
      const bool valid = h >= 0 && h < p.H && w >= 0 && w < p.W;
      const int t = valid ? (b * p.H + h) * p.W + w : -1;
      if (lane == 0) tok[s] = t;
      if (valid) {
        const T* src = xq + (size_t)t * C;
        warp_layernorm([&](int c) { return Elem<T>::ld(src + c); }, C, L.ldx,
                       blob + L.o_lnq_w, blob + L.o_lnq_b, p.eps, XN + (size_t)s * L.ldx);
        if (p.cross) {
          const T* zsrc = zz + (size_t)t * C;
          warp_layernorm([&](int c) { return Elem<T>::ld(zsrc + c); }, C, L.ldx,
                         blob + L.o_lnkv_w, blob + L.o_lnkv_b, p.eps, ZN + (size_t)s * L.ldx);
        }
      } else {
        for (int c = lane; c < L.ldx; c += 32) {
          XN[(size_t)s * L.ldx + c] = 0.f;
          if (p.cross) ZN[(size_t)s * L.ldx + c] = 0.f;
        }
      }
    }
    __syncthreads();

    // ---- projections -------------------------------------------------------
    const int hd = L.hd, hdp = L.hdp, ldq = L.ldq;
    auto scatter = [&](float* dst, const float* bias) {
      return [=](int r, int n, float v) {
        const int hh = n / hd;
        dst[(size_t)r * ldq + hh * hdp + (n - hh * hd)] = v + __ldg(bias + n);
      };
    };
    block_gemm<7, CT>(XN, L.ldx, S, blob + L.o_wq, L.Cp, C, scatter(Qs, blob + L.o_bq));
    block_gemm<7, CT>(ZN, L.ldx, S, blob + L.o_wk, L.Cp, C, scatter(Ks, blob + L.o_bk));
    block_gemm<7, CT>(ZN, L.ldx, S, blob + L.o_wv, L.Cp, C, scatter(Vs, blob + L.o_bv));
    __syncthreads();

    // ---- attention core: one warp per (head, query row) ----------------------
    float* P = Pw + (size_t)warp * Spad;
    for (int pi = warp; pi < L.heads * S; pi += nwarps) {
      const int hh = pi / S, i = pi - hh * S;
      const int ih = i / win, iw = i - ih * win;
      const float* q = Qs + (size_t)i * ldq + hh * hdp;
      const float* rp = RPB + (size_t)hh * L.T;
      float mx = -INFINITY;
      for (int j = lane; j < S; j += 32) {
        const float* k = Ks + (size_t)j * ldq + hh * hdp;
        float s = 0.f;
        for (int d = 0; d < hdp; d += 4) {
          const float4 a = *reinterpret_cast<const float4*>(q + d);
          const float4 bq = *reinterpret_cast<const float4*>(k + d);
          s = fmaf(a.x, bq.x, s); s = fmaf(a.y, bq.y, s);
          s = fmaf(a.z, bq.z, s); s = fmaf(a.w, bq.w, s);
        }
        const int jh = j / win, jw = j - jh * win;
        s += rp[(ih - jh + win - 1) * (2 * win - 1) + (iw - jw + win - 1)];
        if (use_mask && tok[j] < 0) s = -INFINITY;
        P[j] = s;
        mx = fmaxf(mx, s);
      }
      mx = warp_max(mx);
      float sum = 0.f;
      for (int j = lane; j < S; j += 32) {
        const float e = expf(P[j] - mx);
        P[j] = e;
        sum += e;
      }
      const float inv = 1.0f / warp_sum(sum);
      __syncwarp();
      for (int d = lane; d < hdp; d += 32) {
        const float* v = Vs + hh * hdp + d;
        float o = 0.f;
        for (int j = 0; j < S; ++j) o = fmaf(P[j], v[(size_t)j * ldq], o);
        // O overwrites this row's q slice: no other warp reads q(i, hh)
        Qs[(size_t)i * ldq + hh * hdp + d] = o * inv;
      }
      __syncwarp();
    }
    __syncthreads();

    // ---- output projection + residual(s) ----------------------------------
    const float* bo = blob + L.o_bo;
    const bool cross = p.cross;
    block_gemm<7, CT>(Qs, ldq, S, blob + L.o_wo, L.KO, C, [&](int r, int n, float v) {
      const int t = tok[r];
      if (t < 0) return;
      const size_t off = (size_t)t * C + n;
      float y = v + __ldg(bo + n) + Elem<T>::ld(rs + off);
      if (cross) y += Elem<T>::ld(zz + off);
      Elem<T>::st(out + off, y);
    });
    // the pad lanes of Qs were overwritten with O's (zero) pad lanes: still zero.
  }
}

template <typename T>
static int launch_window_attn(const AttnParams& p, cudaStream_t stream) {
  const AttnLayout L(p.C, p.heads, p.win);
  const size_t smem = L.smem_floats(p.cross, kAttnThreads / 32) * sizeof(float);
  const int n_windows = p.B * ceil_div(p.H, p.win) * ceil_div(p.W, p.win);
  const int grid = n_windows < 148 * 16 ? n_windows : 148 * 16;
  auto kern = (p.C % 4 == 0) ? window_attn_kernel<T, 4> : window_attn_kernel<T, 2>;
  HRF_CUDA(ensure_smem((const void*)kern, smem));
  kern<<<grid, kAttnThreads, smem, stream>>>(p);
  count_launch();
  HRF_CUDA(cudaGetLastError());
  return HRF_OK;
}

}  // namespace hrf
