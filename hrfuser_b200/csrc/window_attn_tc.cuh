// Fused window attention (LSA / MWCA) on the tcgen05 tensor cores -- bf16 mode.
//
// One CTA (128 threads) owns a PAIR of windows per iteration: thread r is row r of
// every M=128 tile, rows 0-63 = window A slots (49 real + 15 dead), rows 64-127 =
// window B.  Per pair:
//
//   LN prologue       per-thread LayerNorm of its own token (no shuffles), written
//                     as bf16 into the XN (and ZN) operand tile; pad slots -> zeros; column C
//                     of every row holds a constant 1 (row C of the weight tiles = q/k/v bias)
//   QKV  (UMMA)       [128 x KC] x Wq/Wk/Wv^T  -> TMEM  (N = heads*32 each)
//   epilogue          bf16 -> Q / K / V operand tiles per head (head_dim 18 -> 32, 39 -> 48)
//   per head:
//     S  (UMMA)       Q_h [128x32] x [K_A;K_B]^T -> TMEM 128 cols; each row reads only
//                     the 64 columns of its own window
//     softmax         in registers: + rel-pos bias (table lookups with immediate
//                     offsets), max, ex2, sum; P (unnormalised, bf16) -> 128x64 tile
//     PV (UMMA)       P x V_A and P x V_B (MN-major V, K = 64 keys) -> 2x32 TMEM cols;
//                     each row keeps the product with its own window's V
//     epilogue        x 1/sum, bf16 -> O operand tile
//   out-proj (UMMA)   O [128 x heads*32] x Wo^T -> TMEM
//   epilogue          + bias + residual (+ kv residual), bf16 -> global
//
// Pad / partition / merge / crop are address arithmetic (same closed forms as
// window_attn.cuh).  Accumulation, LayerNorm statistics and softmax are fp32.
// Persistent over window pairs (one resident wave of CTAs); the weight tiles arrive by TMA
// bulk copies behind the first tile's prologue; the next tile's rows are prefetched.
#pragma once
#include <cstdlib>

#include "common.cuh"
#include "umma.cuh"
#include "window_attn.cuh"

namespace hrf {

// HG = heads handled by one CTA.  HG == HEADS: the CTA finishes the block (bias,
// residuals, bf16 store).  HG < HEADS (wide, low-resolution branches: few windows,
// many heads): HEADS/HG CTAs share a window pair, each writes its fp32 partial
// out-projection to a workspace and `attn_reduce_kernel` sums them in fixed order.
template <int C, int HEADS, int HG, bool CROSS>
struct AttnTc {
  static constexpr int HD = C / HEADS;
  static constexpr int HDP = (HD + 15) / 16 * 16;
  static constexpr int KC = (C + 15) / 16 * 16;
  static constexpr int NQ = HEADS * HDP;       // all heads (blob tiles)
  static constexpr int NQG = HG * HDP;         // this CTA's heads
  static constexpr int NG = HEADS / HG;
  static constexpr int NOUT = KC;
  static constexpr int WIN = 7, S = 49;
  static constexpr bool SPLIT = HG < HEADS;
  static constexpr bool BIGC = C > 40;         // LayerNorm streams the row instead of holding it
  // q / k / v biases through the MMA: column C of the LN(x) / LN(z) tiles holds a constant 1
  // (pad and dead rows too: a zero-padded slot projects to the bias, as in the reference) and
  // row C of the weight tiles the bias.  Needs a spare K column (not C = 144).
  static constexpr bool BIAS_MMA = (C + 15) / 16 * 16 > C;
  static_assert(HEADS % HG == 0, "head groups");
  static_assert(HDP == 32 || HDP == 48, "tensor-core attention kernel: head_dim <= 48 (18 | 39 shipped)");
  static_assert(NQG <= 256 && NOUT <= 256, "one UMMA per projection");
  static constexpr int XT = 128 * (KC > NQG ? KC : NQG) * 2;   // XN tile, later aliased by the O tile
  static constexpr int WQ_B = NQG * KC * 2, WO_B = NOUT * NQG * 2;
  static constexpr int HT = 128 * HDP * 2;                    // one head's Q / K / V tile
  // shared-memory map (bytes)
  static constexpr int o_wq = 0, o_wk = o_wq + WQ_B, o_wv = o_wk + WQ_B, o_wo = o_wv + WQ_B;
  static constexpr int o_xn = o_wo + WO_B;
  static constexpr int o_zn = o_xn + XT;
  // per head [Q_h | K_h] (2 x HT); the 128 x 64 bf16 P tile of head h reuses exactly
  // these 16 KB once S = Q_h K_h^T has completed
  static constexpr int o_qk = o_zn + (CROSS ? 128 * KC * 2 : 0);
  static constexpr int o_v = o_qk + HG * 2 * HT;
  static constexpr int o_bias = o_v + HG * HT;                // fp32: bq|bk|bv (this group) | bo
  static_assert(2 * HT >= 128 * 64 * 2, "P tile must fit the Q_h|K_h pair");
  static constexpr int o_rpb = o_bias + (3 * NQG + NOUT) * 4; // fp32 [HG][169]
  static constexpr int o_ln = o_rpb + ((HG * 169 + 3) / 4 * 4) * 4;   // fp32 4 x C4
  static constexpr int C4 = (C + 3) / 4 * 4;
  static constexpr int SMEM = o_ln + 4 * C4 * 4;
  static constexpr int PROJ_COLS = 3 * NQG > NOUT ? 3 * NQG : NOUT;
  static constexpr int TMEM_COLS = (PROJ_COLS <= 128) ? 128 : (PROJ_COLS <= 256 ? 256 : 512);
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
  // CTAs that are resident per SM (shared memory incl. ~1 KB static + 1 KB reserved, TMEM
  // columns): the persistent grid is sized to exactly one wave of them
  static constexpr int BY_SMEM = (227 * 1024) / (SMEM + 2048), BY_TMEM = 512 / TMEM_COLS;
  static constexpr int CTAS_PER_SM = BY_SMEM < 1 ? 1 : (BY_SMEM < BY_TMEM ? BY_SMEM : BY_TMEM);
};

// one token row: C bf16 -> fp32 registers, widest aligned vector loads
template <int C>
__device__ __forceinline__ void load_row_bf16(const __nv_bfloat16* src, float* x) {
  if constexpr ((C * 2) % 16 == 0) {
#pragma unroll
    for (int i = 0; i < C / 8; ++i) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(src) + i);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
        x[i * 8 + j * 2] = f.x; x[i * 8 + j * 2 + 1] = f.y;
      }
    }
  } else if constexpr ((C * 2) % 8 == 0) {
#pragma unroll
    for (int i = 0; i < C / 4; ++i) {
      const uint2 u = __ldg(reinterpret_cast<const uint2*>(src) + i);
      const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
      const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
      x[i * 4] = a.x; x[i * 4 + 1] = a.y; x[i * 4 + 2] = b.x; x[i * 4 + 3] = b.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < C / 2; ++i) {
      const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(src) + i);
      const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
      x[i * 2] = a.x; x[i * 2 + 1] = a.y;
    }
  }
}

// raw variant: the row as C/2 packed bf16 pairs (kept in registers across a tile)
template <int C>
__device__ __forceinline__ void load_row_raw(const __nv_bfloat16* src, uint32_t* w) {
  if constexpr ((C * 2) % 16 == 0) {
#pragma unroll
    for (int i = 0; i < C / 8; ++i) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(src) + i);
      w[4 * i] = u.x; w[4 * i + 1] = u.y; w[4 * i + 2] = u.z; w[4 * i + 3] = u.w;
    }
  } else if constexpr ((C * 2) % 8 == 0) {
#pragma unroll
    for (int i = 0; i < C / 4; ++i) {
      const uint2 u = __ldg(reinterpret_cast<const uint2*>(src) + i);
      w[2 * i] = u.x; w[2 * i + 1] = u.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < C / 2; ++i) w[i] = __ldg(reinterpret_cast<const uint32_t*>(src) + i);
  }
}
template <int C>
__device__ __forceinline__ void unpack_row(const uint32_t* w, float* x) {
#pragma unroll
  for (int i = 0; i < C / 2; ++i) {
    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
    x[2 * i] = f.x; x[2 * i + 1] = f.y;
  }
}

template <int C>
__device__ __forceinline__ void store_row_bf16(__nv_bfloat16* dst, const float* x) {
  uint32_t w[C / 2];
#pragma unroll
  for (int i = 0; i < C / 2; ++i) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
    w[i] = *reinterpret_cast<const uint32_t*>(&h);
  }
  if constexpr ((C * 2) % 16 == 0) {
#pragma unroll
    for (int i = 0; i < C / 8; ++i)
      reinterpret_cast<uint4*>(dst)[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
  } else if constexpr ((C * 2) % 8 == 0) {
#pragma unroll
    for (int i = 0; i < C / 4; ++i) reinterpret_cast<uint2*>(dst)[i] = make_uint2(w[2 * i], w[2 * i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < C / 2; ++i) reinterpret_cast<uint32_t*>(dst)[i] = w[i];
  }
}

// LayerNorm of one row held in registers -> bf16 chunks of an R=128 operand tile.
// ONE: column C of the tile is set to 1 (the bias row of the weight tile multiplies it).
template <int C, int KC, bool ONE = false>
__device__ __forceinline__ void ln_row_to_tile(const float* x, const float* gamma,
                                               const float* beta, float eps, unsigned char* tile,
                                               int row) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) s += x[c];
  const float mean = s * (1.0f / C);
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) { const float d = x[c] - mean; q = fmaf(d, d, q); }
  const float rstd = rsqrtf(q * (1.0f / C) + eps);
#pragma unroll
  for (int ch = 0; ch < KC / 8; ++ch) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = ch * 8 + j;
      v[j] = (c < C) ? fmaf((x[c < C ? c : 0] - mean) * rstd, gamma[c < C ? c : 0], beta[c < C ? c : 0])
                     : (ONE && c == C) ? 1.f : 0.f;
    }
    umma::st_chunk(tile, row, ch, 128, v);
  }
}

// ---- LayerNorm of one token by a PAIR of adjacent lanes ------------------------------------
// half = lane & 1.  Half 0 owns channels [0, CA) (whole 8-channel chunks), half 1 owns [CA, C)
// plus the constant-1 column; the statistics are combined with two shuffles.  Halves the
// serial chain of the row-per-thread version and keeps every thread of a 2 x rows CTA busy.
template <int C>
struct LnPair {
  static constexpr int CA = (C / 2) / 8 * 8;                 // 8 (C=18) | 16 (C=36)
  static constexpr int NA = CA, NB = C - CA;                 // channels of half 0 / half 1
  static constexpr int NM = NB > NA ? NB : NA;
  static constexpr int NW = NM / 2;                          // packed words per thread
  static_assert(C % 2 == 0 && CA >= 8, "pair LayerNorm: even C >= 16");
};
// raw packed words of this thread's part of the row (absent words read as 0)
template <int C>
__device__ __forceinline__ void ln_pair_load(const __nv_bfloat16* row, int half, uint32_t* w) {
  using P = LnPair<C>;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(row) + (half ? P::CA / 2 : 0);
  const int n = half ? P::NB / 2 : P::NA / 2;
#pragma unroll
  for (int j = 0; j < P::NW; ++j) w[j] = (j < n) ? __ldg(src + j) : 0u;
}
// `valid` is uniform over the pair; an invalid (outside / pad) token becomes an all-zero row.
// Contains warp shuffles: every lane of the warp must call it (`store` = false for lanes whose
// row does not exist).
template <int C, int KC, bool ONE>
__device__ __forceinline__ void ln_pair_to_tile(const uint32_t* w, int half, bool valid, bool store,
                                                const float* gamma, const float* beta, float eps,
                                                unsigned char* tile, int row) {
  using P = LnPair<C>;
  constexpr int NM = P::NM, NA = P::NA, NB = P::NB;
  float v[NM];
#pragma unroll
  for (int j = 0; j < P::NW; ++j) {
    const uint32_t wj = valid ? w[j] : 0u;                    // invalid rows hold no data
    v[2 * j] = __uint_as_float(wj << 16);
    v[2 * j + 1] = __uint_as_float(wj & 0xffff0000u);
  }
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int i = 0; i < NM; i += 2) { s0 += v[i]; s1 += v[i + 1]; }
  float s = s0 + s1;
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  const float mean = s * (1.0f / C);
  float q0 = 0.f, q1 = 0.f;
#pragma unroll
  for (int i = 0; i < NM; i += 2) {
    float d0 = v[i] - mean, d1 = v[i + 1] - mean;
    if (i >= NA) { d0 = half ? d0 : 0.f; d1 = half ? d1 : 0.f; }    // half 0 has no such channel
    q0 = fmaf(d0, d0, q0);
    q1 = fmaf(d1, d1, q1);
    v[i] = d0; v[i + 1] = d1;
  }
  float q = q0 + q1;
  q += __shfl_xor_sync(0xffffffffu, q, 1);
  const float rstd = valid ? rsqrtf(q * (1.0f / C) + eps) : 0.f;
  const float one = valid ? 1.f : 0.f;
  const float* ga = gamma + (half ? P::CA : 0);
  const float* be = beta + (half ? P::CA : 0);
  constexpr int ZC0 = (C + (ONE ? 1 : 0) + 7) / 8;            // first all-zero chunk
  constexpr int KA = P::CA / 8 + (KC / 8 - ZC0), KB = ZC0 - P::CA / 8;
  constexpr int KMAX = KA > KB ? KA : KB;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = 8 * k + j;
      const float n = (i < NM) ? fmaf(v[i < NM ? i : 0] * rstd, ga[i < NM ? i : 0], valid ? be[i < NM ? i : 0] : 0.f) : 0.f;
      const float a = (i < NA) ? n : 0.f;
      const float b = (i < NB) ? n : (ONE && i == NB) ? one : 0.f;
      o[j] = (i < NA && i < NB) ? n : (half ? b : a);
    }
    const int chunk = half ? P::CA / 8 + k : (k < P::CA / 8 ? k : ZC0 + k - P::CA / 8);
    if (store && chunk < (half ? ZC0 : KC / 8)) umma::st_chunk(tile, row, chunk, 128, o);
  }
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// LayerNorm of one token row straight from global memory (three cached passes) for
// wide C, where holding the row in registers would spill
template <int C, int KC, bool ONE = false>
__device__ __forceinline__ void ln_row_streamed(const __nv_bfloat16* src, const float* gamma,
                                                const float* beta, float eps, unsigned char* tile,
                                                int row) {
  static_assert(C % 8 == 0, "streamed LN needs 16-byte rows");
  float s = 0.f;
#pragma unroll
  for (int ch = 0; ch < C / 8; ++ch) {
    float v[8];
    load_row_bf16<8>(src + ch * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[j];
  }
  const float mean = s * (1.0f / C);
  float q = 0.f;
#pragma unroll
  for (int ch = 0; ch < C / 8; ++ch) {
    float v[8];
    load_row_bf16<8>(src + ch * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float d = v[j] - mean; q = fmaf(d, d, q); }
  }
  const float rstd = rsqrtf(q * (1.0f / C) + eps);
#pragma unroll
  for (int ch = 0; ch < KC / 8; ++ch) {
    float v[8];
    if (ch < C / 8) {
      load_row_bf16<8>(src + ch * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaf((v[j] - mean) * rstd, gamma[ch * 8 + j], beta[ch * 8 + j]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (ONE && ch * 8 + j == C) ? 1.f : 0.f;
    }
    umma::st_chunk(tile, row, ch, 128, v);
  }
}

// the same for rows that are only 4-byte aligned (C = 78, 156: HRFuser-B): 32-bit loads
template <int C, int KC, bool ONE = false>
__device__ __forceinline__ void ln_row_streamed_w(const __nv_bfloat16* src, const float* gamma,
                                                  const float* beta, float eps, unsigned char* tile,
                                                  int row) {
  static_assert(C % 2 == 0, "even channel count");
  const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
  auto word = [&](int i) {
    const uint32_t u = __ldg(s32 + i);
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
  };
  float s0 = 0.f, s1 = 0.f;
#pragma unroll 13
  for (int i = 0; i < C / 2; ++i) { const float2 f = word(i); s0 += f.x; s1 += f.y; }
  const float mean = (s0 + s1) * (1.0f / C);
  float q0 = 0.f, q1 = 0.f;
#pragma unroll 13
  for (int i = 0; i < C / 2; ++i) {
    const float2 f = word(i);
    const float d0 = f.x - mean, d1 = f.y - mean;
    q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1);
  }
  const float rstd = rsqrtf((q0 + q1) * (1.0f / C) + eps);
#pragma unroll 1
  for (int ch = 0; ch < KC / 8; ++ch) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = ch * 8 + 2 * j;
      if (c < C) {
        const float2 f = word(c / 2);
        v[2 * j] = fmaf((f.x - mean) * rstd, gamma[c], beta[c]);
        v[2 * j + 1] = fmaf((f.y - mean) * rstd, gamma[c + 1], beta[c + 1]);
      } else {
        v[2 * j] = (ONE && c == C) ? 1.f : 0.f;
        v[2 * j + 1] = 0.f;
      }
    }
    umma::st_chunk(tile, row, ch, 128, v);
  }
}

template <int C, int KC, bool BIGC, bool ONE = false>
__device__ __forceinline__ void ln_token(const __nv_bfloat16* src, const float* gamma,
                                         const float* beta, float eps, unsigned char* tile, int row) {
  if constexpr (BIGC && C % 8 != 0) {
    ln_row_streamed_w<C, KC, ONE>(src, gamma, beta, eps, tile, row);
  } else if constexpr (BIGC) {
    ln_row_streamed<C, KC, ONE>(src, gamma, beta, eps, tile, row);
  } else {
    float x[C];
    load_row_bf16<C>(src, x);
    ln_row_to_tile<C, KC, ONE>(x, gamma, beta, eps, tile, row);
  }
}

template <int C, int HEADS, int HG, bool CROSS>
__global__ void __launch_bounds__(128) window_attn_tc_kernel(AttnParams p) {
  using namespace umma;
  using K = AttnTc<C, HEADS, HG, CROSS>;
  constexpr int HDP = K::HDP, KC = K::KC, NQ = K::NQ, NQG = K::NQG, NOUT = K::NOUT;
  constexpr int WIN = K::WIN, S = K::S, NG = K::NG;
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) uint64_t bar, wbar;
  __shared__ uint32_t tmem_base_s;
  __shared__ uint32_t valid_bits[4];

  HRF_PROF_DECL
  pdl_launch_dependents();
  const int tid = threadIdx.x, warp = warp_idx_uniform(), lane = tid & 31;
  const int g = tid >> 6, i = tid & 63;            // window of the pair, slot in the window
  const int hg = blockIdx.x % NG;                  // head group of this CTA
  const int h0 = hg * HG;                          // first (absolute) head
  const AttnLayout L(C, HEADS, WIN);
  const float* blob = p.blob;
  float* sBias = reinterpret_cast<float*>(sm + K::o_bias);
  float* sRpb = reinterpret_cast<float*>(sm + K::o_rpb);
  float* sLn = reinterpret_cast<float*>(sm + K::o_ln);

  // ---- one-time setup: this group's weight slices + small tables -> smem -----------
  // The bf16 weight tiles arrive by bulk async copies (one thread issues them, completion on
  // `wbar`) that run behind the table loads and the first tile's LN prologue.
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_init(&wbar, 1);
    fence_mbar_init();
    constexpr uint32_t piece = NQG * 16;           // one 8-column chunk of this group's rows
    constexpr uint32_t wo_b = (NQG / 8) * NOUT * 16;
    mbar_expect_tx(&wbar, 3 * (KC / 8) * piece + wo_b);
    // q/k/v tiles in the blob are [KC/8][NQ rows][16 B]; take rows h0*32 .. +NQG of each chunk
#pragma unroll 1
    for (int part = 0; part < 3; ++part) {
      const unsigned char* src = reinterpret_cast<const unsigned char*>(
          blob + (part == 0 ? L.o_tc_wq : part == 1 ? L.o_tc_wk : L.o_tc_wv));
      unsigned char* dst = sm + (part == 0 ? K::o_wq : part == 1 ? K::o_wk : K::o_wv);
#pragma unroll 1
      for (int ch = 0; ch < KC / 8; ++ch)
        bulk_g2s(dst + ch * piece, src + ((size_t)ch * NQ + h0 * HDP) * 16, piece, &wbar);
    }
    // out-proj tile [NQ/8][NOUT rows][16 B]: this group's K chunks are contiguous
    bulk_g2s(sm + K::o_wo, reinterpret_cast<const unsigned char*>(blob + L.o_tc_wo) + (size_t)(h0 * HDP / 8) * NOUT * 16,
             wo_b, &wbar);
  }
  {
    if constexpr (!K::BIAS_MMA) {
      for (int e = tid; e < 3 * NQG; e += 128)
        sBias[e] = __ldg(blob + L.o_tc_bias + (e / NQG) * NQ + h0 * HDP + (e % NQG));
    }
    for (int e = tid; e < NOUT; e += 128) sBias[3 * NQG + e] = __ldg(blob + L.o_tc_bias + 3 * NQ + e);
    for (int e = tid; e < HG * 169; e += 128) sRpb[e] = __ldg(blob + L.o_rpb + h0 * 169 + e);
    for (int e = tid; e < K::C4; e += 128) {
      sLn[e] = __ldg(blob + L.o_lnq_w + e);
      sLn[K::C4 + e] = __ldg(blob + L.o_lnq_b + e);
      sLn[2 * K::C4 + e] = __ldg(blob + L.o_lnkv_w + e);
      sLn[3 * K::C4 + e] = __ldg(blob + L.o_lnkv_b + e);
    }
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, K::TMEM_COLS);
  bool w_ready = false;                            // bulk copies observed complete
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);   // this thread's TMEM lane
  uint32_t phase = 0;

  const uint32_t a_wq = smem_u32(sm + K::o_wq), a_wk = smem_u32(sm + K::o_wk);
  const uint32_t a_wv = smem_u32(sm + K::o_wv), a_wo = smem_u32(sm + K::o_wo);
  const uint32_t a_xn = smem_u32(sm + K::o_xn), a_zn = smem_u32(sm + K::o_zn);
  const uint32_t a_qk = smem_u32(sm + K::o_qk), a_v = smem_u32(sm + K::o_v);

  const int nWh = ceil_div(p.H, WIN), nWw = ceil_div(p.W, WIN);
  const int pad_h = nWh * WIN - p.H, pad_w = nWw * WIN - p.W;
  const int pad_t = pad_h / 2, pad_l = pad_w / 2;
  const bool use_mask = p.pad_mask && pad_h > 0 && pad_w > 0;
  const int n_windows = p.B * nWh * nWw;
  const int n_tiles = (n_windows + 1) / 2;
  const size_t n_tok = (size_t)p.B * p.H * p.W;
  const __nv_bfloat16* xq = static_cast<const __nv_bfloat16*>(p.xq);
  const __nv_bfloat16* zz = static_cast<const __nv_bfloat16*>(p.z);
  const __nv_bfloat16* rs = static_cast<const __nv_bfloat16*>(p.resid);
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(p.out);

  // relative-position bias: value for (query slot i, key slot j) is
  // table[(ih-jh+6)*13 + (iw-jw+6)] = table[rp_base - (jh*13+jw)]
  const int ic = i < S ? i : S - 1;
  const int rp_base = (ic / WIN + WIN - 1) * (2 * WIN - 1) + (ic % WIN) + WIN - 1;

  // token held by this row in a given tile (-1: pad slot, dead row, no such window)
  auto row_token = [&](int tile) -> int {
    const int wdx = tile * 2 + g;
    if (tile >= n_tiles || wdx >= n_windows || i >= S) return -1;
    int b, rem, wy, wx;
    p.d_win_img.divmod(wdx, b, rem);
    p.d_win_row.divmod(rem, wy, wx);
    const int h = wy * WIN + i / WIN - pad_t, w = wx * WIN + i % WIN - pad_l;
    return (h >= 0 && h < p.H && w >= 0 && w < p.W) ? (b * p.H + h) * p.W + w : -1;
  };
  // Software pipeline (C <= 40): the raw bf16 rows of the NEXT tile are requested while
  // the current tile computes, and the current rows stay in registers for the residual
  // adds, so no thread ever sits on a global-load latency with the CTA waiting behind it.
  constexpr bool PIPE = !K::BIGC && !K::SPLIT;
  constexpr int NW = PIPE ? C / 2 : 1;
  uint32_t xr[NW], zr[NW];
  const int tile_step = gridDim.x / NG;
  pdl_wait();                                      // everything above only read the weight blob
  int tok = row_token(blockIdx.x / NG);
  if (PIPE && tok >= 0) {
    load_row_raw<C>(xq + (size_t)tok * C, xr);
    if (CROSS) load_row_raw<C>(zz + (size_t)tok * C, zr);
  }

  HRF_PROF(14)                                     // setup
  for (int tile = blockIdx.x / NG; tile < n_tiles; tile += tile_step) {
    HRF_PROF_TILE
    const int wdx = tile * 2 + g;
    {
      const unsigned bal = __ballot_sync(0xffffffffu, tok >= 0);
      if (lane == 0) valid_bits[warp] = bal;
    }
    // ---- LN prologue -------------------------------------------------------------
    if (tok >= 0) {
      if constexpr (PIPE) {
        float x[C];
        unpack_row<C>(xr, x);
        ln_row_to_tile<C, KC, K::BIAS_MMA>(x, sLn, sLn + K::C4, p.eps, sm + K::o_xn, tid);
        if (CROSS) {
          unpack_row<C>(zr, x);
          ln_row_to_tile<C, KC, K::BIAS_MMA>(x, sLn + 2 * K::C4, sLn + 3 * K::C4, p.eps, sm + K::o_zn, tid);
        }
      } else {
        ln_token<C, KC, K::BIGC, K::BIAS_MMA>(xq + (size_t)tok * C, sLn, sLn + K::C4, p.eps, sm + K::o_xn, tid);
        if (CROSS)
          ln_token<C, KC, K::BIGC, K::BIAS_MMA>(zz + (size_t)tok * C, sLn + 2 * K::C4, sLn + 3 * K::C4, p.eps,
                                   sm + K::o_zn, tid);
      }
    } else {
      const float zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ch = 0; ch < KC / 8; ++ch) {
        // zero row; with the bias column it still carries the constant 1
        const bool one = K::BIAS_MMA && ch == C / 8;
        float zc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) zc[j] = (one && j == C % 8) ? 1.f : zero[j];
        st_chunk(sm + K::o_xn, tid, ch, 128, zc);
        if (CROSS) st_chunk(sm + K::o_zn, tid, ch, 128, zc);
      }
    }
    // requests that complete behind the tile's MMAs / softmax: this tile's residual row
    // when it is not the query tensor (MWCA accumulation passes), the next tile's rows
    uint32_t rr[NW], xnext[NW], znext[NW];
    const int tok_next = row_token(tile + tile_step);
    if constexpr (PIPE) {
      if (tok >= 0 && rs != xq) load_row_raw<C>(rs + (size_t)tok * C, rr);
      if (tok_next >= 0) {
        load_row_raw<C>(xq + (size_t)tok_next * C, xnext);
        if (CROSS) load_row_raw<C>(zz + (size_t)tok_next * C, znext);
      }
    }
    HRF_PROF(0)                                    // LN prologue
    if (!w_ready) {                                // first tile: weights have landed?
      mbar_wait(&wbar, 0);
      w_ready = true;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    HRF_PROF(1)

    // ---- q / k / v projections of this head group --------------------------------------
    if (warp == 0 && elect_one()) {
      tc_fence_after();
      constexpr uint32_t idq = idesc_bf16(128, NQG, false, false);
      const uint32_t a_kv = CROSS ? a_zn : a_xn;
#pragma unroll
      for (int s = 0; s < KC / 16; ++s)
        mma_bf16(tmem, desc_kmajor(a_xn, 128, s), desc_kmajor(a_wq, NQG, s), idq, s > 0);
#pragma unroll
      for (int s = 0; s < KC / 16; ++s)
        mma_bf16(tmem + NQG, desc_kmajor(a_kv, 128, s), desc_kmajor(a_wk, NQG, s), idq, s > 0);
#pragma unroll
      for (int s = 0; s < KC / 16; ++s)
        mma_bf16(tmem + 2 * NQG, desc_kmajor(a_kv, 128, s), desc_kmajor(a_wv, NQG, s), idq, s > 0);
      mma_commit(&bar);
    }
    HRF_PROF(2)                                    // q/k/v issue
    cta_wait(&bar, phase);
    phase ^= 1;
    tc_fence_after();
    HRF_PROF(3)                                    // q/k/v wait
#pragma unroll
    for (int part = 0; part < 3; ++part) {
#pragma unroll
      for (int h = 0; h < HG; ++h) {
        unsigned char* dst = part == 0 ? sm + K::o_qk + h * 2 * K::HT
                           : part == 1 ? sm + K::o_qk + h * 2 * K::HT + K::HT
                                       : sm + K::o_v + h * K::HT;
        float v[HDP];
#pragma unroll
        for (int c0 = 0; c0 < HDP; c0 += 16) tmem_ld16(trow + part * NQG + h * HDP + c0, v + c0);
        tmem_ld_wait();
        if constexpr (!K::BIAS_MMA) {
          const float* bs = sBias + part * NQG + h * HDP;
#pragma unroll
          for (int c = 0; c < HDP; ++c) v[c] += bs[c];
        }
#pragma unroll
        for (int ch = 0; ch < HDP / 8; ++ch) st_chunk(dst, tid, ch, 128, v + 8 * ch);
      }
    }

    const unsigned long long vmask =
        (unsigned long long)valid_bits[2 * g] | ((unsigned long long)valid_bits[2 * g + 1] << 32);
    const bool mask_me = use_mask && wdx < n_windows;
    float inv_sum[HG];

#pragma unroll
    for (int h = 0; h < HG; ++h) {
      // ---- S = Q_h [K_A;K_B]^T -------------------------------------------------------
      HRF_PROF(4)                                  // q/k/v epilogue (h = 0) | PV epilogue
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      if (warp == 0 && elect_one()) {
        tc_fence_after();
        constexpr uint32_t ids = idesc_bf16(128, 128, false, false);
#pragma unroll
        for (int s = 0; s < HDP / 16; ++s)
          mma_bf16(tmem, desc_kmajor(a_qk + h * 2 * K::HT, 128, s),
                   desc_kmajor(a_qk + h * 2 * K::HT + K::HT, 128, s), ids, s > 0);
        mma_commit(&bar);
      }
      HRF_PROF(5)                                  // sync + S issue
      cta_wait(&bar, phase);
      phase ^= 1;
      tc_fence_after();
      HRF_PROF(6)                                  // S wait

      // ---- softmax over the 49 keys of this row's window ---------------------------
      float sc[56];
      const uint32_t scol = trow + g * 64;
      tmem_ld32(scol, sc);
      tmem_ld16(scol + 32, sc + 32);
      tmem_ld8(scol + 48, sc + 48);
      tmem_ld_wait();
      const float* tb = sRpb + h * 169 + rp_base;
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < S; ++j) {
        sc[j] += tb[-((j / WIN) * (2 * WIN - 1) + (j % WIN))];
        mx = fmaxf(mx, sc[j]);
      }
      if (use_mask) {             // uniform branch: no shipped config masks pad keys
        mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < S; ++j) {
          if (mask_me && !((vmask >> j) & 1ull)) sc[j] = -INFINITY;
          mx = fmaxf(mx, sc[j]);
        }
      }
      float sum = 0.f;
      const float mxl = mx * 1.4426950408889634f;
#pragma unroll
      for (int j = 0; j < S; ++j) {
        const float e = fast_exp2(fmaf(sc[j], 1.4426950408889634f, -mxl));
        sc[j] = e;
        sum += e;
      }
#pragma unroll
      for (int j = S; j < 56; ++j) sc[j] = 0.f;
      inv_sum[h] = 1.0f / sum;
      unsigned char* sP = sm + K::o_qk + h * 2 * K::HT;     // Q_h | K_h are dead: S is complete
#pragma unroll
      for (int ch = 0; ch < 7; ++ch) st_chunk(sP, tid, ch, 128, sc + 8 * ch);
      {
        const float zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        st_chunk(sP, tid, 7, 128, zero);
      }

      // ---- O = P V (both windows' V; each row keeps its own) -------------------------
      HRF_PROF(7)                                  // softmax
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      if (warp == 0 && elect_one()) {
        tc_fence_after();
        constexpr uint32_t ido = idesc_bf16(128, HDP, false, true);
#pragma unroll
        for (int g2 = 0; g2 < 2; ++g2)
#pragma unroll
          for (int s = 0; s < 4; ++s)
            mma_bf16(tmem + g2 * HDP, desc_kmajor(a_qk + h * 2 * K::HT, 128, s),
                     desc_mnmajor(a_v + h * K::HT + g2 * 64 * 16, 128, s), ido, s > 0);
        mma_commit(&bar);
      }
      HRF_PROF(8)                                  // sync + PV issue
      cta_wait(&bar, phase);
      phase ^= 1;
      tc_fence_after();
      HRF_PROF(9)                                  // PV wait
      {
        float o[HDP];
#pragma unroll
        for (int c0 = 0; c0 < HDP; c0 += 16) tmem_ld16(trow + g * HDP + c0, o + c0);
        tmem_ld_wait();
        const float is = inv_sum[h];
#pragma unroll
        for (int c = 0; c < HDP; ++c) o[c] *= is;
        // O tile aliases the XN tile (dead since the projections completed)
#pragma unroll
        for (int ch = 0; ch < HDP / 8; ++ch)
          st_chunk(sm + K::o_xn, tid, h * (HDP / 8) + ch, 128, o + 8 * ch);
      }
    }

    // ---- output projection (this group's K slice) --------------------------------------
    HRF_PROF(10)                                   // PV epilogue (last head)
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (warp == 0 && elect_one()) {
      tc_fence_after();
      constexpr uint32_t idy = idesc_bf16(128, NOUT, false, false);
#pragma unroll
      for (int s = 0; s < NQG / 16; ++s)
        mma_bf16(tmem, desc_kmajor(a_xn, 128, s), desc_kmajor(a_wo, NOUT, s), idy, s > 0);
      mma_commit(&bar);
    }
    HRF_PROF(11)                                   // sync + out-proj issue
    cta_wait(&bar, phase);
    phase ^= 1;
    tc_fence_after();
    HRF_PROF(12)                                   // out-proj wait
    if constexpr (!K::SPLIT) {
      float y[NOUT];
#pragma unroll
      for (int c0 = 0; c0 < NOUT; c0 += 16) tmem_ld16(trow + c0, y + c0);
      tmem_ld_wait();
      if (tok >= 0) {
        float r[C];
        if constexpr (PIPE) unpack_row<C>(rs != xq ? rr : xr, r);
        else load_row_bf16<C>(rs + (size_t)tok * C, r);
        const float* bo = sBias + 3 * NQG;
#pragma unroll
        for (int c = 0; c < C; ++c) y[c] += bo[c] + r[c];
        if (CROSS) {
          if constexpr (PIPE) unpack_row<C>(zr, r);
          else load_row_bf16<C>(zz + (size_t)tok * C, r);
#pragma unroll
          for (int c = 0; c < C; ++c) y[c] += r[c];
        }
        store_row_bf16<C>(out + (size_t)tok * C, y);
      }
    } else {
      // fp32 partial of this head group -> workspace [NG][n_tok][C]
      float* wrow = static_cast<float*>(p.ws) + ((size_t)hg * n_tok + (size_t)(tok >= 0 ? tok : 0)) * C;
#pragma unroll
      for (int c0 = 0; c0 < C; c0 += 8) {
        float y[8];
        tmem_ld8(trow + c0, y);
        tmem_ld_wait();
        if (tok >= 0) {
          if constexpr (C % 8 == 0) {
            *reinterpret_cast<float4*>(wrow + c0) = make_float4(y[0], y[1], y[2], y[3]);
            *reinterpret_cast<float4*>(wrow + c0 + 4) = make_float4(y[4], y[5], y[6], y[7]);
          } else {                     // C = 78 / 156: 8-byte aligned rows, partial last chunk
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (c0 + 2 * j < C) *reinterpret_cast<float2*>(wrow + c0 + 2 * j) = make_float2(y[2 * j], y[2 * j + 1]);
          }
        }
      }
    }
    HRF_PROF(13)                                   // output epilogue
    // rotate the software pipeline
    tok = tok_next;
    if constexpr (PIPE) {
#pragma unroll
      for (int j = 0; j < NW; ++j) { xr[j] = xnext[j]; zr[j] = znext[j]; }
    }
    // the next iteration's first barrier orders these TMEM reads before its MMAs
  }

  HRF_PROF_END
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, K::TMEM_COLS);
}

// out = resid (+ z) + bo + sum_g partial[g]   (fixed summation order: deterministic)
template <int C>
__global__ void __launch_bounds__(256) attn_reduce_kernel(const float* ws, int ng, size_t n_tok,
                                                          const __nv_bfloat16* rs,
                                                          const __nv_bfloat16* zz,
                                                          const float* __restrict__ bo,
                                                          __nv_bfloat16* out) {
  pdl_launch_dependents();
  pdl_wait();
  // flat over [n_tok][C] in vectors of 8 (a vector may straddle two tokens when C % 8 != 0)
  const size_t n_el = n_tok * C, n_vec = n_el / 8;
  if (blockIdx.x == 0 && threadIdx.x == 0) {         // tail when n_tok * C is not a multiple of 8
    for (size_t e = n_vec * 8; e < n_el; ++e) {
      float a = __bfloat162float(rs[e]) + (zz ? __bfloat162float(zz[e]) : 0.f) + bo[e % C];
      for (int gi = 0; gi < ng; ++gi) a += ws[(size_t)gi * n_el + e];
      out[e] = __float2bfloat16(a);
    }
  }
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_vec;
       v += (size_t)gridDim.x * blockDim.x) {
    const size_t e0 = v * 8;
    const int c0 = (int)(e0 % C);
    float acc[8], t[8];
    load_row_bf16<8>(rs + e0, acc);
    if (zz) {
      load_row_bf16<8>(zz + e0, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += t[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += __ldg(bo + (c0 + j < C ? c0 + j : c0 + j - C));
    for (int gi = 0; gi < ng; ++gi) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(ws + (size_t)gi * n_tok * C + e0));
      const float4 b = __ldg(reinterpret_cast<const float4*>(ws + (size_t)gi * n_tok * C + e0 + 4));
      acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
      acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
    }
    store_row_bf16<8>(out + e0, acc);
  }
}

// HRF_B78_SPLIT=1: C = 78 on the split variants (several CTAs per tile + fp32 workspace), A/B runs
static bool b78_split() {
  const char* e = std::getenv("HRF_B78_SPLIT");
  return e && e[0] == '1';
}

template <int C, int HEADS, int HG>
static int launch_attn_tc_ch(AttnParams p, cudaStream_t stream) {
  constexpr int NG = HEADS / HG;
  p.d_win_row = FastDiv(ceil_div(p.W, 7));
  p.d_win_img = FastDiv(ceil_div(p.H, 7) * ceil_div(p.W, 7));
  const int n_windows = p.B * ceil_div(p.H, 7) * ceil_div(p.W, 7);
  const int n_tiles = (n_windows + 1) / 2;
  // debug knob: HRF_ATTN_CTAS_PER_SM limits the persistent grid (occupancy experiments)
  static const int env_per_sm = [] { const char* e = std::getenv("HRF_ATTN_CTAS_PER_SM"); return e && atoi(e) > 0 ? atoi(e) : 0; }();
  const int fit = p.cross ? AttnTc<C, HEADS, HG, true>::CTAS_PER_SM : AttnTc<C, HEADS, HG, false>::CTAS_PER_SM;
  const int per_sm = env_per_sm > 0 ? env_per_sm : fit;
  const int cap = 148 * per_sm / NG > 0 ? 148 * per_sm / NG : 1;
  const int per_group = n_tiles < cap ? n_tiles : cap;      // one resident wave
  const int grid = per_group * NG;
  if (NG > 1) HRF_REQUIRE(p.ws != nullptr, HRF_EINVAL, "attn_tc: workspace required for C=%d", C);
  HRF_REQUIRE((reinterpret_cast<uintptr_t>(p.blob) & 15) == 0, HRF_EINVAL, "attn_tc: blob must be 16-byte aligned");
  if (p.cross) {
    using K = AttnTc<C, HEADS, HG, true>;
    HRF_CUDA(ensure_smem((const void*)window_attn_tc_kernel<C, HEADS, HG, true>, K::SMEM));
    HRF_CUDA(launch_pdl(window_attn_tc_kernel<C, HEADS, HG, true>, dim3(grid), dim3(128), K::SMEM, stream, p));
  } else {
    using K = AttnTc<C, HEADS, HG, false>;
    HRF_CUDA(ensure_smem((const void*)window_attn_tc_kernel<C, HEADS, HG, false>, K::SMEM));
    HRF_CUDA(launch_pdl(window_attn_tc_kernel<C, HEADS, HG, false>, dim3(grid), dim3(128), K::SMEM, stream, p));
  }
  count_launch();
  HRF_CUDA(cudaGetLastError());
  if constexpr (NG > 1) {
    const AttnLayout L(C, HEADS, 7);
    const size_t n_tok = (size_t)p.B * p.H * p.W;
    const size_t n_vec = n_tok * C / 8;
    const int rgrid = (int)((n_vec + 255) / 256 < 148 * 8 ? (n_vec + 255) / 256 : 148 * 8);
    HRF_CUDA(launch_pdl(attn_reduce_kernel<C>, dim3(rgrid), dim3(256), 0, stream,
                        static_cast<const float*>(p.ws), NG, n_tok, static_cast<const __nv_bfloat16*>(p.resid),
                        p.cross ? static_cast<const __nv_bfloat16*>(p.z) : nullptr,
                        p.blob + L.o_tc_bias + 3 * L.tc_NQ, static_cast<__nv_bfloat16*>(p.out)));
    count_launch();
    HRF_CUDA(cudaGetLastError());
  }
  return HRF_OK;
}

// true when the tensor-core kernel covers this problem
static bool attn_tc_supported(const AttnParams& p) {
  if (p.win != 7 || p.C % p.heads != 0) return false;
  const int hd = p.C / p.heads;
  if (hd == 18) return p.C == 18 || p.C == 36 || p.C == 72 || p.C == 144;     // HRFuser-T
  if (hd == 39) return p.C == 78 || p.C == 156;                                // HRFuser-B
  return false;
}
// fp32 workspace bytes of the split-head variants (0 when the CTA finishes the block itself)
static size_t attn_tc_workspace_bytes(int B, int H, int W, int C, int heads) {
  const int hd = heads > 0 ? C / heads : 0;
  const int ng = hd == 18 ? (C == 72 ? 2 : C == 144 ? 4 : 0) : hd == 39 ? (C == 78 ? 2 : C == 156 ? 4 : 0) : 0;
  return (size_t)ng * B * H * W * C * sizeof(float);
}

static int launch_window_attn_tc(const AttnParams& p, cudaStream_t stream) {
  switch (p.C) {
    case 18: return launch_attn_tc_ch<18, 1, 1>(p, stream);
    case 36: return launch_attn_tc_ch<36, 2, 2>(p, stream);
    case 72: return launch_attn_tc_ch<72, 4, 2>(p, stream);
    case 144: return launch_attn_tc_ch<144, 8, 2>(p, stream);
    case 78:                                                     // head_dim 39 -> padded to 48
      if (b78_split()) return launch_attn_tc_ch<78, 2, 1>(p, stream);
      return launch_attn_tc_ch<78, 2, 2>(p, stream);
    case 156: return launch_attn_tc_ch<156, 4, 1>(p, stream);
  }
  HRF_REQUIRE(false, HRF_EUNSUPPORTED, "attn_tc: C=%d", p.C);
}

}  // namespace hrf
