// Fused MixFFN on the tcgen05 tensor cores -- bf16 mode.
//
//   out = x + GELU(BN3(W2 * GELU(BN2(dw3x3(GELU(BN1(W1*LN(x)+b1)))+bd)) + b2))
//
// One CTA (256 threads) owns a 6 x 14 tile of output tokens per iteration:
//   LN prologue   the 8 x 16 halo (128 tokens) is LayerNormed (a lane pair per token for
//                 C <= 36) and written as one bf16 M=128 operand tile with a constant-1
//                 column that carries b1 through the MMA (all-zero rows outside the image)
//   per chunk of 72 (HRFuser-T) / 78 (HRFuser-B) hidden channels (hidden = 4C):
//     fc1 (UMMA)  [128 x KC] x W1c^T -> TMEM (N = 80)
//     epilogue    GELU (zero outside the image: that is what the reference's zero-padded
//                 depthwise conv sees), bf16 -> H1 [9 chunks][128 tok][8]
//     dw 3x3      CUDA cores on H1: strips of 3 outputs x 4 channels per thread, packed
//                 fp32x2 FMAs, + bd, GELU, bf16 -> H2 operand tile [88 x 80]
//     fc2 (UMMA)  H2 x W2c^T accumulated over the chunks in TMEM (N = NOUT); b2 rides on
//                 the constant-1 column 72 | 78 of H2
//   epilogue      GELU, + residual, bf16 -> global
// The 4C-wide hidden activation never leaves the SM.  BN is folded (eval mode).
#pragma once
#include <cstdlib>

#include "common.cuh"
#include "mixffn.cuh"
#include "umma.cuh"
#include "window_attn_tc.cuh"   // load_row_bf16 / store_row_bf16 / ln_row_to_tile

namespace hrf {

// CPG = hidden chunks handled by one CTA.  CPG == NCH: the CTA finishes the block.
// CPG < NCH (C = 72, 78, 144, 156: few tokens or many chunks): NCH/CPG CTAs share a token tile,
// each writes its fp32 partial fc2 product to a workspace and `ffn_reduce_kernel`
// applies GELU / residual to the fixed-order sum.
//
// Tile geometry: 6 x 14 outputs inside an 8 x 16 = 128 token halo = exactly one M=128 fc1
// tile, 256 threads.  Small tiles keep the per-tile dependency chain short and give the
// low-resolution branches (1 920 .. 30 720 tokens) enough CTAs to cover the 148 SMs; C = 18
// needs 128 TMEM columns and 47 KB of shared memory -> four CTAs (four independent barrier
// domains) per SM, the wider variants two.
template <int C, int CPG>
struct FfnTc {
  // hidden channels per chunk: 72 (HRFuser-T widths 18k) or 78 (HRFuser-B widths 78k); both
  // are padded to N1 = 80 MMA columns, the first padding column (72 | 78) carries the constant 1
  static constexpr int CH = (C % 18 == 0) ? 72 : 78;
  static constexpr int HID = 4 * C, NCH = HID / CH, NG = NCH / CPG;
  static constexpr bool SPLIT = CPG < NCH, BIGC = C > 40;
  static_assert(HID % CH == 0 && NCH % CPG == 0, "hidden must split into 72- or 78-channel chunks");
  static constexpr int NDC = (CH + 7) / 8;           // 8-channel data chunks per hidden chunk: 9 | 10
  static constexpr int ONE_CHUNK = CH / 8, ONE_ELEM = CH % 8;    // where the constant 1 lives in H2
  static constexpr int KC = (C + 15) / 16 * 16, NOUT = KC, N1 = 80;
  static constexpr int TH = 6, TW = 14;
  static constexpr int HH = TH + 2, HW = TW + 2, NHALO = HH * HW;      // 128
  static constexpr int NTOK = TH * TW;                                  // 84 outputs
  static constexpr int NMT = 1;                                         // fc1 M tiles
  static_assert(NHALO == 128, "the halo is one M=128 operand tile");
  static constexpr int NT = 256, NGQ = NT / 128;     // threads, groups per TMEM quadrant
  static constexpr int XT = 128 * KC * 2;            // the LN(x) operand tile
  static constexpr int H1R = NHALO;                  // rows per 16-byte column group of H1
  // fc1 bias through the MMA: XN column C holds 1 for in-image tokens and W1 row C holds b1, so
  // the epilogue needs neither a bias add nor the outside-the-image select (zero rows give
  // GELU(0) = 0).  Needs a spare K column (not C = 144).
  static constexpr bool BIAS_MMA = KC > C;
  static constexpr int H1_B = NDC * H1R * 16;        // GELU(fc1) on the halo, bf16
  // H2 = fc2 A operand: NTOK live rows; the M=128 MMA also reads (and ignores) the rows up to
  // 127, which alias the next chunk / the bytes behind the tile
  static constexpr int H2R = (NTOK + 7) / 8 * 8;     // 88
  // depthwise conv units: SH vertically adjacent outputs x 4 channels per thread
  static constexpr int SH = 3, NSTRIP = TH / SH;
  static_assert(TH % SH == 0, "tile height must split into strips");
  static constexpr int H2_B = 9 * H2R * 16 + 128 * 16;
  // With one chunk per CTA LN(x) is dead once fc1 has run: H1 | H2 reuse its bytes.
  static constexpr bool XN_ALIAS = (CPG == 1);
  static constexpr int ACT_B = XN_ALIAS ? (XT > H1_B + H2_B ? XT : H1_B + H2_B) : XT + H1_B + H2_B;
  // ... in which case a wide LN(x) tile also overwrites the constant-1 column of H2
  static constexpr bool REWRITE_ONE = XN_ALIAS && XT > H1_B + 9 * H2R * 16;
  // shared-memory map (bytes)
  static constexpr int o_w1 = 0;                               // CPG tiles [80 x KC]
  static constexpr int o_w2 = o_w1 + CPG * N1 * KC * 2;        // CPG tiles [NOUT x 80]
  static constexpr int o_xn = o_w2 + CPG * NOUT * N1 * 2;
  static constexpr int o_h1 = XN_ALIAS ? o_xn : o_xn + XT;
  static constexpr int o_h2 = o_h1 + H1_B;
  static constexpr int o_f32 = o_xn + ACT_B;                   // per chunk 880 floats, then b2[NOUT]
  static constexpr int o_ln = o_f32 + (CPG * 880 + NOUT) * 4;  // gamma[C4] beta[C4]
  static constexpr int C4 = (C + 3) / 4 * 4;
  static constexpr int o_in = o_ln + 2 * C4 * 4;               // inside flags [256] bytes
  static constexpr int SMEM = o_in + 256;
  static constexpr int D_COLS = NMT * N1;                      // fc1 accumulators
  static constexpr int Y_COL = D_COLS;                         // fc2 accumulator
  static constexpr int TMEM_COLS = (D_COLS + NOUT <= 128) ? 128 : (D_COLS + NOUT <= 256) ? 256 : 512;
  static constexpr int CTAS_PER_SM = (SMEM + 1024) * 4 <= 227 * 1024 && TMEM_COLS <= 128 ? 4
                                   : (SMEM + 1024) * 3 <= 227 * 1024 && TMEM_COLS <= 128 ? 3
                                   : (SMEM + 1024) * 2 <= 227 * 1024 && TMEM_COLS <= 256 ? 2 : 1;
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

// GELU for the bf16 mode: 0.5 x (1 + tanh(x (a + b x^2))), a / b the minimax fit to the exact
// erf form (max abs error 2.7e-4, tools/fit_gelu.py; the bf16 rounding of the value it
// produces is 4e-3 at the |x| ~ 3 where that maximum sits) on the hardware tanh unit
// (MUFU.TANH).  b > 0 keeps the argument monotone, so no clamp is needed, and the four
// multiply-adds are issued as packed fp32x2 instructions (FMUL2 / FFMA2, sm_100): 7
// instructions per PAIR of values instead of ~17 per value for an erf evaluation.
constexpr float kGeluA = 0.80015708f, kGeluB = 0.03470089f;
__device__ __forceinline__ float2 gelu_as2(float2 x) {
  const float2 x2 = __fmul2_rn(x, x);
  const float2 w = __ffma2_rn(x2, make_float2(kGeluB, kGeluB), make_float2(kGeluA, kGeluA));
  const float2 a = __fmul2_rn(w, x);
  float2 t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(a.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(a.y));
  const float2 hx = __fmul2_rn(x, make_float2(0.5f, 0.5f));
  return __ffma2_rn(hx, t, hx);
}
__device__ __forceinline__ float gelu_as(float x) {
  const float w = fmaf(x * x, kGeluB, kGeluA);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(w * x));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
// in place on 8 values
__device__ __forceinline__ void gelu8(float* v) {
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    const float2 g = gelu_as2(make_float2(v[j], v[j + 1]));
    v[j] = g.x;
    v[j + 1] = g.y;
  }
}

// 16-byte H2 chunk that holds only the constant 1 (bf16) at element `elem`
__device__ __forceinline__ uint4 one_chunk(int elem) {
  const uint32_t w = (elem & 1) ? 0x3F800000u : 0x00003F80u;
  const int i = elem >> 1;
  return make_uint4(i == 0 ? w : 0u, i == 1 ? w : 0u, i == 2 ? w : 0u, i == 3 ? w : 0u);
}

// Warp w reads TMEM lanes 32*(w%4).. (its quadrant q) and belongs to group gq = w/4.
// Epilogue and depthwise work is dealt out in units of (row, 8-channel chunk) so all groups of
// a quadrant stay busy and no thread holds more than 8 accumulators (<= 64 registers).
template <int C, int CPG>
__global__ void __launch_bounds__(FfnTc<C, CPG>::NT, FfnTc<C, CPG>::CTAS_PER_SM)
mixffn_tc_kernel(FfnParams p) {
  using namespace umma;
  using K = FfnTc<C, CPG>;
  constexpr int KC = K::KC, NOUT = K::NOUT, N1 = K::N1, NCH = K::NCH, NG = K::NG;
  constexpr int NT = K::NT, NGQ = K::NGQ, NMT = K::NMT;
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) uint64_t bar, wbar;
  __shared__ uint32_t tmem_base_s;

  HRF_PROF_DECL
  pdl_launch_dependents();
  const int tid = threadIdx.x, warp = warp_idx_uniform(), lane = tid & 31;
  const int gq = warp >> 2, q = warp & 3;          // work group, TMEM quadrant
  const int row = q * 32 + lane;                   // TMEM lane == row of the M=128 tiles
  const int cg = blockIdx.x % NG;                  // chunk group of this CTA
  const FfnLayout L(C, K::HID);
  const float* blob = p.blob;
  float* sF = reinterpret_cast<float*>(sm + K::o_f32);
  float* sLn = reinterpret_cast<float*>(sm + K::o_ln);
  unsigned char* sIn = sm + K::o_in;

  // ---- one-time setup ---------------------------------------------------------------
  // The weight tiles and per-chunk tables (this group's chunks are contiguous in each blob
  // section) arrive by bulk async copies that run behind the first tile's LN prologue.
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_init(&wbar, 1);
    fence_mbar_init();
    constexpr uint32_t b1 = CPG * N1 * KC * 2, b2 = CPG * NOUT * N1 * 2, b3 = CPG * 880 * 4, b4 = NOUT * 4;
    mbar_expect_tx(&wbar, b1 + b2 + b3 + b4);
    bulk_g2s(sm + K::o_w1, reinterpret_cast<const unsigned char*>(blob + L.o_tc_w1) + (size_t)cg * b1, b1, &wbar);
    bulk_g2s(sm + K::o_w2, reinterpret_cast<const unsigned char*>(blob + L.o_tc_w2) + (size_t)cg * b2, b2, &wbar);
    bulk_g2s(sF, blob + L.o_tc_f32 + (size_t)cg * CPG * 880, b3, &wbar);
    bulk_g2s(sF + CPG * 880, blob + L.o_tc_f32 + NCH * 880, b4, &wbar);
  }
  {
    for (int e = tid; e < K::C4; e += NT) {
      sLn[e] = __ldg(blob + L.o_ln_w + e);
      sLn[K::C4 + e] = __ldg(blob + L.o_ln_b + e);
    }
    // (every row and chunk of the LN(x) tile is rewritten per tile: no initialisation)
    // H2: zero; its 10th chunk (the K padding, channels 72..79) holds the constant 1 in
    // channel 72 that multiplies the b2 row of the W2 tile
    uint4* z = reinterpret_cast<uint4*>(sm + K::o_h2);
    for (int e = tid; e < K::H2_B / 16; e += NT)
      z[e] = (e >= 9 * K::H2R && e < 10 * K::H2R) ? one_chunk(K::ONE_ELEM) : make_uint4(0, 0, 0, 0);
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, K::TMEM_COLS);
  bool w_ready = false;                            // bulk copies observed complete
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
  uint32_t phase = 0;
  const uint32_t a_w1 = smem_u32(sm + K::o_w1), a_w2 = smem_u32(sm + K::o_w2);
  const uint32_t a_xn = smem_u32(sm + K::o_xn), a_h2 = smem_u32(sm + K::o_h2);

  const int tiles_x = ceil_div(p.W, K::TW), tiles_y = ceil_div(p.H, K::TH);
  const int n_tiles = p.B * tiles_x * tiles_y;
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(p.x);
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(p.out);
  // epilogue-1 units of this quadrant: (M tile mt, chunk 0..8) for every M tile that has live
  // halo rows in this quadrant
  const int n_units = K::NDC * ((K::NHALO - q * 32 + 127) / 128);

  // Software pipeline (C <= 40): the halo token of the NEXT tile is requested while the
  // current tile computes; no thread waits on global memory with the CTA behind it.  There a
  // PAIR of adjacent lanes normalises one halo token (ln_pair_to_tile); for wide C one thread
  // streams the row.
  constexpr bool PIPE = !K::BIGC;
  constexpr int NW = PIPE ? LnPair<PIPE ? C : 16>::NW : 1;
  const int tile_step = gridDim.x / NG;
  const int ht = PIPE ? tid >> 1 : tid, half = tid & 1;     // halo token of this thread
  // global token index of halo token `ht` in a given tile, or -1 (outside / none)
  auto halo_token = [&](int tile) -> int {
    if (tile >= n_tiles || ht >= K::NHALO) return -1;
    int b, rem, ty, tx;
    p.d_tiles_xy.divmod(tile, b, rem);
    p.d_tiles_x.divmod(rem, ty, tx);
    const int h = ty * K::TH - 1 + ht / K::HW;
    const int w = tx * K::TW - 1 + ht % K::HW;
    return (h >= 0 && h < p.H && w >= 0 && w < p.W) ? (b * p.H + h) * p.W + w : -1;
  };
  pdl_wait();                                      // everything above only read the weight blob
  uint32_t xr[NW];
  int htok = halo_token(blockIdx.x / NG);
  if constexpr (PIPE) {
    if (htok >= 0) ln_pair_load<C>(x + (size_t)htok * C, half, xr);
  }

  HRF_PROF(14)                                     // setup
  for (int tile = blockIdx.x / NG; tile < n_tiles; tile += tile_step) {
    HRF_PROF_TILE
    int b, rem, ty0, tx0;
    p.d_tiles_xy.divmod(tile, b, rem);
    p.d_tiles_x.divmod(rem, ty0, tx0);
    ty0 *= K::TH;
    tx0 *= K::TW;

    // ---- LN prologue ---------------------------------------------------------------------
    if constexpr (PIPE) {
      // whole warps (the pair LayerNorm shuffles): skip only warps with no halo token at all
      if (((tid & ~31) >> 1) < K::NHALO) {
        const bool row_ok = ht < K::NHALO, in = htok >= 0;
        unsigned char* xt = sm + K::o_xn + ((row_ok ? ht : 0) >> 7) * K::XT;
        ln_pair_to_tile<C, KC, K::BIAS_MMA>(xr, half, in, row_ok, sLn, sLn + K::C4, p.eps, xt, ht & 127);
        if (!K::BIAS_MMA && half == 0 && row_ok) sIn[ht] = in ? 1 : 0;
      }
    } else if (ht < K::NHALO) {
      const bool in = htok >= 0;
      unsigned char* xt = sm + K::o_xn + (ht >> 7) * K::XT;
      {
        sIn[ht] = in ? 1 : 0;
        if (in) {
          ln_token<C, KC, true, K::BIAS_MMA>(x + (size_t)htok * C, sLn, sLn + K::C4, p.eps, xt, ht & 127);
        } else {
          const float zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int ch = 0; ch < KC / 8; ++ch) st_chunk(xt, ht & 127, ch, 128, zero);
        }
      }
    }
    // requests that complete behind this tile's work
    const int htok_next = halo_token(tile + tile_step);
    uint32_t xnext[NW];
    if constexpr (PIPE) {
      if (htok_next >= 0) ln_pair_load<C>(x + (size_t)htok_next * C, half, xnext);
    }
    const int oh = ty0 + row / K::TW, ow = tx0 + row % K::TW;
    const bool o_in = row < K::NTOK && oh < p.H && ow < p.W;
    const size_t o_tok = o_in ? (size_t)(b * p.H + oh) * p.W + ow : 0;
    // residual slices of this thread's epilogue-2 units cc = gq + u * NGQ
    constexpr int NU2 = K::SPLIT ? 1 : ((C + 7) / 8 + NGQ - 1) / NGQ;
    uint32_t rres[NU2][4];
    if constexpr (!K::SPLIT) {
#pragma unroll
      for (int u = 0; u < NU2; ++u)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c0 = (gq + u * NGQ) * 8 + 2 * j;
          rres[u][j] = (o_in && c0 < C) ? __ldg(reinterpret_cast<const uint32_t*>(x + o_tok * C + c0)) : 0u;
        }
    }

#pragma unroll 1
    for (int c = 0; c < CPG; ++c) {
      // ---- fc1 on the halo M tile(s) ---------------------------------------------------
      HRF_PROF(0)                                  // LN prologue (first chunk)
      if (!w_ready) {                              // first tile: weights have landed?
        mbar_wait(&wbar, 0);
        w_ready = true;
      }
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      HRF_PROF(1)
      if (warp == 0 && elect_one()) {
        tc_fence_after();
        constexpr uint32_t id1 = idesc_bf16(128, N1, false, false);
        const uint32_t w1c = a_w1 + c * (N1 * KC * 2);
#pragma unroll
        for (int t = 0; t < NMT; ++t)
#pragma unroll
          for (int s = 0; s < KC / 16; ++s)
            mma_bf16(tmem + t * N1, desc_kmajor(a_xn + t * K::XT, 128, s), desc_kmajor(w1c, N1, s),
                     id1, s > 0);
        mma_commit(&bar);
      }
      HRF_PROF(2)                                  // fc1 issue
      cta_wait(&bar, phase);
      phase ^= 1;
      tc_fence_after();
      HRF_PROF(3)                                  // fc1 wait

      // ---- epilogue 1: (+ b1,) GELU, zero outside the image -> H1 -----------------------
      const float* fb = sF + c * 880;
      // two units per iteration: both TMEM loads are in flight before the first GELU chain
      auto epi1_unit = [&](int u, float* v) {
        const int mt = u >= K::NDC ? 1 : 0, ch = u - mt * K::NDC;
        const int t = mt * 128 + row;                 // halo token
        if (t < K::NHALO) {
          if constexpr (K::BIAS_MMA) {
            gelu8(v);
          } else {
            const bool in = sIn[t] != 0;
            const float4 ba = *reinterpret_cast<const float4*>(fb + ch * 8);
            const float4 bb = *reinterpret_cast<const float4*>(fb + ch * 8 + 4);
            const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += bv[j];
            gelu8(v);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = in ? v[j] : 0.f;
          }
          st_chunk(sm + K::o_h1, t, ch, K::H1R, v);
        }
      };
#pragma unroll 1
      for (int u = gq; u < n_units; u += 2 * NGQ) {   // warp-uniform trip count
        const int u2 = u + NGQ;
        const bool two = u2 < n_units;                // warp-uniform
        float va[8], vb[8];
        {
          const int mt = u >= K::NDC ? 1 : 0;
          tmem_ld8(trow + mt * N1 + (u - mt * K::NDC) * 8, va);
        }
        if (two) {
          const int mt = u2 >= K::NDC ? 1 : 0;
          tmem_ld8(trow + mt * N1 + (u2 - mt * K::NDC) * 8, vb);
        }
        tmem_ld_wait();
        epi1_unit(u, va);
        if (two) epi1_unit(u2, vb);
      }
      HRF_PROF(4)                                  // epilogue 1
      tc_fence_before();
      __syncthreads();
      HRF_PROF(5)

      // ---- depthwise 3x3 + GELU -> H2 --------------------------------------------------
      {
        // bf16 H1: unit = (strip of SH vertically adjacent outputs, 4 channels).  Every halo
        // value is loaded and unpacked once per column offset and feeds up to three outputs;
        // the multiply-adds are packed fp32x2 (FFMA2).  Lane pairs cover the two halves of one
        // token's 16-byte chunk, so a warp's 8-byte accesses are contiguous.
        constexpr int SH = K::SH;
        constexpr int NUNIT = K::NDC * K::NSTRIP * K::TW * 2;
        if (K::REWRITE_ONE && tid < K::H2R)          // LN(x) overwrote the constant-1 column
          *reinterpret_cast<uint4*>(sm + K::o_h2 + (9 * K::H2R + tid) * 16) = one_chunk(K::ONE_ELEM);
        const float* wd = fb + 80;
        const float* bd = fb + 800;
#pragma unroll 1
        for (int id = tid; id < NUNIT; id += NT) {
          const int hf = id & 1;
          int r = id >> 1;
          const int ox = r % K::TW;
          r /= K::TW;
          const int st = r % K::NSTRIP, ch = r / K::NSTRIP;
          const int oy0 = st * SH, c4 = ch * 8 + hf * 4;
          const float4 bq = *reinterpret_cast<const float4*>(bd + c4);
          float2 acc[SH][2];
#pragma unroll
          for (int o = 0; o < SH; ++o) {
            acc[o][0] = make_float2(bq.x, bq.y);
            acc[o][1] = make_float2(bq.z, bq.w);
          }
          const unsigned char* hp = sm + K::o_h1 + ((size_t)ch * K::H1R + oy0 * K::HW + ox) * 16 + hf * 8;
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            float2 f[SH + 2][2];
#pragma unroll
            for (int rr = 0; rr < SH + 2; ++rr) {
              const uint2 u = *reinterpret_cast<const uint2*>(hp + (rr * K::HW + dx) * 16);
              f[rr][0] = make_float2(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u));
              f[rr][1] = make_float2(__uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
            }
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              const float4 w = *reinterpret_cast<const float4*>(wd + (dy * 3 + dx) * 80 + c4);
              const float2 w0 = make_float2(w.x, w.y), w1 = make_float2(w.z, w.w);
#pragma unroll
              for (int o = 0; o < SH; ++o) {
                acc[o][0] = __ffma2_rn(f[o + dy][0], w0, acc[o][0]);
                acc[o][1] = __ffma2_rn(f[o + dy][1], w1, acc[o][1]);
              }
            }
          }
#pragma unroll
          for (int o = 0; o < SH; ++o) {
            float2 g0 = gelu_as2(acc[o][0]), g1 = gelu_as2(acc[o][1]);
            if constexpr (K::ONE_ELEM != 0) {      // CH = 78: this half-chunk holds the constant 1
              if (ch == K::ONE_CHUNK && hf == K::ONE_ELEM / 4) {
                if (K::ONE_ELEM % 4 == 0) g0.x = 1.f;
                if (K::ONE_ELEM % 4 == 1) g0.y = 1.f;
                if (K::ONE_ELEM % 4 == 2) g1.x = 1.f;
                if (K::ONE_ELEM % 4 == 3) g1.y = 1.f;
              }
            }
            const __nv_bfloat162 p0 = __floats2bfloat162_rn(g0.x, g0.y), p1 = __floats2bfloat162_rn(g1.x, g1.y);
            uint2 u;
            u.x = *reinterpret_cast<const uint32_t*>(&p0);
            u.y = *reinterpret_cast<const uint32_t*>(&p1);
            *reinterpret_cast<uint2*>(sm + K::o_h2 + ((size_t)ch * K::H2R + (oy0 + o) * K::TW + ox) * 16 + hf * 8) = u;
          }
        }
      }

      // ---- fc2 partial product over this chunk's 72 (padded 80) channels --------------
      HRF_PROF(6)                                  // depthwise conv
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      HRF_PROF(7)
      if (warp == 0 && elect_one()) {
        tc_fence_after();
        constexpr uint32_t id2 = idesc_bf16(128, NOUT, false, false);
        const uint32_t w2c = a_w2 + c * (NOUT * N1 * 2);
#pragma unroll
        for (int s = 0; s < N1 / 16; ++s)
          mma_bf16(tmem + K::Y_COL, desc_kmajor(a_h2, K::H2R, s), desc_kmajor(w2c, NOUT, s), id2,
                   (c > 0) || (s > 0));
        mma_commit(&bar);
      }
      HRF_PROF(8)                                  // fc2 issue
      cta_wait(&bar, phase);        // H2 / XN free again, Y complete after the last chunk
      phase ^= 1;
      tc_fence_after();
      HRF_PROF(9)                                  // fc2 wait
    }

    // ---- epilogue 2: unit = (output token, 8-channel chunk of the C outputs) ----------
    // (b2 came through the MMA: row 72 of the first chunk's W2 tile x the constant-1 column)
    {
      constexpr int NCC = (C + 7) / 8;
#pragma unroll
      for (int u = 0; u < (NCC + NGQ - 1) / NGQ; ++u) {
        const int cc = gq + u * NGQ;
        if (cc < NCC) {                               // warp-uniform
          float y[8];
          tmem_ld8(trow + K::Y_COL + cc * 8, y);
          tmem_ld_wait();
          if (o_in) {
            if constexpr (K::SPLIT) {  // fp32 partial of this chunk group -> workspace [NG][n_tok][C]
              const size_t n_tok = (size_t)p.B * p.H * p.W;
              float* wrow = static_cast<float*>(p.ws) + ((size_t)cg * n_tok + o_tok) * C + cc * 8;
              if constexpr (C % 8 == 0) {
                *reinterpret_cast<float4*>(wrow) = make_float4(y[0], y[1], y[2], y[3]);
                *reinterpret_cast<float4*>(wrow + 4) = make_float4(y[4], y[5], y[6], y[7]);
              } else {                 // C = 78 / 156: rows are 8-byte aligned, the last chunk is partial
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  if (cc * 8 + 2 * j < C) *reinterpret_cast<float2*>(wrow + 2 * j) = make_float2(y[2 * j], y[2 * j + 1]);
              }
            } else {                   // GELU, + residual; rows are only 4-byte aligned
              gelu8(y);
              uint32_t* orow = reinterpret_cast<uint32_t*>(out + o_tok * C + cc * 8);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (cc * 8 + 2 * j < C) {
                  const uint32_t rw = rres[K::SPLIT ? 0 : u][j];
                  const float2 r = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rw));
                  const __nv_bfloat162 hh = __floats2bfloat162_rn(r.x + y[2 * j], r.y + y[2 * j + 1]);
                  orow[j] = *reinterpret_cast<const uint32_t*>(&hh);
                }
              }
            }
          }
        }
      }
    }
    HRF_PROF(10)                                   // epilogue 2
    // rotate the software pipeline
    htok = htok_next;
    if constexpr (PIPE) {
#pragma unroll
      for (int j = 0; j < NW; ++j) xr[j] = xnext[j];
    }
    // next iteration: its first barrier (after the LN prologue) orders these TMEM reads
    // before the next fc1 / fc2 MMAs; the XN tiles were released by the fc2 wait above.
  }

  HRF_PROF_END
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, K::TMEM_COLS);
}

// out = x + GELU(b2 + sum_g partial[g])   (fixed summation order)
template <int C>
__global__ void __launch_bounds__(256) ffn_reduce_kernel(const float* ws, int ng, size_t n_tok,
                                                         const __nv_bfloat16* x,
                                                         const float* __restrict__ b2,
                                                         __nv_bfloat16* out) {
  pdl_launch_dependents();
  pdl_wait();
  // flat over [n_tok][C] (no per-channel term: b2 is part of the first partial, MMA bias row)
  const size_t n_el = n_tok * C, n_vec = n_el / 8;
  if (blockIdx.x == 0 && threadIdx.x == 0) {         // tail when n_tok * C is not a multiple of 8
    for (size_t e = n_vec * 8; e < n_el; ++e) {
      float a = 0.f;
      for (int gi = 0; gi < ng; ++gi) a += ws[(size_t)gi * n_el + e];
      out[e] = __float2bfloat16(__bfloat162float(x[e]) + gelu_as(a));
    }
  }
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_vec;
       v += (size_t)gridDim.x * blockDim.x) {
    const size_t e0 = v * 8;
    float acc[8], r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int gi = 0; gi < ng; ++gi) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(ws + (size_t)gi * n_tok * C + e0));
      const float4 b = __ldg(reinterpret_cast<const float4*>(ws + (size_t)gi * n_tok * C + e0 + 4));
      acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
      acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
    }
    load_row_bf16<8>(x + e0, r);
    gelu8(acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += r[j];
    store_row_bf16<8>(out + e0, acc);
  }
}

static bool ffn_tc_supported(const FfnParams& p) {
  return p.hidden == 4 * p.C && (p.C == 18 || p.C == 36 || p.C == 72 || p.C == 144 ||   // HRFuser-T
                                 p.C == 78 || p.C == 156);                               // HRFuser-B
}
// chunk groups (CTAs per tile) of the split variants; 0: the CTA finishes the block
static int ffn_tc_groups(int C) { return C == 72 || C == 78 ? 2 : C == 144 || C == 156 ? 8 : 0; }   // upper bound (workspace size)
static size_t ffn_tc_workspace_bytes(int B, int H, int W, int C, int hidden) {
  if (hidden != 4 * C) return 0;
  return (size_t)ffn_tc_groups(C) * B * H * W * C * sizeof(float);
}

template <int C, int CPG>
static int launch_ffn_tc_c(FfnParams p, cudaStream_t stream) {
  using K = FfnTc<C, CPG>;
  const int n_tiles = p.B * ceil_div(p.H, K::TH) * ceil_div(p.W, K::TW);
  p.d_tiles_x = FastDiv(ceil_div(p.W, K::TW));
  p.d_tiles_xy = FastDiv(ceil_div(p.H, K::TH) * ceil_div(p.W, K::TW));
  // debug knob: HRF_FFN_CTAS_PER_SM limits the persistent grid (occupancy experiments)
  static const int env_per_sm = [] { const char* e = std::getenv("HRF_FFN_CTAS_PER_SM"); return e ? atoi(e) : 0; }();
  const int per_sm = env_per_sm > 0 ? env_per_sm : K::CTAS_PER_SM;
  const int cap = 148 * per_sm / K::NG > 0 ? 148 * per_sm / K::NG : 1;
  const int grid = (n_tiles < cap ? n_tiles : cap) * K::NG;
  if (K::SPLIT) HRF_REQUIRE(p.ws != nullptr, HRF_EINVAL, "mixffn_tc: workspace required for C=%d", C);
  HRF_REQUIRE((reinterpret_cast<uintptr_t>(p.blob) & 15) == 0, HRF_EINVAL, "mixffn_tc: blob must be 16-byte aligned");
  HRF_CUDA(ensure_smem((const void*)mixffn_tc_kernel<C, CPG>, K::SMEM));
  HRF_CUDA(launch_pdl(mixffn_tc_kernel<C, CPG>, dim3(grid), dim3(K::NT), K::SMEM, stream, p));
  count_launch();
  HRF_CUDA(cudaGetLastError());
  if constexpr (K::SPLIT) {
    const FfnLayout L(C, K::HID);
    const size_t n_tok = (size_t)p.B * p.H * p.W;
    const size_t n_vec = n_tok * C / 8;
    const int rgrid = (int)((n_vec + 255) / 256 < 148 * 8 ? (n_vec + 255) / 256 : 148 * 8);
    HRF_CUDA(launch_pdl(ffn_reduce_kernel<C>, dim3(rgrid), dim3(256), 0, stream,
                        static_cast<const float*>(p.ws), K::NG, n_tok, static_cast<const __nv_bfloat16*>(p.x),
                        p.blob + L.o_tc_f32 + K::NCH * 880, static_cast<__nv_bfloat16*>(p.out)));
    count_launch();
    HRF_CUDA(cudaGetLastError());
  }
  return HRF_OK;
}

static int launch_mixffn_tc(const FfnParams& p, cudaStream_t stream) {
  switch (p.C) {
    case 18: return launch_ffn_tc_c<18, 1>(p, stream);
    case 36: return launch_ffn_tc_c<36, 2>(p, stream);
    case 72: return launch_ffn_tc_c<72, 2>(p, stream);
    case 144: return launch_ffn_tc_c<144, 1>(p, stream);
    case 78:
      // HRF_B78_SPLIT=1: the split variant (two CTAs per tile + fp32 workspace), for A/B runs
      if (b78_split()) return launch_ffn_tc_c<78, 2>(p, stream);
      return launch_ffn_tc_c<78, 4>(p, stream);
    case 156:
      if (b78_split()) return launch_ffn_tc_c<156, 1>(p, stream);
      return launch_ffn_tc_c<156, 2>(p, stream);
  }
  HRF_REQUIRE(false, HRF_EUNSUPPORTED, "mixffn_tc: C=%d", p.C);
}

}  // namespace hrf
