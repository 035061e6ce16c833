// Fused MixFFN, bf16 storage, second generation (C = 18 | 36: the high-resolution branches
// that carry 95 % of the tokens).
//
//   out = x + GELU(BN3(W2 * GELU(BN2(dw3x3(GELU(BN1(W1*LN(x)+b1)))+bd)) + b2))
//
// What changed against mixffn_tc.cuh (which stays for the split low-resolution widths):
//   * the halo tile of raw x arrives by ONE TMA tensor-tile copy (cp.async.bulk.tensor.3d):
//     the token tensor is described as (W*C/2 words, H, B), a box of 24 tokens (16-byte
//     aligned start: 4 tokens left of the tile) x (TH+2) rows lands in shared memory as dense rows and everything outside the image is
//     zero-filled by the TMA unit -- no per-thread 4-byte global loads, no index arithmetic,
//     no boundary predicates; the next tile's box is in flight while this tile computes
//     (two-stage ring); the residual is read back from the same shared-memory tile and the
//     output tile leaves by one TMA tensor store (clipped at the image edge by the unit);
//   * a 12 x 16 output tile inside a 14 x 18 halo (252 tokens = two M=128 fc1 tiles): the
//     fc1 + GELU recompute on the halo falls from 1.52x to 1.31x;
//   * the hidden activation is fp16 on chip (H1, H2, the depthwise and fc2 weights): the
//     depthwise 3x3 runs as packed HFMA2 straight on the 16-byte shared-memory chunks
//     (8 channels per 128-bit load, no unpacking, 3 vertically adjacent outputs share their
//     15 loads), its GELU as packed half2 around MUFU.TANH; fp16 keeps 11 significant bits
//     where bf16 kept 8, so the result is closer to the fp32 reference than before;
//   * LayerNorm's affine is folded into W1 / b1 and the 0.5 of every GELU into the weights
//     in front of it (hrf_ffn_pack), so LN is (x - mean) * rstd and GELU is
//     hx + hx * tanh(hx * (2a + 8b hx^2)): 6 instructions per PAIR of values.
// Per tile: LN (a thread per halo token) -> fc1 (tcgen05, both M tiles, one commit) ->
// epilogue 1 (TMEM -> GELU -> fp16 H1) -> depthwise + GELU -> H2 -> fc2 (tcgen05, fp16) ->
// epilogue 2 (GELU, + residual, bf16 -> output tile) -> TMA store.
// Reference: hrformer.py:267-295 (CrossFFN) with the LayerNorm / residual of hrformer.py:371
// and hrfuser_hrformer_based.py:315.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "mixffn.cuh"
#include "mixffn_tc.cuh"
#include "tmap.cuh"
#include "umma.cuh"

namespace hrf {

template <int C, int TH, int TW, int NT>
struct FfnV2 {
  static constexpr int CH = 72, HID = 4 * C, NCH = HID / CH;
  static_assert(C % 18 == 0 && HID % CH == 0, "HRFuser-T widths");
  static constexpr int NDC = 9;                       // live 8-channel chunks per hidden chunk
  static constexpr int KC = (C + 15) / 16 * 16, NOUT = KC, N1 = 80;
  static_assert(KC > C, "the fc1 bias rides on a spare K column");
  static constexpr int HH = TH + 2, HW = TW + 2, NHALO = HH * HW;
  static constexpr int NMT = (NHALO + 127) / 128;     // fc1 M tiles
  static constexpr int XR = NMT * 128;                // rows of the LN(x) operand tile
  // TMA boxes: the byte offset of the box start along the innermost dimension and the box
  // width in bytes must both be multiples of 16 (measured: tools/probes/tma_probe.cu -- an
  // unaligned start coordinate is an illegal-instruction fault), i.e. multiples of ALIGN_TOK
  // tokens.  The load box therefore starts PADL tokens left of the tile instead of 1.
  static constexpr int ALIGN_TOK = (C * 2) % 16 == 0 ? 1 : (C * 2) % 8 == 0 ? 2 : 4;
  static constexpr int PADL = ALIGN_TOK;
  static constexpr int BOXW = (PADL + TW + 1 + ALIGN_TOK - 1) / ALIGN_TOK * ALIGN_TOK;
  static_assert(TW % ALIGN_TOK == 0 && (C * 2) % 4 == 0, "TMA box alignment");
  static_assert(BOXW * C / 2 <= 256 && TW * C / 2 <= 256, "TMA box dimension limit");
  static constexpr int NTOK = TH * TW, NMT2 = (NTOK + 127) / 128;
  static constexpr int SH = 3, NSTRIP = TH / SH;
  static_assert(TH % SH == 0 && TW % 2 == 0, "tile must split into 3 x 2 depthwise units");
  static constexpr int NW = NT / 32;
  // bytes
  static constexpr int RAW_ROW = BOXW * C * 2;
  static constexpr int RAW_B = (HH * RAW_ROW + 127) / 128 * 128;
  static constexpr int XN_B = (KC / 8) * XR * 16;
  static constexpr int H1_B = NDC * NHALO * 16;
  static constexpr int H2R = NTOK;                                    // live rows of the fc2 A tile
  static constexpr int H2_B = 10 * H2R * 16 + (NMT2 * 128 - H2R) * 16;   // + the last M tile's over-read
  // one hidden chunk: LN(x) is dead once fc1 has completed, H2 reuses its bytes (never the
  // tenth chunk of H2, which holds the constant 1 that multiplies the b2 row of W2)
  static constexpr bool XN_ALIAS = NCH == 1 && XN_B <= 9 * H2R * 16;
  static constexpr int OUT_B = (TH * TW * C * 2 + 127) / 128 * 128;
  static constexpr int W1_B = NCH * N1 * KC * 2, W2_B = NCH * NOUT * N1 * 2, CV_B = NCH * 1600;
  // shared-memory map
  static constexpr int o_raw = 0;
  static constexpr int o_h2 = o_raw + 2 * RAW_B;
  static constexpr int o_xn = XN_ALIAS ? o_h2 : o_h2 + ((H2_B + 127) / 128 * 128);
  static constexpr int o_h1 = (XN_ALIAS ? o_h2 + ((H2_B + 127) / 128 * 128) : o_xn + XN_B);
  static constexpr int o_out = o_h1 + ((H1_B + 127) / 128 * 128);
  static constexpr int o_w1 = o_out + OUT_B;
  static constexpr int o_w2 = o_w1 + W1_B;
  static constexpr int o_cv = o_w2 + W2_B;
  static constexpr int SMEM = o_cv + CV_B;
  static constexpr int D_COLS = NMT * N1, Y_COL = D_COLS;
  static constexpr int NEED_COLS = D_COLS + NMT2 * NOUT;
  static constexpr int TMEM_COLS = NEED_COLS <= 128 ? 128 : NEED_COLS <= 256 ? 256 : 512;
  static_assert(NEED_COLS <= 512, "TMEM columns");
  static constexpr int BY_SMEM = (227 * 1024) / (SMEM + 1024 + 128);
  static constexpr int BY_TMEM = 512 / TMEM_COLS;
  static constexpr int BY_REGS = 65536 / (NT * 56);    // at least 56 registers per thread
  static constexpr int CTAS0 = BY_SMEM < BY_TMEM ? BY_SMEM : BY_TMEM;
  static constexpr int CTAS_PER_SM = CTAS0 < 1 ? 1 : (CTAS0 < BY_REGS ? CTAS0 : BY_REGS);
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
  static_assert(XR <= NT, "a thread per row of the LN(x) tile");
};

// GELU on values that arrive pre-halved (hx = x / 2): hx + hx * tanh(hx * (2a + 8b hx^2))
// with the fitted a / b of mixffn_tc.cuh.
constexpr float kG2A = 2.f * kGeluA, kG8B = 8.f * kGeluB;
__device__ __forceinline__ float2 gelu_hx2(float2 hx) {
  const float2 x2 = __fmul2_rn(hx, hx);
  const float2 w = __ffma2_rn(x2, make_float2(kG8B, kG8B), make_float2(kG2A, kG2A));
  const float2 a = __fmul2_rn(w, hx);
  float2 t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(a.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(a.y));
  return __ffma2_rn(hx, t, hx);
}
__device__ __forceinline__ uint32_t h2_as_u32(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ __half2 u32_as_h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ __half2 gelu_hx_h2(__half2 hx) {
  const __half2 k8b = __floats2half2_rn(kG8B, kG8B), k2a = __floats2half2_rn(kG2A, kG2A);
  const __half2 x2 = __hmul2(hx, hx);
  const __half2 w = __hfma2(x2, k8b, k2a);
  const uint32_t a = h2_as_u32(__hmul2(w, hx));
  uint32_t t;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(a));
  return __hfma2(hx, u32_as_h2(t), hx);
}
// two fp32 -> packed fp16 (x in the low half), saturating to the largest finite value
__device__ __forceinline__ uint32_t pack_f16x2(float x, float y) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(y), "f"(x));
  return r;
}

// Several PROBLEMS of one shape per launch (the camera's branch 0 and the modality streams run the
// same MixFFN on tensors of the same size, each with its own weights): the images of all
// problems form one tile space, every CTA walks ONE contiguous range of it (so it crosses from
// one problem into the next at most once or twice and reloads the 11.6 KB of weights there), and
// the launch, the setup and -- above all -- the partly filled last round are paid once:
// 3 x 640 tiles over 296 resident CTAs are 7 rounds instead of 3 x 3.
constexpr int kFfnMaxProb = 4;      // camera + up to three modality streams (T-stf)
struct FfnV2Group {
  const float* blob[kFfnMaxProb];
  int n_prob, tiles_per_cta;      // tiles_per_cta > 0: contiguous ranges; 0: tiles strided over the grid (one problem)
};
template <int NP>
struct FfnV2Maps {                 // NP = 1: the single-problem launch keeps its two maps (and its speed)
  CUtensorMap x[NP], o[NP];
};

#ifndef HRF_FFN_GELU1_HALF
#define HRF_FFN_GELU1_HALF 1
#endif
__device__ __forceinline__ constexpr bool v2_gelu1_half() { return HRF_FFN_GELU1_HALF != 0; }

template <int C, int TH, int TW, int NT, int NP>
__global__ void __launch_bounds__(NT, FfnV2<C, TH, TW, NT>::CTAS_PER_SM)
mixffn_v2_kernel(FfnParams p, FfnV2Group gp, const __grid_constant__ FfnV2Maps<NP> tm) {
  using namespace umma;
  using K = FfnV2<C, TH, TW, NT>;
  constexpr int KC = K::KC, NOUT = K::NOUT, N1 = K::N1, NCH = K::NCH, NW = K::NW;
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) uint64_t full[2], wbar, bar1, bar2;
  __shared__ uint32_t tmem_base_s;

  HRF_PROF_DECL
  pdl_launch_dependents();
  const int tid = threadIdx.x, warp = warp_idx_uniform(), lane = tid & 31;
  const FfnLayout L(C, K::HID);
  const int tiles_x = ceil_div(p.W, TW), tiles_y = ceil_div(p.H, TH);
  const int total_tiles = gp.n_prob * p.B * tiles_x * tiles_y;
  const bool contig = gp.tiles_per_cta > 0;
  const int t_begin = contig ? blockIdx.x * gp.tiles_per_cta : blockIdx.x;
  const int t_step = contig ? 1 : gridDim.x;
  const int n_tiles = !contig ? total_tiles                       // end of this CTA's tile walk
                              : (t_begin + gp.tiles_per_cta < total_tiles ? t_begin + gp.tiles_per_cta : total_tiles);
  // tile -> (problem, image of the problem, first row, first column)
  auto tile_coords = [&](int tile, int& pr, int& b, int& ty0, int& tx0) {
    int rem, bg;
    p.d_tiles_xy.divmod(tile, bg, rem);
    p.d_tiles_x.divmod(rem, ty0, tx0);
    pr = NP == 1 ? 0 : (bg >= p.B) + (bg >= 2 * p.B) + (bg >= 3 * p.B);
    b = bg - pr * p.B;
    ty0 *= TH;
    tx0 *= TW;
  };
  auto load_weights = [&](int pr) {                    // one thread: the problem's W1 | W2 | conv tables
    const float* blob = gp.blob[pr];
    mbar_expect_tx(&wbar, K::W1_B + K::W2_B + K::CV_B);
    bulk_g2s(sm + K::o_w1, blob + L.o_v2_w1, K::W1_B, &wbar);
    bulk_g2s(sm + K::o_w2, blob + L.o_v2_w2, K::W2_B, &wbar);
    bulk_g2s(sm + K::o_cv, blob + L.o_v2_cv, K::CV_B, &wbar);
  };

  // ---- one-time setup ---------------------------------------------------------------------
  // Thread 0 starts the first tile's box and the weight copies before anything else: their
  // latency (1-2 us) then hides behind the TMEM allocation and the H2 initialisation.
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_init(&wbar, 1);
    mbar_init(&bar1, 1);
    mbar_init(&bar2, 1);
    fence_mbar_init();
    pdl_wait();                                        // x is the predecessor's output
    if (t_begin < n_tiles) {
      int pr, b, ty0, tx0;
      tile_coords(t_begin, pr, b, ty0, tx0);
      mbar_expect_tx(&full[0], K::HH * K::RAW_ROW);
      tma_load_3d(sm + K::o_raw, &tm.x[pr], (tx0 - K::PADL) * (C / 2), ty0 - 1, b, &full[0]);
      load_weights(pr);
      tma_prefetch_desc(&tm.o[pr]);
    }
  }
  {
    // H2: zero, its tenth chunk (hidden channels 72..79 of the chunk, the K padding) holds the
    // constant 1 (fp16) in channel 72 for every row the fc2 MMAs read
    uint4* z = reinterpret_cast<uint4*>(sm + K::o_h2);
    for (int e = tid; e < K::H2_B / 16; e += NT) z[e] = e >= 9 * K::H2R ? make_uint4(0x3C00u, 0, 0, 0) : make_uint4(0, 0, 0, 0);
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, K::TMEM_COLS);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const int q = warp & 3, kq = warp >> 2;              // TMEM quadrant, index among its warps
  const int nq = (NW - q + 3) / 4;                     // warps that share this quadrant
  const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
  const uint32_t a_w1 = smem_u32(sm + K::o_w1), a_w2 = smem_u32(sm + K::o_w2);
  const uint32_t a_xn = smem_u32(sm + K::o_xn), a_h2 = smem_u32(sm + K::o_h2);
  pdl_wait();
  uint32_t ph1 = 0, ph2 = 0, wph = 0;
  int cur_pr = -1;

  HRF_PROF(14)
  int it = 0;
  for (int tile = t_begin; tile < n_tiles; tile += t_step, ++it) {
    HRF_PROF_TILE
    const int s = it & 1;
    int pr, b, ty0, tx0;
    tile_coords(tile, pr, b, ty0, tx0);
    const unsigned char* raw = sm + K::o_raw + s * K::RAW_B;
    // the range crossed into the next problem: its weights (every read of the old ones lies
    // before the previous tile's last barrier)
    if (tid == 0 && cur_pr >= 0 && pr != cur_pr) load_weights(pr);
    // the other stage was last read by the previous tile's epilogue 2 (a barrier ago): refill it
    if (tid == 0 && tile + t_step < n_tiles) {
      int npr, nb, nty0, ntx0;
      tile_coords(tile + t_step, npr, nb, nty0, ntx0);
      mbar_expect_tx(&full[s ^ 1], K::HH * K::RAW_ROW);
      tma_load_3d(sm + K::o_raw + (s ^ 1) * K::RAW_B, &tm.x[npr], (ntx0 - K::PADL) * (C / 2), nty0 - 1, nb, &full[s ^ 1]);
    }
    mbar_wait(&full[s], (it >> 1) & 1, 100 + s);

    // ---- LayerNorm: a thread per row of the operand tile ------------------------------------
    if (tid < K::XR) {
      const int hy = tid / K::HW, hx = tid - hy * K::HW;
      const int gy = ty0 - 1 + hy, gx = tx0 - 1 + hx;
      const bool in = tid < K::NHALO && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
      float2 v[KC / 2];
      if (in) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(raw + hy * K::RAW_ROW + (hx + K::PADL - 1) * (C * 2));
        float2 sa = make_float2(0.f, 0.f), sb = sa, sc = sa;
#pragma unroll
        for (int j = 0; j < C / 2; ++j) {
          const uint32_t u = src[j];
          v[j] = make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
          if (j % 3 == 0) sa = __fadd2_rn(sa, v[j]);
          else if (j % 3 == 1) sb = __fadd2_rn(sb, v[j]);
          else sc = __fadd2_rn(sc, v[j]);
        }
        sa = __fadd2_rn(__fadd2_rn(sa, sb), sc);
        const float nmean = -(sa.x + sa.y) * (1.0f / C);
        const float2 nm2 = make_float2(nmean, nmean);
        float2 qa = make_float2(0.f, 0.f), qb = qa, qc = qa;
#pragma unroll
        for (int j = 0; j < C / 2; ++j) {
          v[j] = __fadd2_rn(v[j], nm2);
          if (j % 3 == 0) qa = __ffma2_rn(v[j], v[j], qa);
          else if (j % 3 == 1) qb = __ffma2_rn(v[j], v[j], qb);
          else qc = __ffma2_rn(v[j], v[j], qc);
        }
        qa = __fadd2_rn(__fadd2_rn(qa, qb), qc);
        const float rstd = rsqrtf((qa.x + qa.y) * (1.0f / C) + p.eps);
        const float2 r2 = make_float2(rstd, rstd);
#pragma unroll
        for (int j = 0; j < C / 2; ++j) v[j] = __fmul2_rn(v[j], r2);
        v[C / 2] = make_float2(1.f, 0.f);              // x the bias row of the W1 tile
#pragma unroll
        for (int j = C / 2 + 1; j < KC / 2; ++j) v[j] = make_float2(0.f, 0.f);
      } else {
#pragma unroll
        for (int j = 0; j < KC / 2; ++j) v[j] = make_float2(0.f, 0.f);   // outside the image: fc1 -> 0 -> GELU -> 0
      }
#pragma unroll
      for (int ch = 0; ch < KC / 8; ++ch) {
        uint4 u;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[4 * ch].x, v[4 * ch].y), h1 = __floats2bfloat162_rn(v[4 * ch + 1].x, v[4 * ch + 1].y);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4 * ch + 2].x, v[4 * ch + 2].y), h3 = __floats2bfloat162_rn(v[4 * ch + 3].x, v[4 * ch + 3].y);
        u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
        u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(sm + K::o_xn + ((size_t)ch * K::XR + tid) * 16) = u;
      }
    }
    HRF_PROF(0)
    if (pr != cur_pr) {                                // first tile / first tile of the next problem
      mbar_wait(&wbar, wph, 102);
      wph ^= 1;
      cur_pr = pr;
    }

#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      HRF_PROF(1)
      // ---- fc1 on the halo M tiles ------------------------------------------------------------
      if (warp == 0 && elect_one()) {
        tc_fence_after();
        constexpr uint32_t id1 = idesc_bf16(128, N1, false, false);
        const uint32_t w1c = a_w1 + c * (N1 * KC * 2);
#pragma unroll
        for (int t = 0; t < K::NMT; ++t)
#pragma unroll
          for (int st = 0; st < KC / 16; ++st)
            mma_bf16(tmem + t * N1, desc_kmajor(a_xn + t * 2048, K::XR, st), desc_kmajor(w1c, N1, st), id1, st > 0);
        mma_commit(&bar1);
      }
      HRF_PROF(2)
      mbar_wait(&bar1, ph1, 103);
      ph1 ^= 1;
      tc_fence_after();
      HRF_PROF(3)

      // ---- epilogue 1: GELU -> fp16 H1; unit = (M tile, 16 accumulator columns) of this quadrant
#pragma unroll 1
      for (int j = kq; j < K::NMT * 5; j += nq) {
        const int mt = j / 5, grp = j - mt * 5;
        const int t = mt * 128 + q * 32 + lane;                  // halo token
        float a[16];
        if (grp < 4) tmem_ld16(trow + mt * N1 + grp * 16, a);
        else tmem_ld8(trow + mt * N1 + 64, a);                   // columns 72..79 are padding
        tmem_ld_wait();
        if (t < K::NHALO) {
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            if (hf == 1 && grp == 4) break;
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (v2_gelu1_half()) {
                // round the fc1 output to fp16 first and evaluate GELU packed (one MUFU.TANH.F16x2 per
                // pair, no separate pack): H1 is stored as fp16 anyway
                w[e] = h2_as_u32(gelu_hx_h2(u32_as_h2(pack_f16x2(a[hf * 8 + 2 * e], a[hf * 8 + 2 * e + 1]))));
              } else {
                const float2 g = gelu_hx2(make_float2(a[hf * 8 + 2 * e], a[hf * 8 + 2 * e + 1]));
                w[e] = pack_f16x2(g.x, g.y);
              }
            }
            *reinterpret_cast<uint4*>(sm + K::o_h1 + ((size_t)(grp * 2 + hf) * K::NHALO + t) * 16) =
                make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
      HRF_PROF(4)
      tc_fence_before();
      __syncthreads();
      HRF_PROF(5)

      // ---- depthwise 3x3 + GELU -> H2: unit = (8-channel chunk, strip of 3 rows, PAIR of columns):
      // the four halo columns of a unit are loaded once (5 rows x 16 bytes each) and feed both
      // outputs columns -- 20 loads per 6 outputs
      {
        constexpr int EW = NT <= 320 ? 2 : 1;                   // output columns per unit
        constexpr int UW = TW / EW, NUNIT = K::NDC * K::NSTRIP * UW;
        const unsigned char* cv = sm + K::o_cv + c * 1600;       // fp16 wd[9][80] | bd[80]
#pragma unroll 1
        for (int id = tid; id < NUNIT; id += NT) {
          const int ch = id / (K::NSTRIP * UW), r = id - ch * (K::NSTRIP * UW);
          const int st = r / UW, ox = (r - st * UW) * EW;
          const int oy0 = st * K::SH;
          const uint4 bq = *reinterpret_cast<const uint4*>(cv + 1440 + ch * 16);
          __half2 acc[EW][K::SH][4];
#pragma unroll
          for (int e = 0; e < EW; ++e)
#pragma unroll
            for (int o = 0; o < K::SH; ++o) {
              acc[e][o][0] = u32_as_h2(bq.x); acc[e][o][1] = u32_as_h2(bq.y);
              acc[e][o][2] = u32_as_h2(bq.z); acc[e][o][3] = u32_as_h2(bq.w);
            }
          const unsigned char* hp = sm + K::o_h1 + ((size_t)ch * K::NHALO + oy0 * K::HW + ox) * 16;
          const unsigned char* wp = cv + ch * 16;
#pragma unroll
          for (int j = 0; j < EW + 2; ++j) {                     // halo column ox + j
            uint4 f[K::SH + 2];
#pragma unroll
            for (int rr = 0; rr < K::SH + 2; ++rr) f[rr] = *reinterpret_cast<const uint4*>(hp + (rr * K::HW + j) * 16);
#pragma unroll
            for (int e = 0; e < EW; ++e) {                       // output column ox + e: tap dx = j - e
              if (j - e < 0 || j - e > 2) continue;
#pragma unroll
              for (int dy = 0; dy < 3; ++dy) {
                const uint4 w = *reinterpret_cast<const uint4*>(wp + (dy * 3 + (j - e)) * 160);
#pragma unroll
                for (int o = 0; o < K::SH; ++o) {
                  acc[e][o][0] = __hfma2(u32_as_h2(f[o + dy].x), u32_as_h2(w.x), acc[e][o][0]);
                  acc[e][o][1] = __hfma2(u32_as_h2(f[o + dy].y), u32_as_h2(w.y), acc[e][o][1]);
                  acc[e][o][2] = __hfma2(u32_as_h2(f[o + dy].z), u32_as_h2(w.z), acc[e][o][2]);
                  acc[e][o][3] = __hfma2(u32_as_h2(f[o + dy].w), u32_as_h2(w.w), acc[e][o][3]);
                }
              }
            }
          }
#pragma unroll
          for (int o = 0; o < K::SH; ++o)
#pragma unroll
            for (int e = 0; e < EW; ++e) {
              uint4 u;
              u.x = h2_as_u32(gelu_hx_h2(acc[e][o][0]));
              u.y = h2_as_u32(gelu_hx_h2(acc[e][o][1]));
              u.z = h2_as_u32(gelu_hx_h2(acc[e][o][2]));
              u.w = h2_as_u32(gelu_hx_h2(acc[e][o][3]));
              *reinterpret_cast<uint4*>(sm + K::o_h2 + ((size_t)ch * K::H2R + (oy0 + o) * TW + ox + e) * 16) = u;
            }
        }
      }
      HRF_PROF(6)
      // (the previous tile's TMA store must have finished reading the output tile before
      // epilogue 2 rewrites it: thread 0 checks on this side of the barrier)
      if (tid == 0 && c == NCH - 1) tma_store_wait_read();
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      HRF_PROF(7)
      // ---- fc2 partial product over this chunk's 72 (padded 80) hidden channels ---------------
      if (warp == 0 && elect_one()) {
        tc_fence_after();
        constexpr uint32_t id2 = idesc_f16(128, NOUT, false, false);
        const uint32_t w2c = a_w2 + c * (NOUT * N1 * 2);
#pragma unroll
        for (int t = 0; t < K::NMT2; ++t)
#pragma unroll
          for (int st = 0; st < N1 / 16; ++st)
            mma_bf16(tmem + K::Y_COL + t * NOUT, desc_kmajor(a_h2 + t * 2048, K::H2R, st),
                     desc_kmajor(w2c, NOUT, st), id2, (c > 0) || (st > 0));
        mma_commit(&bar2);
      }
      HRF_PROF(8)
      mbar_wait(&bar2, ph2, 104);        // H2 (and LN(x) behind it) free again; Y complete after the last chunk
      ph2 ^= 1;
      tc_fence_after();
      HRF_PROF(9)
    }

    // ---- epilogue 2: GELU, + residual (from the raw tile), bf16 -> output tile ----------------
    if (warp < NW / 4 * 4) {
#pragma unroll 1
      for (int g = warp; g < (K::NTOK + 31) / 32; g += NW / 4 * 4) {
        const int ot = g * 32 + lane;                            // output token of the tile
        const int mt2 = g >> 2;
        const int oy = ot / TW, ox = ot - oy * TW;
        float y[NOUT];
#pragma unroll
        for (int c0 = 0; c0 < NOUT; c0 += 16) tmem_ld16(trow + K::Y_COL + mt2 * NOUT + c0, y + c0);
        tmem_ld_wait();
        if (ot >= K::NTOK) continue;
        const uint32_t* res = reinterpret_cast<const uint32_t*>(raw + (oy + 1) * K::RAW_ROW + (ox + K::PADL) * (C * 2));
        uint32_t* dst = reinterpret_cast<uint32_t*>(sm + K::o_out + (size_t)ot * (C * 2));
#pragma unroll
        for (int j = 0; j < C / 2; ++j) {
          const float2 g2 = gelu_hx2(make_float2(y[2 * j], y[2 * j + 1]));
          const uint32_t u = res[j];
          const __nv_bfloat162 hh = __floats2bfloat162_rn(__uint_as_float(u << 16) + g2.x,
                                                          __uint_as_float(u & 0xffff0000u) + g2.y);
          dst[j] = *reinterpret_cast<const uint32_t*>(&hh);
        }
      }
    }
    HRF_PROF(10)
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tma_store_3d(&tm.o[pr], tx0 * (C / 2), ty0, b, sm + K::o_out);
      tma_store_commit();
    }
    HRF_PROF(11)
  }

  HRF_PROF_END
  if (tid == 0) tma_store_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, K::TMEM_COLS);
}


// ---- host side --------------------------------------------------------------------------------
// true when the token tensor can be described to the TMA unit this way
static bool tmap_tokens_ok(const void* base, int W, int C) {
  return (reinterpret_cast<uintptr_t>(base) & 15) == 0 && C % 2 == 0 && ((size_t)W * C * 2) % 16 == 0;
}
static int make_tmap_tokens(CUtensorMap* m, const void* base, int B, int H, int W, int C, int box_tokens,
                            int box_rows) {
  PFN_tmapEncodeTiled enc = tmap_encoder();
  HRF_REQUIRE(enc != nullptr, HRF_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t gdim[3] = {(cuuint64_t)W * C / 2, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t gstr[2] = {(cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  const cuuint32_t box[3] = {(cuuint32_t)(box_tokens * C / 2), (cuuint32_t)box_rows, 1u};
  const cuuint32_t est[3] = {1u, 1u, 1u};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<void*>(base), gdim, gstr, box, est,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  HRF_REQUIRE(r == CUDA_SUCCESS, HRF_ECUDA, "cuTensorMapEncodeTiled failed (%d) for B=%d H=%d W=%d C=%d box=%dx%d",
              (int)r, B, H, W, C, box_tokens, box_rows);
  return HRF_OK;
}

// HRF_FFN_V2=0 keeps the first-generation kernel (A/B runs)
static bool ffn_v2_enabled() {
  const char* e = std::getenv("HRF_FFN_V2");        // read per call: A/B within one process
  return !(e && e[0] == '0');
}
static bool ffn_v2_supported(const FfnParams& p) {
  return ffn_v2_enabled() && p.hidden == 4 * p.C && p.C == 18 && tmap_tokens_ok(p.x, p.W, p.C) &&
         tmap_tokens_ok(p.out, p.W, p.C) && p.W * p.C / 2 >= 1;
}

// one problem: (p.x, p.blob, p.out); several: n_prob entries of xs / blobs / outs
struct FfnV2Problems {
  int n = 1;
  const void* x[kFfnMaxProb];
  const float* blob[kFfnMaxProb];
  void* out[kFfnMaxProb];
};

template <int C, int TH, int TW, int NT, int NP>
static int launch_ffn_v2_np(FfnParams p, const FfnV2Problems& pb, cudaStream_t stream) {
  using K = FfnV2<C, TH, TW, NT>;
  const int tiles_x = ceil_div(p.W, TW), tiles_y = ceil_div(p.H, TH);
  const int n_tiles = pb.n * p.B * tiles_x * tiles_y;
  p.d_tiles_x = FastDiv(tiles_x);
  p.d_tiles_xy = FastDiv(tiles_x * tiles_y);
  FfnV2Maps<NP> tm;
  FfnV2Group gp{};
  gp.n_prob = pb.n;
  for (int q = 0; q < kFfnMaxProb; ++q) gp.blob[q] = pb.blob[q < pb.n ? q : 0];
  for (int q = 0; q < NP; ++q) {
    const int s = q < pb.n ? q : 0;
    int rc = make_tmap_tokens(&tm.x[q], pb.x[s], p.B, p.H, p.W, C, K::BOXW, K::HH);
    if (rc) return rc;
    rc = make_tmap_tokens(&tm.o[q], pb.out[s], p.B, p.H, p.W, C, TW, TH);
    if (rc) return rc;
    HRF_REQUIRE((reinterpret_cast<uintptr_t>(pb.blob[s]) & 15) == 0, HRF_EINVAL, "mixffn_v2: blob must be 16-byte aligned");
  }
  static const int env_per_sm = [] { const char* e = std::getenv("HRF_FFN_CTAS_PER_SM"); return e ? atoi(e) : 0; }();
  const int per_sm = env_per_sm > 0 && env_per_sm < K::CTAS_PER_SM ? env_per_sm : K::CTAS_PER_SM;
  const int cap = 148 * per_sm;
  // HRF_BALANCED_GRID=1 (experiment): ceil(tiles / rounds) CTAs that all walk `rounds` tiles instead of
  // one full resident wave -- the same step time (2.799 vs 2.802 ms: the slots it frees for the other
  // streams' kernels buy nothing), and a slower kernel on its own (the last, partly filled round of
  // a full wave runs uncontended: MixFFN 19.8 vs 23.0 us), so the full wave is the default
  static const bool balanced = [] { const char* e = std::getenv("HRF_BALANCED_GRID"); return e && e[0] == '1'; }();
  const int rounds = ceil_div(n_tiles, cap);
  int grid = !balanced ? (n_tiles < cap ? n_tiles : cap) : ceil_div(n_tiles, rounds);
  if (pb.n > 1) {                                      // several problems: contiguous ranges (weights reload at a crossing)
    gp.tiles_per_cta = ceil_div(n_tiles, grid);
    grid = ceil_div(n_tiles, gp.tiles_per_cta);
  }
  HRF_CUDA(ensure_smem((const void*)mixffn_v2_kernel<C, TH, TW, NT, NP>, K::SMEM));
  HRF_CUDA(launch_pdl(mixffn_v2_kernel<C, TH, TW, NT, NP>, dim3(grid), dim3(NT), K::SMEM, stream, p, gp, tm));
  count_launch();
  HRF_CUDA(cudaGetLastError());
  return HRF_OK;
}

template <int C, int TH, int TW, int NT>
static int launch_ffn_v2_t(FfnParams p, const FfnV2Problems& pb, cudaStream_t stream) {
  return pb.n == 1 ? launch_ffn_v2_np<C, TH, TW, NT, 1>(p, pb, stream)
                   : launch_ffn_v2_np<C, TH, TW, NT, kFfnMaxProb>(p, pb, stream);
}

static int launch_mixffn_v2(const FfnParams& p, cudaStream_t stream, const FfnV2Problems* group = nullptr) {
  FfnV2Problems pb;
  if (group) pb = *group;
  else { pb.n = 1; pb.x[0] = p.x; pb.blob[0] = p.blob; pb.out[0] = p.out; }
  // 12 x 16 tile, 576 threads (two CTAs = 36 warps per SM at 56 registers; depthwise units of 3 x 1
  // outputs): 21.7 us at 96 x 160 x 8 by CUDA events against 23.5 us for 288 threads / 3 x 2 units and
  // 27.0 us for the first-generation kernel.  The other tile shapes that were measured (6 x 16, 9 x 16,
  // 12 x 16 at 288 / 384 / 448 threads; 12 x 18) are in the git history.
  switch (p.C) {
    case 18: return launch_ffn_v2_t<18, 12, 16, 576>(p, pb, stream);
  }
  HRF_REQUIRE(false, HRF_EUNSUPPORTED, "mixffn_v2: C=%d", p.C);
}

}  // namespace hrf
