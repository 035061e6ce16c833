// MixFFN (CrossFFN) with ALL THREE layers on the tensor cores (bf16 mode).
//
// Reference: HRFuser `CrossFFN` (/root/reference/mmdet/models/backbones/hrformer.py:248-306,
// used by HRFomerBlock.forward :399-406):   out = x + GELU(BN(fc2(GELU(BN(dw3x3(GELU(BN(fc1(LN(x))))))))))
//
// `mixffn_tc.cuh` runs fc1 / fc2 on tcgen05 and the depthwise 3x3 on the CUDA cores, where it
// is ~60 % of the kernel's issued instructions.  Here the depthwise conv is tensor-core work
// as well: with the hidden activations H1 stored token-major in the chunk-major operand layout
// (16 bytes per token and 8-channel chunk), "the same tile shifted by (dy, dx) tokens" is just
// the A-operand descriptor's start address moved by (dy*16 + dx) * 16 bytes, so
//
//     D2[r, ch] = sum_tap  H1[r + dy*16 + dx, ch] * wd[tap][ch]
//               = sum_tap  (A = H1 shifted by tap)  x  (B = diag(wd[tap]))
//
// is 9 taps x 5 column groups of M=128, N=16, K=16 MMAs (8 cycles each by the tcgen05 floor of
// 128*N/256 cycles) that run asynchronously next to the CUDA-core epilogues of the co-resident
// CTAs.  The tile is 6 x 14 outputs inside an 8 x 16 = 128 token halo, so output token
// (oy, ox) is row r = oy*16 + ox <= 93 of every M=128 tile and reads halo rows <= 127.
//
// All biases ride through the MMAs: b1 via a constant-1 column C of LN(x) (zero for halo
// tokens outside the image, so H1 = GELU(0) = 0 there, the conv's zero padding), bd via the
// constant-1 activation column 72 (the K padding of the 72-channel chunk) and one extra
// "bias tile" MMA per column group.  One 20 KB buffer holds LN(x), then H1, then H2 in place.
#pragma once
#include "mixffn_tc.cuh"

namespace hrf {

template <int C, int CPG>
struct FfnTcd {
  static constexpr int HID = 4 * C, NCH = HID / 72, NG = NCH / CPG;
  static constexpr bool SPLIT = CPG < NCH, BIGC = C > 40;
  static_assert(HID % 72 == 0 && NCH % CPG == 0, "hidden must split into 72-channel chunks");
  static constexpr int KC = (C + 15) / 16 * 16, NOUT = KC, N1 = 80;
  static_assert(KC > C, "needs a spare K column for the fc1 bias");
  static constexpr int TH = 6, TW = 14, HW = 16, HH = 8, NHALO = 128;
  static constexpr int NT = 256, NGQ = NT / 128;
  static constexpr int H_B = 10 * 128 * 16;           // 80 channels x 128 rows, bf16
  static constexpr int XN_B = 128 * KC * 2;
  // LN(x) may live in the H buffer when it is consumed once (one chunk) and does not reach
  // the constant-1 column (chunk 9)
  static constexpr bool XN_ALIAS = (CPG == 1) && (KC / 8 <= 9);
  static constexpr int DG_B = 50 * 512;               // per chunk: 5 groups x (9 taps + bias)
  // shared-memory map (bytes).  H first: the shifted A operands read up to 34 rows past it.
  static constexpr int o_h = 0;
  static constexpr int o_xn = XN_ALIAS ? o_h : o_h + H_B;
  static constexpr int o_w1 = XN_ALIAS ? o_h + H_B : o_xn + XN_B;   // CPG tiles [80 x KC]
  static constexpr int o_w2 = o_w1 + CPG * N1 * KC * 2;             // CPG tiles [NOUT x 80]
  static constexpr int o_dg = o_w2 + CPG * NOUT * N1 * 2;           // CPG x 50 tiles [16 x 16]
  static constexpr int o_f32 = o_dg + CPG * DG_B;                   // b2[NOUT]
  static constexpr int C4 = (C + 3) / 4 * 4;
  static constexpr int o_ln = o_f32 + NOUT * 4;                     // gamma[C4] beta[C4]
  static constexpr int SMEM = o_ln + 2 * C4 * 4;
  static constexpr int D_COL = 0;                     // fc1 accumulator, then the conv's
  static constexpr int Y_COL = N1;                    // fc2 accumulator
  static constexpr int TMEM_COLS = (N1 + NOUT <= 128) ? 128 : 256;
  static constexpr int CTAS_PER_SM = (SMEM + 1040) * 4 <= 228 * 1024 && TMEM_COLS <= 128 ? 4
                                   : (SMEM + 1040) * 3 <= 228 * 1024 && TMEM_COLS <= 128 ? 3
                                   : (SMEM + 1040) * 2 <= 228 * 1024 ? 2 : 1;
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

template <int C, int CPG>
__global__ void __launch_bounds__(FfnTcd<C, CPG>::NT, FfnTcd<C, CPG>::CTAS_PER_SM)
mixffn_tcd_kernel(FfnParams p) {
  using namespace umma;
  using K = FfnTcd<C, CPG>;
  constexpr int KC = K::KC, NOUT = K::NOUT, N1 = K::N1, NCH = K::NCH, NG = K::NG;
  constexpr int NT = K::NT, NGQ = K::NGQ;
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;

  HRF_PROF_DECL
  pdl_launch_dependents();
  const int tid = threadIdx.x, warp = warp_idx_uniform(), lane = tid & 31;
  const int gq = warp >> 2, q = warp & 3;          // work group, TMEM quadrant
  const int row = q * 32 + lane;                   // TMEM lane == halo token == tile row
  const int cg = blockIdx.x % NG;                  // chunk group of this CTA
  const FfnLayout L(C, K::HID);
  const float* blob = p.blob;
  float* sB2 = reinterpret_cast<float*>(sm + K::o_f32);
  float* sLn = reinterpret_cast<float*>(sm + K::o_ln);

  // ---- one-time setup ---------------------------------------------------------------
  {
    const uint4* s1 = reinterpret_cast<const uint4*>(blob + L.o_tc_w1) + (size_t)cg * CPG * (N1 * KC * 2 / 16);
    const uint4* s2 = reinterpret_cast<const uint4*>(blob + L.o_tc_w2) + (size_t)cg * CPG * (NOUT * N1 * 2 / 16);
    const uint4* s3 = reinterpret_cast<const uint4*>(blob + L.o_tc_dg) + (size_t)cg * CPG * (K::DG_B / 16);
    uint4* d1 = reinterpret_cast<uint4*>(sm + K::o_w1);
    uint4* d2 = reinterpret_cast<uint4*>(sm + K::o_w2);
    uint4* d3 = reinterpret_cast<uint4*>(sm + K::o_dg);
    for (int e = tid; e < CPG * N1 * KC * 2 / 16; e += NT) d1[e] = __ldg(s1 + e);
    for (int e = tid; e < CPG * NOUT * N1 * 2 / 16; e += NT) d2[e] = __ldg(s2 + e);
    for (int e = tid; e < CPG * K::DG_B / 16; e += NT) d3[e] = __ldg(s3 + e);
    for (int e = tid; e < NOUT; e += NT) sB2[e] = __ldg(blob + L.o_tc_f32 + NCH * 880 + e);
    for (int e = tid; e < K::C4; e += NT) {
      sLn[e] = __ldg(blob + L.o_ln_w + e);
      sLn[K::C4 + e] = __ldg(blob + L.o_ln_b + e);
    }
    // H: zero, then channel 72 (chunk 9, element 0) = 1 in every row
    uint4* z = reinterpret_cast<uint4*>(sm + K::o_h);
    for (int e = tid; e < 9 * 128; e += NT) z[e] = make_uint4(0, 0, 0, 0);
    for (int e = tid; e < 128; e += NT) z[9 * 128 + e] = make_uint4(0x00003F80u, 0, 0, 0);
    if (!K::XN_ALIAS) {
      z = reinterpret_cast<uint4*>(sm + K::o_xn);
      for (int e = tid; e < K::XN_B / 16; e += NT) z[e] = make_uint4(0, 0, 0, 0);
    }
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, K::TMEM_COLS);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
  uint32_t phase = 0;
  const uint32_t a_w1 = smem_u32(sm + K::o_w1), a_w2 = smem_u32(sm + K::o_w2);
  const uint32_t a_xn = smem_u32(sm + K::o_xn), a_h = smem_u32(sm + K::o_h);
  const uint32_t a_dg = smem_u32(sm + K::o_dg);

  const int tiles_x = ceil_div(p.W, K::TW), tiles_y = ceil_div(p.H, K::TH);
  const int n_tiles = p.B * tiles_x * tiles_y;
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(p.x);
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(p.out);

  // fc1 of chunk c: D[128 x 80] = LN(x)[128 x KC] . W1_c^T   (one thread)
  auto issue_fc1 = [&](int c) {
    constexpr uint32_t id1 = idesc_bf16(128, N1, false, false);
    const uint32_t w1c = a_w1 + c * (N1 * KC * 2);
#pragma unroll
    for (int s = 0; s < KC / 16; ++s)
      mma_bf16(tmem + K::D_COL, desc_kmajor(a_xn, 128, s), desc_kmajor(w1c, N1, s), id1, s > 0);
  };

  // Software pipeline (C <= 40): halo token `tid` of the NEXT tile is requested while the
  // current tile computes.
  constexpr bool PIPE = !K::BIGC;
  constexpr int NW = PIPE ? C / 2 : 1;
  const int tile_step = gridDim.x / NG;
  auto halo_token = [&](int tile) -> int {
    if (tile >= n_tiles || tid >= K::NHALO) return -1;
    int b, rem, ty, tx;
    p.d_tiles_xy.divmod(tile, b, rem);
    p.d_tiles_x.divmod(rem, ty, tx);
    const int h = ty * K::TH - 1 + (tid >> 4);
    const int w = tx * K::TW - 1 + (tid & 15);
    return (h >= 0 && h < p.H && w >= 0 && w < p.W) ? (b * p.H + h) * p.W + w : -1;
  };
  pdl_wait();
  uint32_t xr[NW];
  int htok = halo_token(blockIdx.x / NG);
  if (PIPE && htok >= 0) load_row_raw<C>(x + (size_t)htok * C, xr);

  // output token of this thread's row: r = oy*16 + ox
  const int oy = row >> 4, ox = row & 15;
  const bool o_row = (oy < K::TH) && (ox < K::TW);

  HRF_PROF(14)                                     // setup
  for (int tile = blockIdx.x / NG; tile < n_tiles; tile += tile_step) {
    int b, rem, ty0, tx0;
    p.d_tiles_xy.divmod(tile, b, rem);
    p.d_tiles_x.divmod(rem, ty0, tx0);
    ty0 *= K::TH;
    tx0 *= K::TW;
    HRF_PROF_TILE

    // ---- LN prologue: halo token `tid` -> row `tid` of the LN(x) tile -----------------
    if (tid < K::NHALO) {
      unsigned char* xt = sm + K::o_xn;
      if (htok >= 0) {
        if constexpr (PIPE) {
          float v[C];
          unpack_row<C>(xr, v);
          ln_row_to_tile<C, KC, true>(v, sLn, sLn + K::C4, p.eps, xt, tid);
        } else {
          ln_token<C, KC, true, true>(x + (size_t)htok * C, sLn, sLn + K::C4, p.eps, xt, tid);
        }
      } else {
        const float zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ch = 0; ch < KC / 8; ++ch) st_chunk(xt, tid, ch, 128, zero);
      }
    }
    const int htok_next = halo_token(tile + tile_step);
    uint32_t xnext[NW];
    if (PIPE && htok_next >= 0) load_row_raw<C>(x + (size_t)htok_next * C, xnext);
    const int oh = ty0 + oy, ow = tx0 + ox;
    const bool o_in = o_row && oh < p.H && ow < p.W;
    const size_t o_tok = o_in ? (size_t)(b * p.H + oh) * p.W + ow : 0;
    uint32_t rres[4] = {0u, 0u, 0u, 0u};           // residual of epilogue-2 unit cc = gq
    if (!K::SPLIT && o_in && gq * 8 < C) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (gq * 8 + 2 * j < C) rres[j] = __ldg(reinterpret_cast<const uint32_t*>(x + o_tok * C + gq * 8) + j);
    }

    HRF_PROF(0)                                    // LN prologue
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    HRF_PROF(1)                                    // barrier skew
    if (warp == 0 && elect_one()) {
      tc_fence_after();
      issue_fc1(0);
      mma_commit(&bar);
    }
    HRF_PROF(2)                                    // fc1 issue
    cta_wait(&bar, phase);
    phase ^= 1;
    tc_fence_after();
    HRF_PROF(3)                                    // fc1 wait

#pragma unroll 1
    for (int c = 0; c < CPG; ++c) {
      // ---- epilogue 1: H1 = GELU(fc1) (bias and zero padding came through the MMA) -----
#pragma unroll 1
      for (int ch = gq; ch < 9; ch += NGQ) {
        float v[8];
        tmem_ld8(trow + K::D_COL + ch * 8, v);
        tmem_ld_wait();
        gelu8(v);
        st_chunk(sm + K::o_h, row, ch, 128, v);
      }

      HRF_PROF(4)                                  // epilogue 1
      // ---- depthwise 3x3 (+ bd) on the tensor cores: D = sum_tap shift_tap(H1) . diag(wd_tap)
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      HRF_PROF(5)
      if (warp == 0 && elect_one()) {
        tc_fence_after();
        constexpr uint32_t idd = idesc_bf16(128, 16, false, false);
        const uint32_t dgc = a_dg + c * K::DG_B;
#pragma unroll 1
        for (int s = 0; s < 5; ++s) {
          const uint32_t a_s = a_h + (uint32_t)s * (2 * 128 * 16);      // channels 16s..16s+15
          const uint32_t b_s = dgc + (uint32_t)s * (10 * 512);
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const uint32_t shift = (uint32_t)((t / 3) * K::HW + (t % 3)) * 16u;
            mma_bf16(tmem + K::D_COL + s * 16, smem_desc(a_s + shift, 128 * 16, 128),
                     smem_desc(b_s + t * 512, 256, 128), idd, t > 0);
          }
          // bias tile against the constant-1 column: A = chunks 8, 9 (un-shifted)
          mma_bf16(tmem + K::D_COL + s * 16, smem_desc(a_h + 8 * (128 * 16), 128 * 16, 128),
                   smem_desc(b_s + 9 * 512, 256, 128), idd, true);
        }
        mma_commit(&bar);
      }
      HRF_PROF(6)                                  // conv issue
      cta_wait(&bar, phase);
      phase ^= 1;
      tc_fence_after();
      HRF_PROF(7)                                  // conv wait

      // ---- epilogue dw: H2 = GELU(conv), in place over H1 (rows of real output tokens) ---
      if (q < 3) {                                   // rows 96..127 hold no output token
#pragma unroll 1
        for (int ch = gq; ch < 9; ch += NGQ) {
          float v[8];
          tmem_ld8(trow + K::D_COL + ch * 8, v);
          tmem_ld_wait();
          if (o_row) {
            gelu8(v);
            st_chunk(sm + K::o_h, row, ch, 128, v);
          }
        }
      }

      HRF_PROF(8)                                  // epilogue dw
      // ---- fc2 partial product over this chunk (and fc1 of the next chunk behind it) -----
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      HRF_PROF(9)
      if (warp == 0 && elect_one()) {
        tc_fence_after();
        constexpr uint32_t id2 = idesc_bf16(128, NOUT, false, false);
        const uint32_t w2c = a_w2 + c * (NOUT * N1 * 2);
#pragma unroll
        for (int s = 0; s < N1 / 16; ++s)
          mma_bf16(tmem + K::Y_COL, desc_kmajor(a_h, 128, s), desc_kmajor(w2c, NOUT, s), id2,
                   (c > 0) || (s > 0));
        if (c + 1 < CPG) issue_fc1(c + 1);
        mma_commit(&bar);
      }
      HRF_PROF(10)                                 // fc2 issue
      cta_wait(&bar, phase);
      phase ^= 1;
      tc_fence_after();
      HRF_PROF(11)                                 // fc2 wait
    }

    // ---- epilogue 2: unit = (output token, 8-channel chunk of the C outputs) ----------
    if (q < 3) {
#pragma unroll 1
      for (int cc = gq; cc * 8 < C; cc += NGQ) {
        float y[8];
        tmem_ld8(trow + K::Y_COL + cc * 8, y);
        tmem_ld_wait();
        if (o_in) {
        if constexpr (K::SPLIT) {      // fp32 partial of this chunk group -> workspace [NG][n_tok][C]
          const size_t n_tok = (size_t)p.B * p.H * p.W;
          float* wrow = static_cast<float*>(p.ws) + ((size_t)cg * n_tok + o_tok) * C + cc * 8;
          *reinterpret_cast<float4*>(wrow) = make_float4(y[0], y[1], y[2], y[3]);
          *reinterpret_cast<float4*>(wrow + 4) = make_float4(y[4], y[5], y[6], y[7]);
        } else {                       // + b2, GELU, + residual; rows are only 4-byte aligned
          const uint32_t* xr4 = reinterpret_cast<const uint32_t*>(x + o_tok * C + cc * 8);
          uint32_t* orow = reinterpret_cast<uint32_t*>(out + o_tok * C + cc * 8);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (cc * 8 + 2 * j < C) {
              const uint32_t u = (cc == gq) ? rres[j] : __ldg(xr4 + j);
              const float2 r = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
              const float a0 = r.x + gelu_as(y[2 * j]);       // b2 came through the MMA
              const float a1 = r.y + gelu_as(y[2 * j + 1]);
              const __nv_bfloat162 hh = __floats2bfloat162_rn(a0, a1);
              orow[j] = *reinterpret_cast<const uint32_t*>(&hh);
            }
          }
        }
        }
      }
    }
    HRF_PROF(12)                                   // epilogue 2
    htok = htok_next;
    if constexpr (PIPE) {
#pragma unroll
      for (int j = 0; j < NW; ++j) xr[j] = xnext[j];
    }
    // the next tile's first barrier orders these TMEM reads before its MMAs
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, K::TMEM_COLS);
}

static bool ffn_tcd_supported(const FfnParams& p) {
  // experimental variant, selected with HRF_FFN_DW_TC=1: it is bound by the tensor pipe's
  // operand fetch (50 M=128,N=16 MMAs of ~50 cycles per tile) and loses to the CUDA-core conv
  static const bool on = [] { const char* e = std::getenv("HRF_FFN_DW_TC"); return e && atoi(e) != 0; }();
  return on && p.hidden == 4 * p.C && p.C == 18;
}

template <int C, int CPG>
static int launch_ffn_tcd_c(FfnParams p, cudaStream_t stream) {
  using K = FfnTcd<C, CPG>;
  const int n_tiles = p.B * ceil_div(p.H, K::TH) * ceil_div(p.W, K::TW);
  p.d_tiles_x = FastDiv(ceil_div(p.W, K::TW));
  p.d_tiles_xy = FastDiv(ceil_div(p.H, K::TH) * ceil_div(p.W, K::TW));
  static const int env_per_sm = [] { const char* e = std::getenv("HRF_FFN_CTAS_PER_SM"); return e ? atoi(e) : 0; }();
  const int per_sm = env_per_sm > 0 ? env_per_sm : K::CTAS_PER_SM;
  const int cap = 148 * per_sm / K::NG > 0 ? 148 * per_sm / K::NG : 1;
  const int grid = (n_tiles < cap ? n_tiles : cap) * K::NG;
  if (K::SPLIT) HRF_REQUIRE(p.ws != nullptr, HRF_EINVAL, "mixffn_tcd: workspace required for C=%d", C);
  HRF_CUDA(ensure_smem((const void*)mixffn_tcd_kernel<C, CPG>, K::SMEM));
  HRF_CUDA(launch_pdl(mixffn_tcd_kernel<C, CPG>, dim3(grid), dim3(K::NT), K::SMEM, stream, p));
  count_launch();
  HRF_CUDA(cudaGetLastError());
  return HRF_OK;
}

static int launch_mixffn_tcd(const FfnParams& p, cudaStream_t stream) {
  switch (p.C) {
    case 18: return launch_ffn_tcd_c<18, 1>(p, stream);
  }
  HRF_REQUIRE(false, HRF_EUNSUPPORTED, "mixffn_tcd: C=%d", p.C);
}

}  // namespace hrf
