// Transition conv (3x3, pad 1, stride 1|2, + folded BN, ReLU) on the tcgen05 tensor cores for
// the C = 18 modality stream -- bf16 mode.
//
// Every fusion stage sends the 18-channel modality tokens through chains of these convs
// (18 -> 18 -> .. -> {36, 72, 144}, reference hrfuser_hrformer_based.py
// `_make_transition_layer_modality`), 18 launches per step.  On the CUDA cores they are bound
// by the shared-memory weight reads (~14 TFLOP/s, tools/conv_bench.py); as an implicit GEMM
//
//     M = 128 output tokens,  K = 9 taps x 32 (18 channels + bias column + zero pad),
//     N = Cout rounded up to 16
//
// one CTA gathers the nine taps of its 128 tokens straight into nine UMMA operand tiles (a
// lane pair per token, every load of the tile in flight at once), issues 18 MMAs and stores
// ReLU(acc).  The bias rides on a constant-1 column of the centre tap (always inside the
// image).  One tile per CTA; the weight tiles arrive by a bulk async copy behind the gather.
#pragma once
#include "common.cuh"
#include "hrfuse.cuh"
#include "umma.cuh"

namespace hrf {

// Supported (Cin, Cout): the 18-channel modality stream into any width <= 144, and the camera's
// new-branch convs 36 -> 72 and 72 -> 144 (the latter with the output channels split over two
// CTAs and the taps gathered in three passes, for shared memory).
struct ConvTcCfg {
  int KC, NOUT, NSPLIT, TPP;   // K per tap, N per CTA (padded), CTAs per tile along N, taps per pass
  bool ok;
  __host__ __device__ ConvTcCfg(int cin, int cout) {
    KC = round_up(cin + 1, 16);                       // + the bias column
    NSPLIT = (cin == 72 && cout == 144) ? 2 : 1;
    NOUT = round_up(cout / NSPLIT, 16);
    TPP = cin == 72 ? 3 : 9;
    ok = (cin == 18 && cout % 2 == 0 && (NOUT == 32 || NOUT == 48 || NOUT == 80 || NOUT == 144)) ||
         (cin == 36 && cout == 72) || (cin == 72 && cout == 144);
  }
};

// TC section of the conv3x3 blob (appended to PwLayout(9*Cin, Cout)): per N split, nine bf16 B
// tiles [KC/8][NOUT rows][8] (chunk-major, K = channel; row Cin of the centre tap = folded bias)
struct ConvTcLayout {
  int NOUT, KC, NSPLIT, o_w, total;   // floats
  __host__ __device__ ConvTcLayout(int cin, int cout) {
    const int base = PwLayout(9 * cin, cout).total;
    const ConvTcCfg c(cin, cout);
    NOUT = c.NOUT; KC = c.KC; NSPLIT = c.NSPLIT;
    o_w = round_up(base, 4);
    total = c.ok ? o_w + NSPLIT * 9 * NOUT * KC / 2 : base;
  }
};

template <int CIN, int NOUT, int TPP>
__global__ void __launch_bounds__(256) conv3x3_tc_kernel(Conv3Params p) {
  using namespace umma;
  constexpr int KC = (CIN + 1 + 15) / 16 * 16;
  constexpr int TCOLS = NOUT <= 32 ? 32 : NOUT <= 64 ? 64 : NOUT <= 128 ? 128 : 256;
  constexpr int A_B = 128 * KC * 2;                 // one tap's A tile
  constexpr int W_B = NOUT * KC * 2;                // one tap's B tile
  constexpr int NWR = CIN / 2;                      // packed words per token row
  constexpr int WH0 = (NWR / 2) / 4 * 4;            // half 0: words [0, WH0), whole chunks
  constexpr int NW1 = NWR - WH0;                    // half 1: words [WH0, NWR) + the bias word
  constexpr int NWT = NW1 > WH0 ? NW1 : WH0;        // words per thread
  constexpr int NCH0 = WH0 / 4, NCH1 = (NW1 + 1 + 3) / 4;
  static_assert(NCH0 + NCH1 <= KC / 8 && WH0 >= 4, "chunk split");
  extern __shared__ __align__(128) unsigned char sm[];   // [TPP A tiles][9 B tiles]
  __shared__ __align__(8) uint64_t bar, wbar;
  __shared__ uint32_t tmem_base_s;
  unsigned char* sA = sm;
  unsigned char* sW = sm + TPP * A_B;

  pdl_launch_dependents();
  const int tid = threadIdx.x, warp = warp_idx_uniform(), lane = tid & 31;
  const ConvTcLayout L(CIN, p.Cout);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_init(&wbar, 1);
    fence_mbar_init();
    mbar_expect_tx(&wbar, 9 * W_B);
    bulk_g2s(sW, reinterpret_cast<const unsigned char*>(p.blob + L.o_w) + (size_t)blockIdx.y * 9 * W_B, 9 * W_B, &wbar);
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, TCOLS);
  // chunks past the data + bias column are zero padding (written once)
  constexpr int NZ = KC / 8 - NCH0 - NCH1;           // all-zero chunks per tap
  if constexpr (NZ > 0) {
    for (int e = tid; e < TPP * NZ * 128; e += 256) {
      const int tp = e / (NZ * 128), rest = e - tp * (NZ * 128);
      *reinterpret_cast<uint4*>(sA + tp * A_B + (NCH0 + NCH1 + (rest >> 7)) * 2048 + (rest & 127) * 16) =
          make_uint4(0, 0, 0, 0);
    }
  }
  __syncthreads();                                  // barriers initialised, TMEM address published
  pdl_wait();

  // ---- gather: lane pair (2r, 2r+1) owns token r ---------------------------------------------
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(p.x);
  const int ntok = p.B * p.Ho * p.Wo;
  const int r = tid >> 1, half = tid & 1;
  const int t = blockIdx.x * 128 + r;
  const bool tv = t < ntok;
  int bb, rem, oy, ox;
  p.d_howo.divmod(tv ? t : 0, bb, rem);
  p.d_wo.divmod(rem, oy, ox);
  const uint32_t tmem = tmem_base_s;
  const int nw = half ? NW1 : WH0;
  uint32_t phase = 0;

#pragma unroll 1
  for (int pass = 0; pass < 9 / TPP; ++pass) {
    uint32_t w[TPP][NWT];
#pragma unroll
    for (int tp = 0; tp < TPP; ++tp) {
      const int tap = pass * TPP + tp;
      const int iy = oy * p.stride - 1 + tap / 3, ix = ox * p.stride - 1 + tap % 3;
      const bool in = tv && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
      const uint32_t* src = reinterpret_cast<const uint32_t*>(
          x + ((size_t)(bb * p.H + (in ? iy : 0)) * p.W + (in ? ix : 0)) * CIN) + half * WH0;
#pragma unroll
      for (int j = 0; j < NWT; ++j) w[tp][j] = (in && j < nw) ? __ldg(src + j) : 0u;
    }
    if (pass > 0) {                                 // the previous pass's MMAs still read the A tiles
      cta_wait(&bar, phase);
      phase ^= 1;
    }
#pragma unroll
    for (int tp = 0; tp < TPP; ++tp) {
      const int tap = pass * TPP + tp;
      unsigned char* a = sA + tp * A_B + r * 16;
      // constant-1 column (channel CIN) on the centre tap carries the bias
      const uint32_t one = (tap == 4 && tv) ? 0x00003F80u : 0u;
      if (half == 0) {
#pragma unroll
        for (int c = 0; c < NCH0; ++c)
          *reinterpret_cast<uint4*>(a + c * 2048) =
              make_uint4(w[tp][4 * c], w[tp][4 * c + 1], w[tp][4 * c + 2], w[tp][4 * c + 3]);
      } else {
#pragma unroll
        for (int c = 0; c < NCH1; ++c) {
          uint32_t v[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int j = 4 * c + k;
            v[k] = j < NW1 ? w[tp][j < NWT ? j : 0] : (j == NW1 ? one : 0u);
          }
          *reinterpret_cast<uint4*>(a + (NCH0 + c) * 2048) = make_uint4(v[0], v[1], v[2], v[3]);
        }
      }
    }
    if (pass == 0) mbar_wait(&wbar, 0);             // weights have landed
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (warp == 0 && elect_one()) {
      tc_fence_after();
      constexpr uint32_t id = idesc_bf16(128, NOUT, false, false);
      const uint32_t a_a = smem_u32(sA), a_w = smem_u32(sW);
#pragma unroll
      for (int tp = 0; tp < TPP; ++tp)
#pragma unroll
        for (int s = 0; s < KC / 16; ++s)
          mma_bf16(tmem, desc_kmajor(a_a + tp * A_B, 128, s),
                   desc_kmajor(a_w + (pass * TPP + tp) * W_B, NOUT, s), id, pass > 0 || tp > 0 || s > 0);
      mma_commit(&bar);
    }
  }
  cta_wait(&bar, phase);
  tc_fence_after();

  // ---- epilogue: ReLU, bf16 -> global.  Warp w: TMEM quadrant w % 4, column group w / 4 ------
  {
    const int q = warp & 3, gq = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    const int to = blockIdx.x * 128 + row;
    const int cper = p.Cout / (int)gridDim.y, co0 = blockIdx.y * cper;   // this CTA's output channels
    __nv_bfloat16* out = static_cast<__nv_bfloat16*>(p.out);
    const int ncc = (cper + 7) / 8;
#pragma unroll 1
    for (int cc = gq; cc < NOUT / 8; cc += 2) {      // warp-uniform trip count
      float y[8];
      tmem_ld8(trow + cc * 8, y);
      tmem_ld_wait();
      if (to < ntok && cc < ncc) {
        uint32_t* orow = reinterpret_cast<uint32_t*>(out + (size_t)to * p.Cout + co0 + cc * 8);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (cc * 8 + 2 * j < cper) {
            float a0 = y[2 * j], a1 = y[2 * j + 1];
            if (p.relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
            const __nv_bfloat162 hh = __floats2bfloat162_rn(a0, a1);
            orow[j] = *reinterpret_cast<const uint32_t*>(&hh);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TCOLS);
}

static bool conv3x3_tc_supported(const Conv3Params& p) {
  return ConvTcCfg(p.Cin, p.Cout).ok && !tc_disabled();
}

template <int CIN, int NOUT, int TPP>
static int launch_conv3x3_tc_n(const Conv3Params& p, cudaStream_t stream) {
  constexpr int KC = (CIN + 1 + 15) / 16 * 16;
  const size_t smem = (size_t)TPP * 128 * KC * 2 + (size_t)9 * NOUT * KC * 2;
  auto kern = conv3x3_tc_kernel<CIN, NOUT, TPP>;
  HRF_CUDA(ensure_smem((const void*)kern, smem));
  const int ntok = p.B * p.Ho * p.Wo;
  const ConvTcCfg c(p.Cin, p.Cout);
  HRF_CUDA(launch_pdl(kern, dim3(ceil_div(ntok, 128), c.NSPLIT), dim3(256), smem, stream, p));
  count_launch();
  HRF_CUDA(cudaGetLastError());
  return HRF_OK;
}

// p.Ho / p.Wo / FastDivs already set by conv3x3_prepare
static int launch_conv3x3_tc(const Conv3Params& p, cudaStream_t stream) {
  HRF_REQUIRE((reinterpret_cast<uintptr_t>(p.blob) & 15) == 0, HRF_EINVAL, "conv3x3_tc: blob must be 16-byte aligned");
  const ConvTcCfg c(p.Cin, p.Cout);
  if (p.Cin == 18) {
    switch (c.NOUT) {
      case 32: return launch_conv3x3_tc_n<18, 32, 9>(p, stream);
      case 48: return launch_conv3x3_tc_n<18, 48, 9>(p, stream);
      case 80: return launch_conv3x3_tc_n<18, 80, 9>(p, stream);
      case 144: return launch_conv3x3_tc_n<18, 144, 9>(p, stream);
    }
  }
  if (p.Cin == 36 && p.Cout == 72) return launch_conv3x3_tc_n<36, 80, 9>(p, stream);
  if (p.Cin == 72 && p.Cout == 144) return launch_conv3x3_tc_n<72, 80, 3>(p, stream);
  HRF_REQUIRE(false, HRF_EUNSUPPORTED, "conv3x3_tc: %d -> %d", p.Cin, p.Cout);
}

}  // namespace hrf
