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

constexpr int kConvTcCin = 18, kConvTcKC = 32;

// TC section of the conv3x3 blob (appended to PwLayout(9*Cin, Cout)): nine bf16 B tiles
// [KC/8][NOUT rows][8] (chunk-major, K = channel; row 18 of the centre tap = folded bias)
struct ConvTcLayout {
  int NOUT, o_w, total;   // floats
  __host__ __device__ ConvTcLayout(int cin, int cout) {
    const int base = PwLayout(9 * cin, cout).total;
    NOUT = round_up(cout, 16);
    o_w = round_up(base, 4);
    total = (cin == kConvTcCin && NOUT <= 256) ? o_w + 9 * NOUT * kConvTcKC / 2 : base;
  }
};

template <int NOUT>
__global__ void __launch_bounds__(256) conv3x3_tc_kernel(Conv3Params p) {
  using namespace umma;
  constexpr int KC = kConvTcKC, CIN = kConvTcCin;
  constexpr int TCOLS = NOUT <= 32 ? 32 : NOUT <= 64 ? 64 : NOUT <= 128 ? 128 : 256;
  constexpr int A_B = 128 * KC * 2;                 // one tap's A tile
  constexpr int W_B = NOUT * KC * 2;                // one tap's B tile
  extern __shared__ __align__(128) unsigned char sm[];   // [9 A tiles][9 B tiles]
  __shared__ __align__(8) uint64_t bar, wbar;
  __shared__ uint32_t tmem_base_s;
  unsigned char* sA = sm;
  unsigned char* sW = sm + 9 * A_B;

  pdl_launch_dependents();
  const int tid = threadIdx.x, warp = warp_idx_uniform(), lane = tid & 31;
  const ConvTcLayout L(CIN, p.Cout);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_init(&wbar, 1);
    fence_mbar_init();
    mbar_expect_tx(&wbar, 9 * W_B);
    bulk_g2s(sW, p.blob + L.o_w, 9 * W_B, &wbar);
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, TCOLS);
  // chunk 3 (columns 24..31) of every tap is zero padding
  for (int e = tid; e < 9 * 128; e += 256)
    *reinterpret_cast<uint4*>(sA + (e >> 7) * A_B + 3 * 2048 + (e & 127) * 16) = make_uint4(0, 0, 0, 0);
  __syncthreads();                                  // barriers initialised, TMEM address published
  pdl_wait();

  // ---- gather: lane pair (2r, 2r+1) owns token r; half 0 = channels 0..7, half 1 = 8..17 ----
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(p.x);
  const int ntok = p.B * p.Ho * p.Wo;
  const int r = tid >> 1, half = tid & 1;
  const int t = blockIdx.x * 128 + r;
  const bool tv = t < ntok;
  int bb, rem, oy, ox;
  p.d_howo.divmod(tv ? t : 0, bb, rem);
  p.d_wo.divmod(rem, oy, ox);
  uint32_t w[9][5];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int iy = oy * p.stride - 1 + tap / 3, ix = ox * p.stride - 1 + tap % 3;
    const bool in = tv && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(
        x + ((size_t)(bb * p.H + (in ? iy : 0)) * p.W + (in ? ix : 0)) * CIN) + half * 4;
#pragma unroll
    for (int j = 0; j < 5; ++j) w[tap][j] = (in && (j < 4 || half)) ? __ldg(src + j) : 0u;
  }
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    unsigned char* a = sA + tap * A_B + r * 16;
    if (half == 0) {
      *reinterpret_cast<uint4*>(a) = make_uint4(w[tap][0], w[tap][1], w[tap][2], w[tap][3]);
    } else {
      *reinterpret_cast<uint4*>(a + 2048) = make_uint4(w[tap][0], w[tap][1], w[tap][2], w[tap][3]);
      const uint32_t one = (tap == 4 && tv) ? 0x00003F80u : 0u;     // channel 18 = 1.0
      *reinterpret_cast<uint4*>(a + 2 * 2048) = make_uint4(w[tap][4], one, 0u, 0u);
    }
  }
  mbar_wait(&wbar, 0);                              // weights have landed
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 0 && elect_one()) {
    constexpr uint32_t id = idesc_bf16(128, NOUT, false, false);
    const uint32_t a_a = smem_u32(sA), a_w = smem_u32(sW);
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
      for (int s = 0; s < KC / 16; ++s)
        mma_bf16(tmem, desc_kmajor(a_a + tap * A_B, 128, s), desc_kmajor(a_w + tap * W_B, NOUT, s), id,
                 tap > 0 || s > 0);
    mma_commit(&bar);
  }
  cta_wait(&bar, 0);
  tc_fence_after();

  // ---- epilogue: ReLU, bf16 -> global.  Warp w: TMEM quadrant w % 4, column group w / 4 ------
  {
    const int q = warp & 3, gq = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    const int to = blockIdx.x * 128 + row;
    __nv_bfloat16* out = static_cast<__nv_bfloat16*>(p.out);
    const int ncc = (p.Cout + 7) / 8;
#pragma unroll 1
    for (int cc = gq; cc < NOUT / 8; cc += 2) {      // warp-uniform trip count
      float y[8];
      tmem_ld8(trow + cc * 8, y);
      tmem_ld_wait();
      if (to < ntok && cc < ncc) {
        uint32_t* orow = reinterpret_cast<uint32_t*>(out + (size_t)to * p.Cout + cc * 8);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (cc * 8 + 2 * j < p.Cout) {
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
  const int n = round_up(p.Cout, 16);
  return p.Cin == kConvTcCin && p.Cout % 2 == 0 && (n == 32 || n == 48 || n == 80 || n == 144) &&
         !tc_disabled();
}

template <int NOUT>
static int launch_conv3x3_tc_n(const Conv3Params& p, cudaStream_t stream) {
  const size_t smem = 9 * (128 * kConvTcKC * 2 + NOUT * kConvTcKC * 2);
  auto kern = conv3x3_tc_kernel<NOUT>;
  HRF_CUDA(ensure_smem((const void*)kern, smem));
  const int ntok = p.B * p.Ho * p.Wo;
  HRF_CUDA(launch_pdl(kern, dim3(ceil_div(ntok, 128)), dim3(256), smem, stream, p));
  count_launch();
  HRF_CUDA(cudaGetLastError());
  return HRF_OK;
}

// p.Ho / p.Wo / FastDivs already set by launch_conv3x3
static int launch_conv3x3_tc(const Conv3Params& p, cudaStream_t stream) {
  HRF_REQUIRE((reinterpret_cast<uintptr_t>(p.blob) & 15) == 0, HRF_EINVAL, "conv3x3_tc: blob must be 16-byte aligned");
  switch (round_up(p.Cout, 16)) {
    case 32: return launch_conv3x3_tc_n<32>(p, stream);
    case 48: return launch_conv3x3_tc_n<48>(p, stream);
    case 80: return launch_conv3x3_tc_n<80>(p, stream);
    case 144: return launch_conv3x3_tc_n<144>(p, stream);
  }
  HRF_REQUIRE(false, HRF_EUNSUPPORTED, "conv3x3_tc: Cout=%d", p.Cout);
}

}  // namespace hrf
