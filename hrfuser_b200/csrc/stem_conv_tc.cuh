// Stem convolution on the tcgen05 tensor cores -- bf16 mode.
//
// conv 3x3, stride 2, pad 1, Cin <= 3 (camera 3, lidar 3, radar 2, gated 1 channels;
// reference hrnet.py:341-348 `conv1`, hrfuser_hrformer_based.py:380-387 `conv_a`)
// + folded BatchNorm + ReLU, reading the caller's fp32 NCHW image directly and writing
// bf16 channels-last: the dtype / layout conversion passes and cuDNN's channel padding
// of a 3-channel input disappear.  Implicit GEMM: M = 128 output pixels per tile,
// K = 9*Cin <= 27 (padded to 32), N = Cout; the im2col rows are gathered by one thread
// per pixel straight into the UMMA operand tile.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace hrf {

struct StemLayout {
  int Cin, Cout, o_bias, o_w, total;      // floats; W is a bf16 tile [32/8][Cout][8]
  __host__ __device__ StemLayout(int cin, int cout) {
    Cin = cin; Cout = cout;
    o_bias = 0; o_w = round_up(cout, 4); total = o_w + cout * 32 / 2;
  }
};

struct StemParams {
  const float* x;      // (B, Cin, H, W) fp32
  const float* blob;
  __nv_bfloat16* out;  // (B, Ho, Wo, Cout) bf16
  int B, Cin, H, W, Ho, Wo, Cout, relu;
  FastDiv d_wo, d_ho;  // filled by launch_stem_conv
};

template <int COUT>
__global__ void __launch_bounds__(128, 6) stem_conv_tc_kernel(StemParams p) {
  using namespace umma;
  static_assert(COUT % 16 == 0 && COUT <= 256, "Cout");
  constexpr int TCOLS = COUT <= 32 ? 32 : COUT <= 64 ? 64 : COUT <= 128 ? 128 : 256;
  __shared__ __align__(128) unsigned char sA[128 * 32 * 2];
  __shared__ __align__(128) unsigned char sW[COUT * 32 * 2];
  __shared__ float sBias[COUT];
  __shared__ __align__(16) unsigned char sOut[128 * (2 * COUT + 16)];   // epilogue staging
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  HRF_PROF_DECL
  const int tid = threadIdx.x, warp = warp_idx_uniform();
  const StemLayout L(p.Cin, COUT);
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.blob + L.o_w);
    uint4* dst = reinterpret_cast<uint4*>(sW);
    for (int e = tid; e < COUT * 32 * 2 / 16; e += 128) dst[e] = __ldg(src + e);
    for (int e = tid; e < COUT; e += 128) sBias[e] = __ldg(p.blob + L.o_bias + e);
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, TCOLS);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
  const uint32_t a_a = smem_u32(sA), a_w = smem_u32(sW);
  uint32_t phase = 0;
  const int n_pix = p.B * p.Ho * p.Wo;
  const int n_tiles = ceil_div(n_pix, 128);
  const size_t plane = (size_t)p.H * p.W;

  // this pixel's 3x3xCin patch (k = (ci*3 + ky)*3 + kx), zero outside the image / the tensor
  auto gather = [&](int pix, float* v) {
#pragma unroll
    for (int k = 0; k < 27; ++k) v[k] = 0.f;
    if (pix < n_pix) {
      int ox, t, oy, b;
      p.d_wo.divmod(pix, t, ox);
      p.d_ho.divmod(t, b, oy);
      const float* xb = p.x + (size_t)b * p.Cin * plane;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        if (ci < p.Cin) {
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const int iy = oy * 2 - 1 + ky;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const int ix = ox * 2 - 1 + kx;
              if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W)
                v[(ci * 3 + ky) * 3 + kx] = __ldg(xb + ci * plane + (size_t)iy * p.W + ix);
            }
          }
        }
      }
    }
  };
  // software pipeline: the next tile's patch is requested before this tile's MMA / epilogue
  float vn[27];
  gather(blockIdx.x * 128 + tid, vn);

  HRF_PROF(14)                                     // setup + first gather
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    HRF_PROF_TILE
    float v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = k < 27 ? vn[k] : 0.f;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) st_chunk(sA, tid, ch, 128, v + 8 * ch);
    HRF_PROF(0)                                    // patch -> operand tile (waits for the loads)
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    HRF_PROF(1)
    if (warp == 0 && elect_one()) {
      tc_fence_after();
      constexpr uint32_t id = idesc_bf16(128, COUT, false, false);
      mma_bf16(tmem, desc_kmajor(a_a, 128, 0), desc_kmajor(a_w, COUT, 0), id, false);
      mma_bf16(tmem, desc_kmajor(a_a, 128, 1), desc_kmajor(a_w, COUT, 1), id, true);
      mma_commit(&bar);
    }
    HRF_PROF(2)                                    // MMA issue
    gather((tile + gridDim.x) * 128 + tid, vn);     // lands behind the MMA and the epilogue
    HRF_PROF(3)                                    // next tile's loads issued
    cta_wait(&bar, phase);
    phase ^= 1;
    tc_fence_after();
    HRF_PROF(4)                                    // MMA wait
    // ---- epilogue: + bias, ReLU, bf16.  Each thread owns one pixel row of 2*COUT bytes; the
    // warp's 32 rows are staged in shared memory (pitch 2*COUT + 16: conflict-free both ways)
    // and written out 512 contiguous bytes per store instruction (8 lanes x 16 B per row)
    // instead of 32 scattered 16-byte pieces.
    {
      constexpr int PITCH = 2 * COUT + 16;
      unsigned char* srow = sOut + (size_t)tid * PITCH;
#pragma unroll
      for (int c0 = 0; c0 < COUT; c0 += 32) {
        float y[32];
        tmem_ld32(trow + c0, y);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          y[j] += sBias[c0 + j];
          if (p.relu) y[j] = fmaxf(y[j], 0.f);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u;
          __nv_bfloat162 h0 = __floats2bfloat162_rn(y[8 * q], y[8 * q + 1]);
          __nv_bfloat162 h1 = __floats2bfloat162_rn(y[8 * q + 2], y[8 * q + 3]);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(y[8 * q + 4], y[8 * q + 5]);
          __nv_bfloat162 h3 = __floats2bfloat162_rn(y[8 * q + 6], y[8 * q + 7]);
          u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
          u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
          *reinterpret_cast<uint4*>(srow + (c0 + 8 * q) * 2) = u;
        }
      }
      __syncwarp();
      constexpr int PIECES = 2 * COUT / 16;          // 16-byte pieces per row (8 for COUT = 64)
      constexpr int RPI = 32 / PIECES;               // rows per store instruction
      const int lane = tid & 31, w0 = tid & ~31;
#pragma unroll
      for (int k = 0; k < 32 / RPI; ++k) {
        const int r = w0 + k * RPI + lane / PIECES, piece = lane % PIECES;
        const int px = tile * 128 + r;
        if (px < n_pix)
          *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(p.out) + (size_t)px * (2 * COUT) + piece * 16) =
              *reinterpret_cast<const uint4*>(sOut + (size_t)r * PITCH + piece * 16);
      }
      __syncwarp();
    }
    HRF_PROF(5)                                    // epilogue
    // the barrier before the next MMA orders these TMEM reads and the sA rewrite
    tc_fence_before();
    __syncthreads();
    HRF_PROF(6)
  }
  HRF_PROF_END
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TCOLS);
}

static int launch_stem_conv(const StemParams& p, cudaStream_t stream) {
  HRF_REQUIRE(p.Cin >= 1 && p.Cin <= 3, HRF_EUNSUPPORTED, "stem_conv: Cin=%d (1..3)", p.Cin);
  HRF_REQUIRE(p.Cout == 64, HRF_EUNSUPPORTED, "stem_conv: Cout=%d (64)", p.Cout);
  const int n_tiles = ceil_div(p.B * p.Ho * p.Wo, 128);
  // persistent: exactly the CTAs that are resident at once (6 per SM by registers); a larger
  // grid runs as a full wave plus a mostly idle second one
  const int grid = n_tiles < 148 * 6 ? n_tiles : 148 * 6;
  HRF_CUDA(ensure_smem((const void*)stem_conv_tc_kernel<64>, 0));
  StemParams q = p;
  q.d_wo = FastDiv(p.Wo);
  q.d_ho = FastDiv(p.Ho);
  stem_conv_tc_kernel<64><<<grid, 128, 0, stream>>>(q);
  count_launch();
  HRF_CUDA(cudaGetLastError());
  return HRF_OK;
}

}  // namespace hrf
