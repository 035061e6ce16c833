// Dense 1x1 / 3x3 convolution (+ folded BN, + residual, + ReLU) of the stems, the Bottlenecks,
// the 256-channel transitions and the HRFPN convs as a warp-specialised, TMA-fed tcgen05
// implicit GEMM -- bf16 mode (reference hrnet.py:341-371,419-463,
// resnet.py:263-302, necks/hrfpn.py:87-100).
//
//   out[b,h,w,n] = act( sum_tap sum_c x[b, h*s+dy-1, w*s+dx-1, c] * W[tap][n][c] + bias[n] (+ resid[b,h,w,n]) )
//
//   M tile   128 output tokens = 8 rows x 16 columns of one image
//   N        all output channels of the layer, padded to 32 | 64 | 128 | 256 accumulator columns
//   K loop   taps x (Cin / 64): per step one A tile [128 tokens x 64 channels] and one B tile
//            [N x 64] in 128-byte-swizzled K-major shared memory, four K = 16 UMMAs
//
// Roles (192 threads -- 320 for 128 / 256 accumulator columns: a second set of epilogue warps, see cg::epi_groups --
// one CTA per SM, persistent over the tiles):
//   warp 0      TMA producer: for every K step one 4-D tensor-tile copy of the activations --
//               box (64 channels, 16, 8, 1) at (c0, w0 + dx - 1, h0 + dy - 1, b); the unit
//               zero-fills what lies outside the image, which IS the convolution's zero padding;
//               a stride-2 convolution reads through a map with element strides (1, 2, 2, 1) --
//               and one 3-D copy of the weights (64, N, 1) at (c0, 0, tap), into a ring of
//               STAGES stages (full / empty mbarriers)
//   warp 1      MMA issuer: waits for a stage, issues the four UMMAs, commits the stage's
//               `empty` barrier; after the last K step commits the accumulator's `full` barrier.
//               The accumulator is double-buffered in TMEM (2 x N columns): the MMAs of tile
//               i + 1 run while the epilogue drains tile i
//   warps 2-5   epilogue: TMEM -> registers (a thread per token row), + bias, + residual, ReLU,
//               bf16.  Channel counts that are multiples of 64 go through shared-memory "C
//               slots" of [128 tokens x 64 channels] in the same 128-byte swizzle: the producer
//               TMA-loads the residual sub-tile into the slot ahead of time, the epilogue adds
//               in place (conflict-free 16-byte accesses) and one elected thread TMA-stores the
//               slot; per-thread global accesses (a 512-byte stride between lanes) ran the
//               64 -> 256 + residual layer at 97 us against 36 us for cuDNN.  Other widths (the
//               18 / 36-channel transitions) store straight from registers.
// No CTA-wide barrier inside the tile loop.
//
// Row-reuse mode (RR: 3x3, stride 1, weights that fit shared memory -- Bottleneck conv2, the
// 256 -> 18 transition).  The plain implicit GEMM reads every activation tile nine times and
// the layer's weights once per tile from L2, and those layers ran at L2 -> SM bandwidth, 5x
// off their MMA time (profiles/r02_timeline.txt: 187 us for the 256 -> 18 transition).  Here
//   * the weights of the layer [taps][Cin/64][NPAD x 64] are loaded ONCE per CTA (and again when
//     its tile range crosses into the next problem of a grouped launch);
//   * a K step is (dx, 64-channel chunk): ONE box of TM_H + 2 = 10 rows x 16 tokens feeds the
//     three vertical taps -- the A tile of tap dy starts dy rows (dy x 2 KB, a whole number of
//     swizzle atoms) into the box -- so a tile costs 3 x 10 / 8 = 3.75 activation tile loads
//     per channel chunk instead of 9;
//   * CTAs walk contiguous tile ranges (neighbouring tiles share halo rows in L2).
//
// Halo mode (HALO, the default for those layers; HRF_CONV_HALO=0 falls back to RR).  The M tile is
// 16 rows x 8 tokens and a K step is one 64-channel chunk: ONE box of 18 rows x 10 tokens feeds all
// nine taps.  Tap (dy, dx) is that box read through a descriptor that starts (dy * 10 + dx) * 128
// bytes into it with an 8-row-group stride of 10 * 128 = 1 280 bytes (the box's row pitch).  The
// tensor core, like TMA, applies the 128-byte-swizzle XOR to absolute shared-memory address bits,
// so starts and strides that are not multiples of the 1 024-byte atom read the right rows
// (tools/probes/swz_probe.cu).  1.44 activation tile loads per channel chunk instead of 3.75.
//
// Concatenated K (kc_split, 1x1 only): K chunks past kc_split come from a second activation tensor
// -- Bottleneck conv3 and its downsample conv as one GEMM (hrf_convgemm_grouped_cat_fwd).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "tmap.cuh"
#include "umma.cuh"

namespace hrf {

// blob: fp32 bias[NPAD] | bf16 W [taps][NPAD rows][Cin] (K-major rows, zero rows beyond Cout)
struct ConvGemmLayout {
  int NPAD, o_bias, o_w, total;   // floats
  __host__ __device__ ConvGemmLayout(int cin, int cout, int taps) {
    NPAD = cout <= 32 ? 32 : cout <= 64 ? 64 : cout <= 128 ? 128 : 256;
    o_bias = 0;
    o_w = NPAD;
    total = o_w + taps * NPAD * cin / 2;
  }
};

// Up to kMaxProb problems of ONE shape per launch (the camera stream and the modality streams
// run the same stem / Bottleneck / transition layers on tensors of the same size, each with its
// own weights): the persistent CTAs walk the tiles of all of them, so the layer is one launch
// for (1 + M) streams instead of (1 + M) launches that each take the whole GPU in turn.
constexpr int kMaxProb = 4;
struct ConvGemmParams {
  const float* blob[kMaxProb];
  const void* resid[kMaxProb];      // [B][Ho][Wo][Cout] bf16 or nullptr (all or none)
  void* out[kMaxProb];              // [B][Ho][Wo][Cout] bf16
  int n_prob;
  int B, Ho, Wo, Cin, Cout, taps, relu, stride;   // relu: bit q = ReLU on problem q
  int tiles_w, tiles_h, n_tiles;    // n_tiles: per problem
  int n_stages, tiles_per_cta;      // RR mode: ring depth (runtime), contiguous tile range per CTA
  int kc_split;                     // > 0: K chunks >= kc_split come from a SECOND activation tensor
                                    // (maps.r), i.e. the conv of the concatenation [x | x2] (1x1 only)
  FastDiv d_tiles_prob, d_tiles_img, d_tiles_w;
};
struct ConvGemmMaps {               // 4 x 4 tensor maps = 2 KB of kernel parameters
  CUtensorMap x[kMaxProb], w[kMaxProb], c[kMaxProb], r[kMaxProb];
};

namespace cg {
constexpr int TM_H = 8, TM_W = 16;                    // the 128-token M tile
constexpr int A_BYTES = 128 * 128;                    // 128 tokens x 64 channels bf16
constexpr int A_RR_BYTES = (TM_H + 2) * TM_W * 128;   // row-reuse box: 10 rows x 16 tokens x 64 channels
constexpr int MAXST = 8;
// warp 0 = TMA producer, warp 1 = MMA issuer, then `groups` sets of four epilogue warps (one warp
// per TMEM lane quadrant).  Outputs of 128 / 256 accumulator columns get two sets that take
// alternate 64-channel sub-tiles: one warp per scheduler has nothing to hide the TMEM-load, shared-
// memory and barrier latencies of the epilogue behind, and at 256 columns the epilogue -- not HBM
// -- bounded the kernel (8 100 cycles per tile against 4 300 at the HBM rate).
__host__ __device__ constexpr int epi_groups(int npad) { return npad >= 128 ? 2 : 1; }
__host__ __device__ constexpr int nt(int npad) { return 64 + 128 * epi_groups(npad); }
// halo mode: the M tile is 16 rows x 8 tokens; ONE box of 18 rows x 10 tokens x 64 channels per K
// chunk feeds all nine taps (tap (dy, dx) starts (dy * 10 + dx) tokens into the box)
constexpr int HALO_TH = 16, HALO_TW = 8;
constexpr int HALO_PITCH = (HALO_TW + 2) * 128;                // bytes between image rows of the box
constexpr int HALO_BOX = (HALO_TH + 2) * HALO_PITCH;           // 23 040 bytes landed per box
constexpr int HALO_STAGE = (HALO_BOX + 1023) / 1024 * 1024;    // ring stride: stages stay 1024-byte aligned

// shared-memory descriptor of a 128-byte-swizzled K-major operand tile (rows of 128 bytes, 8-row
// groups 1024 bytes apart), K step `ks` (16 elements = 32 bytes) -- sm_100 descriptor version 1,
// layout type 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, int ks) {
  uint64_t d = 0;
  d |= (uint64_t)(((saddr + (uint32_t)ks * 32u) & 0x3FFFF) >> 4);
  d |= (uint64_t)(1024u >> 4) << 32;                  // stride byte offset: next 8-row group
  d |= (uint64_t)1 << 46;                             // descriptor version
  d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
  return d;
}

// the same with a start address that is any multiple of 128 bytes and any 8-row-group stride:
// the hardware applies the 128-byte-swizzle XOR to ABSOLUTE shared-memory address bits (7..9 into
// 4..6), exactly as TMA wrote the box, so a row-granular shift and a 1 280-byte group stride read
// the right rows with base_offset = 0 (tools/probes/swz_probe.cu, measured on B200)
__device__ __forceinline__ uint64_t desc_sw128_sbo(uint32_t saddr, int ks, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)(((saddr + (uint32_t)ks * 32u) & 0x3FFFF) >> 4);
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void tma_load_4d(void* dst_smem, const void* tmap, int c0, int c1, int c2, int c3,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(umma::smem_u32(dst_smem)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(umma::smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}
}  // namespace cg

template <int NPAD>
struct ConvGemmCfg {
  static constexpr int B_BYTES = NPAD * 128;
  static constexpr int STAGE = cg::A_BYTES + B_BYTES;
  static constexpr int NSUB = NPAD / 64;                      // 64-channel sub-tiles of the output (0: direct stores)
  // C slots (residual in, result out): at least two per epilogue group -- a group frees a slot
  // when it issues its NEXT store, so with one slot it would wait for itself
  static constexpr int NSLOT = NSUB == 0 ? 0 : (NSUB <= 2 ? 2 * NSUB : NSUB);
  static constexpr int C_BYTES = NSLOT * cg::A_BYTES;
  static constexpr int ROOM = 218 * 1024 - C_BYTES;     // 227 KB - static (bias table, barriers) - slack
  static constexpr int STAGES = ROOM / STAGE > 8 ? 8 : ROOM / STAGE;
  static constexpr int SMEM = STAGES * STAGE + C_BYTES + 1024;          // + alignment slack
  static constexpr int TMEM_COLS = 2 * NPAD < 32 ? 32 : 2 * NPAD;
  static_assert(SMEM + kMaxProb * NPAD * 4 + 1024 <= 227 * 1024, "dynamic + static shared memory");
};

// MODE 0: plain (activation tile + weight tile per K step); 1: row reuse (RR); 2: halo box (HALO).
// RR and HALO keep the layer's weights resident and walk contiguous tile ranges.
template <int NPAD, int MODE>
__global__ void __launch_bounds__(cg::nt(NPAD), 1)
conv_gemm_tc_kernel(const __grid_constant__ ConvGemmParams p, const __grid_constant__ ConvGemmMaps tm) {
  using namespace umma;
  using K = ConvGemmCfg<NPAD>;
  constexpr bool RR = MODE == 1, HALO = MODE == 2, RES = MODE != 0;
  constexpr int EG = cg::epi_groups(NPAD);
  constexpr int TH = HALO ? cg::HALO_TH : cg::TM_H, TW = HALO ? cg::HALO_TW : cg::TM_W;   // the 128-token M tile
  const int STAGES = RES ? p.n_stages : K::STAGES;
  constexpr int STAGE_B = RR ? cg::A_RR_BYTES : HALO ? cg::HALO_STAGE : K::STAGE;
  extern __shared__ unsigned char sm_raw[];
  constexpr int NSLOT = K::NSLOT > 0 ? K::NSLOT : 1, NSUB = K::NSUB;
  __shared__ __align__(8) uint64_t full[cg::MAXST], empty[cg::MAXST], acc_full[2], acc_empty[2], w_full;
  __shared__ __align__(8) uint64_t c_full[NSLOT], c_empty[NSLOT];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_bias[kMaxProb * NPAD];

  pdl_launch_dependents();
  const int tid = threadIdx.x, warp = warp_idx_uniform(), lane = tid & 31;
  // operand tiles need 1024-byte alignment (128-byte swizzle atom = 8 rows x 128 bytes)
  unsigned char* sm = sm_raw + ((1024u - (smem_u32(sm_raw) & 1023u)) & 1023u);
  const int kc = p.Cin / 64, n_k = RR ? 3 * kc : HALO ? kc : p.taps * kc;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 4 * EG);               // one arrival per epilogue warp
    }
    for (int c = 0; c < NSLOT; ++c) {
      mbar_init(&c_full[c], 1);
      mbar_init(&c_empty[c], 1);
    }
    mbar_init(&w_full, 1);
    fence_mbar_init();
    for (int q = 0; q < p.n_prob; ++q) {
      tma_prefetch_desc(&tm.x[q]);
      tma_prefetch_desc(&tm.w[q]);
      if (NSUB > 0) tma_prefetch_desc(&tm.c[q]);
    }
  }
  const int total_tiles = p.n_tiles * p.n_prob;
  const bool staged = NSUB > 0 && p.Cout == NPAD;    // whole 64-channel sub-tiles: the C-slot epilogue
  unsigned char* c_slots = sm + STAGES * STAGE_B;
  unsigned char* w_res = c_slots + K::C_BYTES;       // RR: the layer's weights, resident
  // tiles of this CTA: strided over the grid, or (RR) one contiguous range
  const int t_begin = RES ? blockIdx.x * p.tiles_per_cta : blockIdx.x;
  const int t_step = RES ? 1 : gridDim.x;
  const int t_end = RES ? (t_begin + p.tiles_per_cta < total_tiles ? t_begin + p.tiles_per_cta : total_tiles) : total_tiles;
  for (int e = tid; e < p.n_prob * NPAD; e += cg::nt(NPAD)) s_bias[e] = __ldg(p.blob[e / NPAD] + (e % NPAD));
  if (warp == 1) tmem_alloc(&tmem_base_s, K::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  pdl_wait();

  if (warp == 0) {
    // ===== TMA producer ===========================================================================
    if (elect_one()) {
      int it = 0, cit = 0, cur_pr = -1;
      for (int tile = t_begin; tile < t_end; tile += t_step) {
        int pr, tl, b, rem, ty, tx;
        p.d_tiles_prob.divmod(tile, pr, tl);
        p.d_tiles_img.divmod(tl, b, rem);
        p.d_tiles_w.divmod(rem, ty, tx);
        const int h0 = ty * TH, w0 = tx * TW;
        if constexpr (RES) {
          if (pr != cur_pr) {
            // (re)load the resident weights; the MMAs of the previous problem's tiles must have
            // read the old ones: the commit behind the last stage issued covers all of them
            if (it > 0) mbar_wait(&empty[(it - 1) % STAGES], ((it - 1) / STAGES) & 1, 270);
            mbar_expect_tx(&w_full, (uint32_t)(9 * kc * K::B_BYTES));
            for (int t = 0; t < 9 * kc; ++t)
              tma_load_3d(w_res + t * K::B_BYTES, &tm.w[pr], (t % kc) * 64, 0, t / kc, &w_full);
            cur_pr = pr;
          }
          for (int k = 0; k < n_k; ++k, ++it) {
            const int s = it % STAGES;
            if (it >= STAGES) mbar_wait(&empty[s], ((it / STAGES) - 1) & 1, 200 + s);
            if constexpr (HALO) {                   // k = channel chunk: the whole 18 x 10 halo box
              mbar_expect_tx(&full[s], cg::HALO_BOX);
              cg::tma_load_4d(sm + s * STAGE_B, &tm.x[pr], k * 64, w0 - 1, h0 - 1, b, &full[s]);
            } else {
              const int dxi = k / kc, c0 = (k - dxi * kc) * 64;
              mbar_expect_tx(&full[s], cg::A_RR_BYTES);
              cg::tma_load_4d(sm + s * STAGE_B, &tm.x[pr], c0, w0 + dxi - 1, h0 - 1, b, &full[s]);
            }
          }
        } else {
        for (int k = 0; k < n_k; ++k, ++it) {
          const int s = it % STAGES;
          if (it >= STAGES) mbar_wait(&empty[s], ((it / STAGES) - 1) & 1, 200 + s);
          const int tap = k / kc, c0 = (k - tap * kc) * 64;
          const int dy = p.taps == 9 ? tap / 3 - 1 : 0, dx = p.taps == 9 ? tap % 3 - 1 : 0;
          unsigned char* a = sm + s * STAGE_B;
          mbar_expect_tx(&full[s], cg::A_BYTES + K::B_BYTES);
          const bool second = p.kc_split > 0 && k >= p.kc_split;      // taps == 1: k is the chunk
          cg::tma_load_4d(a, second ? &tm.r[pr] : &tm.x[pr], second ? c0 - p.kc_split * 64 : c0,
                          w0 * p.stride + dx, h0 * p.stride + dy, b, &full[s]);
          tma_load_3d(a + cg::A_BYTES, &tm.w[pr], c0, 0, tap, &full[s]);
        }
        }
        if (NSUB > 0 && staged && p.resid[0] != nullptr) {
          // residual sub-tiles of THIS tile into the C slots (after its K steps: the slots are
          // released by the previous tile's stores, which must not hold back this tile's MMAs)
          for (int j = 0; j < NSUB; ++j, ++cit) {
            const int cs = cit % NSLOT;
            if (cit >= NSLOT) mbar_wait(&c_empty[cs], ((cit / NSLOT) - 1) & 1, 240 + cs);
            mbar_expect_tx(&c_full[cs], cg::A_BYTES);
            cg::tma_load_4d(c_slots + cs * cg::A_BYTES, &tm.r[pr], j * 64, w0, h0, b, &c_full[cs]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer ==============================================================================
    constexpr uint32_t idesc = idesc_bf16(128, NPAD, false, false);
    int it = 0, tcount = 0, cur_pr = -1, wcount = 0;
    for (int tile = t_begin; tile < t_end; tile += t_step, ++tcount) {
      const int acc = tcount & 1;
      if constexpr (RES) {
        const int pr = p.d_tiles_prob.div(tile);
        if (pr != cur_pr) {                           // this problem's weights have landed
          mbar_wait(&w_full, wcount & 1, 280);
          ++wcount;
          cur_pr = pr;
        }
      }
      if (tcount >= 2) mbar_wait(&acc_empty[acc], ((tcount >> 1) - 1) & 1, 210 + acc);
      tc_fence_after();
      for (int k = 0; k < n_k; ++k, ++it) {
        const int s = it % STAGES;
        mbar_wait(&full[s], (it / STAGES) & 1, 220 + s);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a = smem_u32(sm + s * STAGE_B);
          if constexpr (HALO) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const uint32_t at = a + (uint32_t)((tap / 3) * cg::HALO_PITCH + (tap % 3) * 128);
              const uint32_t bq = smem_u32(w_res) + (uint32_t)((tap * kc + k) * K::B_BYTES);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                mma_bf16(tmem + acc * NPAD, cg::desc_sw128_sbo(at, ks, cg::HALO_PITCH), cg::desc_sw128(bq, ks), idesc,
                         (k | tap | ks) != 0);
            }
          } else if constexpr (RR) {
            const int dxi = k / kc, c = k - dxi * kc;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              const uint32_t bq = smem_u32(w_res) + (uint32_t)(((dy * 3 + dxi) * kc + c) * K::B_BYTES);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                mma_bf16(tmem + acc * NPAD, cg::desc_sw128(a + dy * (cg::TM_W * 128), ks), cg::desc_sw128(bq, ks), idesc,
                         (k | dy | ks) != 0);
            }
          } else {
            const uint32_t bq = a + cg::A_BYTES;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              mma_bf16(tmem + acc * NPAD, cg::desc_sw128(a, ks), cg::desc_sw128(bq, ks), idesc, (k | ks) != 0);
          }
          mma_commit(&empty[s]);                      // frees the stage when these MMAs have read it
          if (k == n_k - 1) mma_commit(&acc_full[acc]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue warps ==========================================================================
    const int q = warp & 3;                           // TMEM lane quadrant of this warp
    const int eg = (warp - 2) >> 2;                   // epilogue group: sub-tiles j with j % EG == eg
    const int row = q * 32 + lane;                    // token of the tile
    const int hh = row / TW, ww = row % TW;
    int tcount = 0, cit = 0, prev_cit = -1;           // prev_cit: the last slot THIS group stored
    const bool solo = EG > 1 && NSUB > 0 && staged && p.resid[0] != nullptr;
    for (int tile = t_begin; tile < t_end; tile += t_step, ++tcount) {
      const int acc = tcount & 1;
      int pr, tl, b, rem, ty, tx;
      p.d_tiles_prob.divmod(tile, pr, tl);
      p.d_tiles_img.divmod(tl, b, rem);
      p.d_tiles_w.divmod(rem, ty, tx);
      const __nv_bfloat16* resid = static_cast<const __nv_bfloat16*>(p.resid[pr]);
      __nv_bfloat16* out = static_cast<__nv_bfloat16*>(p.out[pr]);
      const float* bias = s_bias + pr * NPAD;
      const bool relu = (p.relu >> pr) & 1;
      const int h = ty * TH + hh, w = tx * TW + ww;
      const bool live = h < p.Ho && w < p.Wo;
      const size_t tok = ((size_t)(b * p.Ho + (live ? h : 0)) * p.Wo + (live ? w : 0)) * p.Cout;
      mbar_wait(&acc_full[acc], (tcount >> 1) & 1, 230 + acc);
      tc_fence_after();
      const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + acc * NPAD;
      if (NSUB > 0 && staged) {
        // ---- C-slot epilogue: 64-channel sub-tiles through swizzled shared memory + TMA store ----
        const bool has_r = resid != nullptr;
#pragma unroll 1
        for (int j = 0; j < NSUB; ++j, ++cit) {
          // the other group's sub-tile.  With a residual, group 0 takes them all: the producer
          // reloads a slot when its store has been issued AND the same group has issued the next
          // one, and with alternating groups that hand-over put two exposed HBM latencies into
          // every tile (conv3 + identity 28 -> 37 us); without one, two groups: 23.5 -> 15 us
          if (EG > 1 && (solo ? eg != 0 : (j & (EG - 1)) != eg)) continue;
          const int cs = cit % NSLOT;
          unsigned char* slot = c_slots + cs * cg::A_BYTES;
          float v[64];
          tmem_ld32(trow + j * 64, v);
          tmem_ld32(trow + j * 64 + 32, v + 32);
          if (has_r) mbar_wait(&c_full[cs], (cit / NSLOT) & 1, 250 + cs);          // residual has landed
          else if (cit >= NSLOT) mbar_wait(&c_empty[cs], ((cit / NSLOT) - 1) & 1, 260 + cs);   // slot drained
          tmem_ld_wait();
          unsigned char* rowp = slot + row * 128;
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            uint4* cp = reinterpret_cast<uint4*>(rowp + ((ch ^ (row & 7)) << 4));
            float a[8];
            const float4 b0 = *reinterpret_cast<const float4*>(bias + j * 64 + ch * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(bias + j * 64 + ch * 8 + 4);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) a[e] = v[ch * 8 + e] + bb[e];
            if (has_r) {
              const uint4 r = *cp;
              const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                a[2 * e] += __uint_as_float(rw[e] << 16);
                a[2 * e + 1] += __uint_as_float(rw[e] & 0xffff0000u);
              }
            }
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float a0 = a[2 * e], a1 = a[2 * e + 1];
              if (relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
              const __nv_bfloat162 hb = __floats2bfloat162_rn(a0, a1);
              o[e] = *reinterpret_cast<const uint32_t*>(&hb);
            }
            *cp = make_uint4(o[0], o[1], o[2], o[3]);
          }
          fence_proxy_async();                                  // generic writes -> the TMA store's reads
          asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");   // the four warps of this group
          if (q == 2 && elect_one()) {                          // warp 2 / warp 6: first warp of the group
            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                         ::"l"(&tm.c[pr]), "r"(smem_u32(slot)), "r"(j * 64), "r"(tx * TW), "r"(ty * TH), "r"(b)
                         : "memory");
            tma_store_commit();
            if (prev_cit >= 0) {                                // this group's PREVIOUS store has read its slot
              asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
              cg::mbar_arrive(&c_empty[prev_cit % NSLOT]);
            }
          }
          prev_cit = cit;
        }
      } else {
#pragma unroll 1
      for (int c0 = 0; c0 < NPAD; c0 += 32) {
        if (EG > 1 && ((c0 >> 5) & (EG - 1)) != eg) continue;
        float v[32];
        tmem_ld32(trow + c0, v);
        tmem_ld_wait();
        if (live && c0 < p.Cout) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += bias[c0 + j];
          if (((p.Cout * 2) & 15) == 0) {             // 16-byte vector path (channel count % 8 == 0)
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
              if (c0 + g8 * 8 < p.Cout) {
                if (resid) {
                  const uint4 r = __ldg(reinterpret_cast<const uint4*>(resid + tok + c0 + g8 * 8));
                  const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    v[g8 * 8 + 2 * j] += __uint_as_float(rw[j] << 16);
                    v[g8 * 8 + 2 * j + 1] += __uint_as_float(rw[j] & 0xffff0000u);
                  }
                }
                uint32_t o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  float a0 = v[g8 * 8 + 2 * j], a1 = v[g8 * 8 + 2 * j + 1];
                  if (relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
                  const __nv_bfloat162 hb = __floats2bfloat162_rn(a0, a1);
                  o[j] = *reinterpret_cast<const uint32_t*>(&hb);
                }
                *reinterpret_cast<uint4*>(out + tok + c0 + g8 * 8) = make_uint4(o[0], o[1], o[2], o[3]);
              }
            }
          } else {                                    // 4-byte path (18 / 36-channel transitions)
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (c0 + 2 * j < p.Cout) {
                float a0 = v[2 * j], a1 = v[2 * j + 1];
                if (resid) {
                  const uint32_t r = __ldg(reinterpret_cast<const uint32_t*>(resid + tok + c0 + 2 * j));
                  a0 += __uint_as_float(r << 16);
                  a1 += __uint_as_float(r & 0xffff0000u);
                }
                if (relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
                const __nv_bfloat162 hb = __floats2bfloat162_rn(a0, a1);
                *reinterpret_cast<uint32_t*>(out + tok + c0 + 2 * j) = *reinterpret_cast<const uint32_t*>(&hb);
              }
            }
          }
        }
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) cg::mbar_arrive(&acc_empty[acc]);
    }
    if (NSUB > 0 && staged && q == 2 && elect_one()) tma_store_wait_all();      // same thread that issued them
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, K::TMEM_COLS);
}

// ---- host side ------------------------------------------------------------------------------------
static bool conv_gemm_supported(int Cin, int Cout, int k, int stride, int H, int W) {
  return (k == 1 || k == 3) && (stride == 1 || stride == 2) && Cin % 64 == 0 && Cin >= 64 && Cout % 2 == 0 &&
         Cout <= 256 && H > 0 && W > 0 && (k == 3 || stride == 1);
}

// HRF_CONV_RR=0 keeps the plain implicit GEMM for the 3x3 stride-1 layers (A/B runs)
static bool conv_rr_enabled() {
  const char* e = std::getenv("HRF_CONV_RR");
  return !(e && e[0] == '0');
}

// HRF_CONV_HALO=0 keeps the row-reuse mode for them (A/B runs)
static bool conv_halo_enabled() {
  const char* e = std::getenv("HRF_CONV_HALO");
  return !(e && e[0] == '0');
}

template <int NPAD>
static int launch_conv_gemm_n(ConvGemmParams p, const void* const* xs, const void* const* xs2, int cin1, int Hi,
                              int Wi, int stride, cudaStream_t stream) {
  using K = ConvGemmCfg<NPAD>;
  // row-reuse mode: 3x3, stride 1, the layer's weights + at least two ring stages fit
  const int wres = 9 * (p.Cin / 64) * K::B_BYTES;
  const int room = 218 * 1024 - K::C_BYTES - wres;
  // halo mode: the same layers, one 18 x 10 box per channel chunk for all nine taps (2.6x less
  // L2 -> shared-memory traffic than the three 10 x 16 boxes of the row-reuse mode)
  const bool halo = p.taps == 9 && stride == 1 && room >= 2 * cg::HALO_STAGE && conv_rr_enabled() && conv_halo_enabled();
  const bool rr = !halo && p.taps == 9 && stride == 1 && room >= 2 * cg::A_RR_BYTES && conv_rr_enabled();
  const int res_stage = halo ? cg::HALO_STAGE : cg::A_RR_BYTES;
  p.n_stages = (halo || rr) ? (room / res_stage > cg::MAXST ? cg::MAXST : room / res_stage) : K::STAGES;
  const int TH = halo ? cg::HALO_TH : cg::TM_H, TW = halo ? cg::HALO_TW : cg::TM_W;
  p.tiles_w = ceil_div(p.Wo, TW);
  p.tiles_h = ceil_div(p.Ho, TH);
  p.n_tiles = p.B * p.tiles_w * p.tiles_h;
  p.d_tiles_prob = FastDiv(p.n_tiles);
  p.d_tiles_img = FastDiv(p.tiles_w * p.tiles_h);
  p.d_tiles_w = FastDiv(p.tiles_w);
  PFN_tmapEncodeTiled enc = tmap_encoder();
  HRF_REQUIRE(enc != nullptr, HRF_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
  ConvGemmMaps tm;
  const bool staged = K::NSUB > 0 && p.Cout == NPAD;
  p.kc_split = xs2 ? cin1 / 64 : 0;
  // activations [B][Hi][Wi][C] bf16 as (C, Wi, Hi, B); a stride-2 convolution steps over every
  // second token / row (element strides), so its coordinates stay in INPUT units
  auto act_map = [&](CUtensorMap* m, const void* ptr, int Cx) -> CUresult {
    const cuuint64_t gdim[4] = {(cuuint64_t)Cx, (cuuint64_t)Wi, (cuuint64_t)Hi, (cuuint64_t)p.B};
    const cuuint64_t gstr[3] = {(cuuint64_t)Cx * 2, (cuuint64_t)Wi * Cx * 2, (cuuint64_t)Hi * Wi * Cx * 2};
    const cuuint32_t box[4] = {64u, (cuuint32_t)(halo ? TW + 2 : TW * stride),
                               (cuuint32_t)(halo || rr ? TH + 2 : TH * stride), 1u};
    const cuuint32_t est[4] = {1u, (cuuint32_t)stride, (cuuint32_t)stride, 1u};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, box, est,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  for (int q = 0; q < p.n_prob; ++q) {
    {
      const CUresult r = act_map(&tm.x[q], xs[q], xs2 ? cin1 : p.Cin);
      HRF_REQUIRE(r == CUDA_SUCCESS, HRF_ECUDA, "conv_gemm: activation tensor map failed (%d)", (int)r);
    }
    {
      const ConvGemmLayout L(p.Cin, p.Cout, p.taps);
      const cuuint64_t gdim[3] = {(cuuint64_t)p.Cin, (cuuint64_t)NPAD, (cuuint64_t)p.taps};
      const cuuint64_t gstr[2] = {(cuuint64_t)p.Cin * 2, (cuuint64_t)NPAD * p.Cin * 2};
      const cuuint32_t box[3] = {64u, (cuuint32_t)NPAD, 1u};
      const cuuint32_t est[3] = {1u, 1u, 1u};
      const CUresult r = enc(&tm.w[q], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<float*>(p.blob[q] + L.o_w), gdim,
                             gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      HRF_REQUIRE(r == CUDA_SUCCESS, HRF_ECUDA, "conv_gemm: weight tensor map failed (%d)", (int)r);
    }
    tm.c[q] = tm.x[q];                                  // placeholders when the direct epilogue runs
    tm.r[q] = tm.x[q];
    if (staged) {
      const cuuint64_t gdim[4] = {(cuuint64_t)p.Cout, (cuuint64_t)p.Wo, (cuuint64_t)p.Ho, (cuuint64_t)p.B};
      const cuuint64_t gstr[3] = {(cuuint64_t)p.Cout * 2, (cuuint64_t)p.Wo * p.Cout * 2, (cuuint64_t)p.Ho * p.Wo * p.Cout * 2};
      const cuuint32_t box[4] = {64u, (cuuint32_t)TW, (cuuint32_t)TH, 1u};
      const cuuint32_t est[4] = {1u, 1u, 1u, 1u};
      CUresult r = enc(&tm.c[q], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, p.out[q], gdim, gstr, box, est,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      HRF_REQUIRE(r == CUDA_SUCCESS, HRF_ECUDA, "conv_gemm: output tensor map failed (%d)", (int)r);
      if (p.resid[q]) {
        r = enc(&tm.r[q], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p.resid[q]), gdim, gstr, box, est,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        HRF_REQUIRE(r == CUDA_SUCCESS, HRF_ECUDA, "conv_gemm: residual tensor map failed (%d)", (int)r);
      }
    }
    if (xs2) {                                          // second K source rides in the residual's map slot
      const CUresult r = act_map(&tm.r[q], xs2[q], p.Cin - cin1);
      HRF_REQUIRE(r == CUDA_SUCCESS, HRF_ECUDA, "conv_gemm: second activation tensor map failed (%d)", (int)r);
    }
  }
  for (int q = p.n_prob; q < kMaxProb; ++q) { tm.x[q] = tm.x[0]; tm.w[q] = tm.w[0]; tm.c[q] = tm.c[0]; tm.r[q] = tm.r[0]; }
  const int total = p.n_tiles * p.n_prob;
  int grid = total < 148 ? total : 148;
  if (halo || rr) {
    p.tiles_per_cta = ceil_div(total, grid);
    grid = ceil_div(total, p.tiles_per_cta);
    const size_t smem = (size_t)p.n_stages * res_stage + K::C_BYTES + wres + 1024;
    if (halo) {
      HRF_CUDA(ensure_smem((const void*)conv_gemm_tc_kernel<NPAD, 2>, smem));
      HRF_CUDA(launch_pdl(conv_gemm_tc_kernel<NPAD, 2>, dim3(grid), dim3(cg::nt(NPAD)), smem, stream, p, tm));
    } else {
      HRF_CUDA(ensure_smem((const void*)conv_gemm_tc_kernel<NPAD, 1>, smem));
      HRF_CUDA(launch_pdl(conv_gemm_tc_kernel<NPAD, 1>, dim3(grid), dim3(cg::nt(NPAD)), smem, stream, p, tm));
    }
    count_launch();
    HRF_CUDA(cudaGetLastError());
    return HRF_OK;
  }
  p.tiles_per_cta = 0;
  HRF_CUDA(ensure_smem((const void*)conv_gemm_tc_kernel<NPAD, 0>, K::SMEM));
  HRF_CUDA(launch_pdl(conv_gemm_tc_kernel<NPAD, 0>, dim3(grid), dim3(cg::nt(NPAD)), K::SMEM, stream, p, tm));
  count_launch();
  HRF_CUDA(cudaGetLastError());
  return HRF_OK;
}

}  // namespace hrf
