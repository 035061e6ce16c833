// Fused window attention, third generation: C = 18, one head (the branch-0 blocks of HRFuser-T:
// 95 % of the attention tokens) -- LSA and the whole MWCA (every modality) in ONE launch.
// Reference: hrformer.py:96-131,184-236,369 (LSA), hrfuser_hrformer_based.py:106-151,189-248,
// 305-313 (MWCA).
//
// What changed against window_attn_tc.cuh (which stays for the multi-head widths):
//   * three tensor-core round trips per window pair instead of four, and 64 TMEM columns
//     instead of 128 (LSA), so 5 CTAs are resident per SM instead of 4 (the register budget: at 6,
//     80 registers, nvcc spills the prefetched row right behind its load and LayerNorm stalls on it):
//       - with ONE head the output projection commutes with the softmax average:
//           out_i = sum_j P_ij (v_j Wo^T) / sum_j P_ij + bo,
//         so V' = LN(z) (Wo Wv)^T is projected alongside Q and K (weights multiplied on the
//         host) and P V' IS the block's output: no O tile, no out-projection MMA;
//       - Q | K | V' come out of one N = 64 MMA with their 18 + 18 + 19 columns packed;
//       - S = Q K^T of the two windows of a pair is ONE M = 128, N = 64 MMA: the K dimension is
//         [window A features | window B features] (48 = 2 x 24), a query row carries zeros in
//         the other window's half, the key tile row j holds K_A[j] | K_B[j].  A row's 64 score
//         columns are the keys of its own window (the previous kernel computed 128 and threw
//         half away);
//   * LayerNorm's affine is folded into the projection weights (x^ = (x - mean) rstd goes to
//     the tensor cores); the reference pads with zeros AFTER the affine, so the tile carries
//     two constant columns: 18 = 1 on every row (x the bias row: bq, bk, Wo bv + bo) and
//     19 = 1 on real tokens only (x the beta row: Wq beta, ...): a padded slot projects to the
//     bias alone, exactly as in the reference;
//   * the softmax denominator comes out of the P V' MMA (V' column 18 is the constant 1), the
//     softmax scale and log2(e) are folded into Wq and the relative-position table: the
//     exponent is one FADD + EX2 per key, no row sum in registers;
//   * MWCA: the camera row is normalised once per tile, the modalities loop inside the tile
//     with their own Q / K / V' weights, the per-modality results accumulate in registers and
//     the block writes its output once (the previous version launched once per modality and
//     re-read the camera tokens and the accumulator every time).
#pragma once
#include "common.cuh"
#include "umma.cuh"
#include "window_attn.cuh"
#include "window_attn_tc.cuh"

namespace hrf {

struct AttnV3 {
  static constexpr int C = 18, KC = 32, S = 49, WIN = 7, MAXMOD = 4;
  static constexpr int NV = 19;                                 // V' columns: 18 features + the constant 1
  // blob sections (bytes): bf16 operand tiles (chunk-major, umma.cuh) + the scaled rpb table
  static constexpr int W_SELF_B = 64 * KC * 2;                  // rows: Wq 0..17 | Wk 18..35 | Wv' 36..54
  static constexpr int W_Q_B = 32 * KC * 2;                     // rows: Wq 0..17
  static constexpr int W_KV_B = 48 * KC * 2;                    // rows: Wk 0..17 | Wv' 18..36
  static constexpr int RPB_B = 172 * 4;                         // fp32 [169] * log2(e), padded
  static constexpr int SEC_SELF = W_SELF_B + RPB_B;             // 4784
  static constexpr int SEC_CROSS = W_Q_B + W_KV_B + RPB_B;      // 5808
  static_assert(SEC_SELF % 16 == 0 && SEC_CROSS % 16 == 0, "bulk-copy granularity");
  // shared-memory tiles
  static constexpr int XT = 128 * KC * 2;                       // [128 tokens x 32] operand tile, 8 KB
  static constexpr int QT = 6 * 128 * 16;                       // Q  [128 rows x 48], 12 KB
  static constexpr int KT = 6 * 64 * 16;                        // K  [64 key slots x 48], 6 KB
  static constexpr int REGION = QT + KT;                        // the P tile [128 x 64] (16 KB) aliases Q | K
  static_assert(REGION >= 128 * 64 * 2, "P tile");
  __host__ __device__ static constexpr int smem_bytes(bool cross, int n_mod) {
    return (cross ? 2 : 1) * XT + REGION + n_mod * (cross ? SEC_CROSS : SEC_SELF);
  }
  __host__ __device__ static constexpr int tmem_cols(bool cross) { return cross ? 128 : 64; }
  static bool applies(int C_, int heads, int win) { return C_ == 18 && heads == 1 && win == 7; }
};

// Self-attention may run several PROBLEMS of one shape in one launch (the camera's branch 0 and
// the modality streams walk the same HRFormer blocks on tensors of the same size, each with its
// own weights: grouped, they share the launch, the setup and the tail: 3 x 13.7 us -> ~27 us).
struct AttnV3Params {
  const void* x;                        // cross: camera / query tokens, also the residual
  const void* z[AttnV3::MAXMOD];        // cross: key / value tokens per modality
  const float* blob[AttnV3::MAXMOD];    // packed blob per modality (cross) / per problem (self)
  const void* xs[AttnV3::MAXMOD];       // self: tokens of every problem
  void* outs[AttnV3::MAXMOD];           // self: output of every problem
  int n_prob;                           // self: problems in this launch (cross: 1)
  int v3_off;                           // float offset of the v3 section inside a blob
  void* out;
  int n_mod;                            // 1 for self-attention
  int B, H, W, pad_mask;
  float eps;
  FastDiv d_win_img, d_win_row, d_tiles;   // d_tiles: tiles per problem (global tile -> problem)
};

namespace v3 {
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// x^ = (x - mean) * rstd of one 18-channel token (9 packed bf16 words) -> chunks 0..2 of row
// `row` of an R = 128 operand tile; column 18 = 1, column 19 = real.  Packed fp32 arithmetic.
__device__ __forceinline__ void ln18_to_tile(const uint32_t* w, bool real, float eps, unsigned char* tile, int row) {
  uint32_t o[12];
  if (real) {
    float2 v[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) v[j] = make_float2(__uint_as_float(w[j] << 16), __uint_as_float(w[j] & 0xffff0000u));
    float2 sa = __fadd2_rn(v[0], v[1]), sb = __fadd2_rn(v[2], v[3]), sc = __fadd2_rn(v[4], v[5]);
    sa = __fadd2_rn(sa, v[6]); sb = __fadd2_rn(sb, v[7]); sc = __fadd2_rn(sc, v[8]);
    sa = __fadd2_rn(__fadd2_rn(sa, sb), sc);
    const float nmean = -(sa.x + sa.y) * (1.0f / 18);
    const float2 nm2 = make_float2(nmean, nmean);
    float2 qa = make_float2(0.f, 0.f), qb = qa, qc = qa;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      v[j] = __fadd2_rn(v[j], nm2);
      if (j % 3 == 0) qa = __ffma2_rn(v[j], v[j], qa);
      else if (j % 3 == 1) qb = __ffma2_rn(v[j], v[j], qb);
      else qc = __ffma2_rn(v[j], v[j], qc);
    }
    qa = __fadd2_rn(__fadd2_rn(qa, qb), qc);
    const float rstd = rsqrtf((qa.x + qa.y) * (1.0f / 18) + eps);
    const float2 r2 = make_float2(rstd, rstd);
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const float2 n = __fmul2_rn(v[j], r2);
      o[j] = pack_bf16(n.x, n.y);
    }
    o[9] = 0x3F803F80u;                                // columns 18, 19 = 1, 1
  } else {
#pragma unroll
    for (int j = 0; j < 9; ++j) o[j] = 0u;
    o[9] = 0x00003F80u;                                // a padded slot: zeros, column 18 = 1
  }
  o[10] = 0u; o[11] = 0u;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch)
    *reinterpret_cast<uint4*>(tile + ch * (128 * 16) + row * 16) = make_uint4(o[4 * ch], o[4 * ch + 1], o[4 * ch + 2], o[4 * ch + 3]);
}

// 18 fp32 values -> three 16-byte chunks (columns 18..23 zero)
__device__ __forceinline__ void pack18(const float* v, uint4& c0, uint4& c1, uint4& c2) {
  c0 = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  c1 = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
  c2 = make_uint4(pack_bf16(v[16], v[17]), 0u, 0u, 0u);
}
}  // namespace v3

template <bool CROSS>
__global__ void __launch_bounds__(128, CROSS ? 4 : 5) window_attn_v3_kernel(const __grid_constant__ AttnV3Params p) {
  using namespace umma;
  using A = AttnV3;
  constexpr int S = A::S, WIN = A::WIN, C = A::C;
  constexpr int SEC = CROSS ? A::SEC_CROSS : A::SEC_SELF;
  constexpr int TMEM_COLS = CROSS ? 128 : 64;
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) uint64_t bar, wbar;
  __shared__ uint32_t tmem_base_s;
  __shared__ uint32_t valid_bits[4];

  HRF_PROF_DECL
  pdl_launch_dependents();
  const int tid = threadIdx.x, warp = warp_idx_uniform(), lane = tid & 31;
  const int g = tid >> 6, i = tid & 63;              // window of the pair, slot in the window
  constexpr int o_xn = 0, o_zn = A::XT;              // self: K / V' come from the x^ tile itself
  constexpr int o_q = (CROSS ? 2 : 1) * A::XT, o_k = o_q + A::QT, o_w = o_q + A::REGION;
  constexpr int o_kvsrc = CROSS ? o_zn : o_xn;       // tile K / V' are projected from; V' then overwrites it
  const int n_mod = CROSS ? p.n_mod : 1;
  const int n_prob = CROSS ? 1 : p.n_prob;
  const int n_w = CROSS ? n_mod : n_prob;          // weight sections resident in shared memory

  // ---- one-time setup: weights + tables by bulk copies, zero columns, TMEM ---------------------
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_init(&wbar, 1);
    fence_mbar_init();
    mbar_expect_tx(&wbar, (uint32_t)(n_w * SEC));
    for (int m = 0; m < n_w; ++m)
      bulk_g2s(sm + o_w + m * SEC,
               reinterpret_cast<const unsigned char*>(p.blob[m] + p.v3_off) + (CROSS ? A::SEC_SELF : 0), SEC, &wbar);
  }
  {
    // chunk 3 (columns 24..31) of the x^ / z^ / V' tiles is never written again: zero
    const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(sm + o_xn + 3 * (128 * 16) + tid * 16) = z4;
    if (CROSS) *reinterpret_cast<uint4*>(sm + o_zn + 3 * (128 * 16) + tid * 16) = z4;
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, TMEM_COLS);
  bool w_ready = false;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
  uint32_t phase = 0;

  const uint32_t a_xn = smem_u32(sm + o_xn), a_kv = smem_u32(sm + o_kvsrc);
  const uint32_t a_q = smem_u32(sm + o_q), a_k = smem_u32(sm + o_k), a_w = smem_u32(sm + o_w);

  const int nWh = ceil_div(p.H, WIN), nWw = ceil_div(p.W, WIN);
  const int pad_h = nWh * WIN - p.H, pad_w = nWw * WIN - p.W;
  const int pad_t = pad_h / 2, pad_l = pad_w / 2;
  const bool use_mask = p.pad_mask && pad_h > 0 && pad_w > 0;
  const int n_windows = p.B * nWh * nWw;
  const int n_tiles = (n_windows + 1) / 2;
  const int total_tiles = n_tiles * n_prob;          // global tile = problem * n_tiles + tile
  auto x_of = [&](int pr) { return static_cast<const __nv_bfloat16*>(CROSS ? p.x : p.xs[pr]); };
  auto out_of = [&](int pr) { return static_cast<__nv_bfloat16*>(CROSS ? p.out : p.outs[pr]); };

  // relative-position bias of (query slot i, key slot j): table[rp_base - (jh * 13 + jw)]
  const int ic = i < S ? i : S - 1;
  const int rp_base = (ic / WIN + WIN - 1) * (2 * WIN - 1) + (ic % WIN) + WIN - 1;

  auto row_token = [&](int tile) -> int {
    const int wdx = tile * 2 + g;
    if (tile >= n_tiles || wdx >= n_windows || i >= S) return -1;
    int b, rem, wy, wx;
    p.d_win_img.divmod(wdx, b, rem);
    p.d_win_row.divmod(rem, wy, wx);
    const int h = wy * WIN + i / WIN - pad_t, w = wx * WIN + i % WIN - pad_l;
    return (h >= 0 && h < p.H && w >= 0 && w < p.W) ? (b * p.H + h) * p.W + w : -1;
  };

  uint32_t xr[9], zr[9];
  pdl_wait();                                        // everything above read only the weight blobs
  int tok = -1;
  if ((int)blockIdx.x < total_tiles) {
    const int pr0 = n_prob > 1 ? p.d_tiles.div(blockIdx.x) : 0;
    tok = row_token(blockIdx.x - pr0 * n_tiles);
    if (tok >= 0) {
      load_row_raw<C>(x_of(pr0) + (size_t)tok * C, xr);
      if (CROSS) load_row_raw<C>(static_cast<const __nv_bfloat16*>(p.z[0]) + (size_t)tok * C, zr);
    }
  }

  HRF_PROF(14)
  for (int gt = blockIdx.x; gt < total_tiles; gt += gridDim.x) {
    HRF_PROF_TILE
    const int pr = n_prob > 1 ? p.d_tiles.div(gt) : 0, tile = gt - pr * n_tiles;
    __nv_bfloat16* out = out_of(pr);
    const int wdx = tile * 2 + g;
    {
      const unsigned bal = __ballot_sync(0xffffffffu, tok >= 0);
      if (lane == 0) valid_bits[warp] = bal;
    }
    // ---- x^ -> operand tile; the next tile's rows are requested behind it --------------------------
    v3::ln18_to_tile(xr, tok >= 0, p.eps, sm + o_xn, tid);
    // Global loads are requested only BEHIND the last proxy fence of a tile (fence.proxy.async
    // compiles to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC, and the MEMBAR waits for every load the
    // thread has in flight): the next rows travel while the P V' MMA and the output epilogue run.
    int tok_next = -1;
    uint32_t xnext[9];
    float acc[C];                                    // cross: sum over the modalities of z + attention
    if (CROSS) {
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] = 0.f;
    }

#pragma unroll 1
    for (int m = 0; m < n_mod; ++m) {
      const int wi = CROSS ? m : pr;                 // weight section: modality (cross) / problem (self)
      const unsigned char* wsec = sm + o_w + wi * SEC;
      const uint32_t a_wm = a_w + wi * SEC;
      uint32_t zcur[9];                              // this modality's raw rows (the "+ z" residual)
      if (CROSS) {
#pragma unroll
        for (int j = 0; j < 9; ++j) zcur[j] = zr[j];
        v3::ln18_to_tile(zcur, tok >= 0, p.eps, sm + o_zn, tid);
      }
      if (!w_ready) {
        mbar_wait(&wbar, 0);
        w_ready = true;
      }
      HRF_PROF(0)
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      HRF_PROF(1)

      // ---- projections: Q | K | V' --------------------------------------------------------------------
      if (warp == 0 && elect_one()) {
        tc_fence_after();
        if constexpr (!CROSS) {
          constexpr uint32_t idp = idesc_bf16(128, 64, false, false);
#pragma unroll
          for (int s = 0; s < 2; ++s) mma_bf16(tmem, desc_kmajor(a_xn, 128, s), desc_kmajor(a_wm, 64, s), idp, s > 0);
        } else {
          constexpr uint32_t idq = idesc_bf16(128, 32, false, false), idk = idesc_bf16(128, 48, false, false);
#pragma unroll
          for (int s = 0; s < 2; ++s) mma_bf16(tmem, desc_kmajor(a_xn, 128, s), desc_kmajor(a_wm, 32, s), idq, s > 0);
#pragma unroll
          for (int s = 0; s < 2; ++s)
            mma_bf16(tmem + 32, desc_kmajor(a_kv, 128, s), desc_kmajor(a_wm + A::W_Q_B, 48, s), idk, s > 0);
        }
        mma_commit(&bar);
      }
      HRF_PROF(2)
      cta_wait(&bar, phase);
      phase ^= 1;
      tc_fence_after();
      HRF_PROF(3)
      {
        // self: Q at columns 0..17, K 18..35, V' 36..54; cross: Q 0..17, K 32..49, V' 50..68
        constexpr int cK = CROSS ? 32 : 18, cV = CROSS ? 50 : 36;
        const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
        float v[32];
        uint4 c0, c1, c2;
        // Q: its window's half of the K dimension, zeros in the other half
        tmem_ld32(trow, v);
        tmem_ld_wait();
        v3::pack18(v, c0, c1, c2);
        unsigned char* qrow = sm + o_q + tid * 16;
        *reinterpret_cast<uint4*>(qrow + (3 * g + 0) * 2048) = c0;
        *reinterpret_cast<uint4*>(qrow + (3 * g + 1) * 2048) = c1;
        *reinterpret_cast<uint4*>(qrow + (3 * g + 2) * 2048) = c2;
        *reinterpret_cast<uint4*>(qrow + (3 * (1 - g) + 0) * 2048) = z4;
        *reinterpret_cast<uint4*>(qrow + (3 * (1 - g) + 1) * 2048) = z4;
        *reinterpret_cast<uint4*>(qrow + (3 * (1 - g) + 2) * 2048) = z4;
        // K: row = key slot, columns = this window's half
        if constexpr (!CROSS) {
          // columns 18..35 straddle the first load: 18..31 are in v, 32..35 come with the next
          float k2[32];
          tmem_ld32(trow + 32, k2);
          tmem_ld_wait();
          float kk[18];
#pragma unroll
          for (int c = 0; c < 14; ++c) kk[c] = v[18 + c];
#pragma unroll
          for (int c = 0; c < 4; ++c) kk[14 + c] = k2[c];
          v3::pack18(kk, c0, c1, c2);
          unsigned char* krow = sm + o_k + i * 16;
          *reinterpret_cast<uint4*>(krow + (3 * g + 0) * 1024) = c0;
          *reinterpret_cast<uint4*>(krow + (3 * g + 1) * 1024) = c1;
          *reinterpret_cast<uint4*>(krow + (3 * g + 2) * 1024) = c2;
          // V': columns 36..54 = k2[4..22]; column 18 of V' is the constant 1
          v3::pack18(k2 + 4, c0, c1, c2);
          c2.y = v3::pack_bf16(k2[22], 0.f);
          unsigned char* vrow = sm + o_kvsrc + tid * 16;
          *reinterpret_cast<uint4*>(vrow) = c0;
          *reinterpret_cast<uint4*>(vrow + 2048) = c1;
          *reinterpret_cast<uint4*>(vrow + 4096) = c2;
        } else {
          tmem_ld32(trow + cK, v);                   // K 32..49 -> v[0..17]; V' 50..63 -> v[18..31]
          float v2[8];
          tmem_ld8(trow + 64, v2);                   // V' 64..68 -> v2[0..4]
          tmem_ld_wait();
          v3::pack18(v, c0, c1, c2);
          unsigned char* krow = sm + o_k + i * 16;
          *reinterpret_cast<uint4*>(krow + (3 * g + 0) * 1024) = c0;
          *reinterpret_cast<uint4*>(krow + (3 * g + 1) * 1024) = c1;
          *reinterpret_cast<uint4*>(krow + (3 * g + 2) * 1024) = c2;
          float vv[19];
#pragma unroll
          for (int c = 0; c < 14; ++c) vv[c] = v[18 + c];
#pragma unroll
          for (int c = 0; c < 5; ++c) vv[14 + c] = v2[c];
          v3::pack18(vv, c0, c1, c2);
          c2.y = v3::pack_bf16(vv[18], 0.f);
          unsigned char* vrow = sm + o_kvsrc + tid * 16;
          *reinterpret_cast<uint4*>(vrow) = c0;
          *reinterpret_cast<uint4*>(vrow + 2048) = c1;
          *reinterpret_cast<uint4*>(vrow + 4096) = c2;
        }
        (void)cV;
      }

      // ---- S = Q K^T: both windows in one M = 128, N = 64, K = 48 MMA ---------------------------------
      HRF_PROF(4)
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      if (warp == 0 && elect_one()) {
        tc_fence_after();
        constexpr uint32_t ids = idesc_bf16(128, 64, false, false);
#pragma unroll
        for (int s = 0; s < 3; ++s) mma_bf16(tmem, desc_kmajor(a_q, 128, s), desc_kmajor(a_k, 64, s), ids, s > 0);
        mma_commit(&bar);
      }
      HRF_PROF(5)
      cta_wait(&bar, phase);
      phase ^= 1;
      tc_fence_after();
      HRF_PROF(6)

      // ---- softmax numerators over the 49 keys of this row's window (base 2: scale and log2 e are in Wq)
      {
        float sc[56];
        tmem_ld32(trow, sc);
        tmem_ld16(trow + 32, sc + 32);
        tmem_ld8(trow + 48, sc + 48);
        tmem_ld_wait();
        const float* tb = reinterpret_cast<const float*>(wsec + (CROSS ? A::W_Q_B + A::W_KV_B : A::W_SELF_B)) + rp_base;
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < S; ++j) {
          sc[j] += tb[-((j / WIN) * (2 * WIN - 1) + (j % WIN))];
          mx = fmaxf(mx, sc[j]);
        }
        if (use_mask) {                              // uniform branch: no shipped config masks pad keys
          const unsigned long long vmask =
              (unsigned long long)valid_bits[2 * g] | ((unsigned long long)valid_bits[2 * g + 1] << 32);
          const bool mask_me = wdx < n_windows;
          mx = -INFINITY;
#pragma unroll
          for (int j = 0; j < S; ++j) {
            if (mask_me && !((vmask >> j) & 1ull)) sc[j] = -INFINITY;
            mx = fmaxf(mx, sc[j]);
          }
        }
#pragma unroll
        for (int j = 0; j < S; ++j) sc[j] = fast_exp2(sc[j] - mx);
#pragma unroll
        for (int j = S; j < 56; ++j) sc[j] = 0.f;
        unsigned char* sP = sm + o_q;                // Q | K are dead: S is complete
#pragma unroll
        for (int ch = 0; ch < 7; ++ch) st_chunk(sP, tid, ch, 128, sc + 8 * ch);
        *reinterpret_cast<uint4*>(sP + 7 * 2048 + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
      }

      // ---- P V' (both windows' V'; a row keeps the product with its own) ------------------------------
      HRF_PROF(7)
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      if (warp == 0 && elect_one()) {
        tc_fence_after();
        constexpr uint32_t ido = idesc_bf16(128, 32, false, true);
#pragma unroll
        for (int g2 = 0; g2 < 2; ++g2)
#pragma unroll
          for (int s = 0; s < 4; ++s)
            mma_bf16(tmem + g2 * 32, desc_kmajor(a_q, 128, s), desc_mnmajor(a_kv + g2 * 64 * 16, 128, s), ido, s > 0);
        mma_commit(&bar);
      }
      if (m + 1 < n_mod) {                           // cross: the next modality's rows of this tile
        if (CROSS && tok >= 0) load_row_raw<C>(static_cast<const __nv_bfloat16*>(p.z[CROSS ? m + 1 : 0]) + (size_t)tok * C, zr);
      } else {                                       // the next tile's rows
        const int gn = gt + gridDim.x;
        if (gn < total_tiles) {
          const int prn = n_prob > 1 ? p.d_tiles.div(gn) : 0;
          tok_next = row_token(gn - prn * n_tiles);
          if (tok_next >= 0) {
            load_row_raw<C>(x_of(prn) + (size_t)tok_next * C, xnext);
            if (CROSS) load_row_raw<C>(static_cast<const __nv_bfloat16*>(p.z[0]) + (size_t)tok_next * C, zr);
          }
        }
      }
      HRF_PROF(8)
      cta_wait(&bar, phase);
      phase ^= 1;
      tc_fence_after();
      HRF_PROF(9)
      {
        float o[24];
        tmem_ld16(trow + g * 32, o);
        tmem_ld8(trow + g * 32 + 16, o + 16);
        tmem_ld_wait();
        const float inv = 1.0f / o[18];              // the constant-1 column of V': sum_j P_ij
        if constexpr (!CROSS) {
          if (tok >= 0) {
            uint32_t w[9];
#pragma unroll
            for (int j = 0; j < 9; ++j)
              w[j] = v3::pack_bf16(fmaf(o[2 * j], inv, __uint_as_float(xr[j] << 16)),
                                   fmaf(o[2 * j + 1], inv, __uint_as_float(xr[j] & 0xffff0000u)));
            uint32_t* dst = reinterpret_cast<uint32_t*>(out + (size_t)tok * C);
#pragma unroll
            for (int j = 0; j < 9; ++j) dst[j] = w[j];
          }
        } else {
#pragma unroll
          for (int j = 0; j < 9; ++j) {
            acc[2 * j] += fmaf(o[2 * j], inv, __uint_as_float(zcur[j] << 16));
            acc[2 * j + 1] += fmaf(o[2 * j + 1], inv, __uint_as_float(zcur[j] & 0xffff0000u));
          }
        }
      }
    }
    if constexpr (CROSS) {
      if (tok >= 0) {
        uint32_t w[9];
#pragma unroll
        for (int j = 0; j < 9; ++j)
          w[j] = v3::pack_bf16(acc[2 * j] + __uint_as_float(xr[j] << 16), acc[2 * j + 1] + __uint_as_float(xr[j] & 0xffff0000u));
        uint32_t* dst = reinterpret_cast<uint32_t*>(out + (size_t)tok * C);
#pragma unroll
        for (int j = 0; j < 9; ++j) dst[j] = w[j];
      }
    }
    // rotate the software pipeline; the next iteration's first barrier orders this tile's TMEM
    // reads before its MMAs
    tok = tok_next;
#pragma unroll
    for (int j = 0; j < 9; ++j) xr[j] = xnext[j];
    HRF_PROF(10)
  }
  HRF_PROF_END

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

static int launch_window_attn_v3(AttnV3Params p, bool cross, cudaStream_t stream) {
  using A = AttnV3;
  p.d_win_row = FastDiv(ceil_div(p.W, 7));
  p.d_win_img = FastDiv(ceil_div(p.H, 7) * ceil_div(p.W, 7));
  const int n_windows = p.B * ceil_div(p.H, 7) * ceil_div(p.W, 7);
  if (cross) p.n_prob = 1;
  p.d_tiles = FastDiv((n_windows + 1) / 2);
  const int n_tiles = ((n_windows + 1) / 2) * p.n_prob;
  const int smem = A::smem_bytes(cross, cross ? p.n_mod : p.n_prob);
  static const int env_per_sm = [] { const char* e = std::getenv("HRF_ATTN_CTAS_PER_SM"); return e && atoi(e) > 0 ? atoi(e) : 0; }();
  const int by_smem = (227 * 1024) / (smem + 1024 + 64), by_tmem = 512 / A::tmem_cols(cross);
  int per_sm = by_smem < by_tmem ? by_smem : by_tmem;
  if (per_sm > (cross ? 4 : 5)) per_sm = cross ? 4 : 5;          // __launch_bounds__ register budget
  if (env_per_sm > 0 && env_per_sm < per_sm) per_sm = env_per_sm;
  const int cap = 148 * (per_sm > 0 ? per_sm : 1);
  // HRF_BALANCED_GRID=1 (experiment): ceil(tiles / rounds) CTAs that all walk `rounds` tiles instead of
  // one full resident wave -- the same step time (2.799 vs 2.802 ms: the slots it frees for the other
  // streams' kernels buy nothing), and a slower kernel on its own (the last, partly filled round of
  // a full wave runs uncontended: MixFFN 19.8 vs 23.0 us), so the full wave is the default
  static const bool balanced = [] { const char* e = std::getenv("HRF_BALANCED_GRID"); return e && e[0] == '1'; }();
  const int rounds = ceil_div(n_tiles, cap);
  const int grid = !balanced ? (n_tiles < cap ? n_tiles : cap) : ceil_div(n_tiles, rounds);
  for (int m = 0; m < (cross ? p.n_mod : p.n_prob); ++m)
    HRF_REQUIRE((reinterpret_cast<uintptr_t>(p.blob[m] + p.v3_off) & 15) == 0, HRF_EINVAL, "attn_v3: blob must be 16-byte aligned");
  if (cross) {
    HRF_CUDA(ensure_smem((const void*)window_attn_v3_kernel<true>, A::smem_bytes(true, A::MAXMOD)));
    HRF_CUDA(launch_pdl(window_attn_v3_kernel<true>, dim3(grid), dim3(128), (size_t)smem, stream, p));
  } else {
    HRF_CUDA(ensure_smem((const void*)window_attn_v3_kernel<false>, A::smem_bytes(false, A::MAXMOD)));
    HRF_CUDA(launch_pdl(window_attn_v3_kernel<false>, dim3(grid), dim3(128), (size_t)smem, stream, p));
  }
  count_launch();
  HRF_CUDA(cudaGetLastError());
  return HRF_OK;
}

}  // namespace hrf
