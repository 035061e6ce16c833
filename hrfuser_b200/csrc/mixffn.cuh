// Fused MixFFN (CrossFFN) kernel, fp32-math SIMT version.
//
//   out = x + GELU(BN3(W2 * GELU(BN2(dw3x3(GELU(BN1(W1*LN(x)+b1)))+bd)) + b2))
//
// One CTA owns an 8x8 spatial tile of tokens.  The 10x10 halo of LayerNormed
// tokens is staged in shared memory once; the 4C hidden channels are processed
// in chunks (fc1 on the halo -> GELU -> depthwise 3x3 -> GELU -> partial fc2
// accumulated in registers), so the hidden activation (4C wide, six HBM round
// trips in the reference: hrformer.py:267-282) never leaves the SM.
// Halo tokens outside the image are zeros *after* fc1/BN/GELU, which is what
// the reference's zero-padded depthwise conv sees (hrformer.py:271-277).
#pragma once
#include "common.cuh"

namespace hrf {

struct FfnLayout {
  int C, Cp, hidden, HC;
  int o_ln_w, o_ln_b, o_w1, o_b1, o_wd, o_bd, o_w2, o_b2, total;
  int ldx, ldh;
  __host__ __device__ FfnLayout(int C_, int hidden_) {
    C = C_; hidden = hidden_; Cp = round_up(C, 4);
    // hidden chunk: largest divisor of `hidden` that is a multiple of 4 and <= 96
    HC = 4;
    for (int c = 4; c <= 96 && c <= hidden; c += 4)
      if (hidden % c == 0) HC = c;
    const int c4 = round_up(C, 4);
    int o = 0;
    o_ln_w = o; o += c4; o_ln_b = o; o += c4;
    o_w1 = o; o += Cp * hidden; o_b1 = o; o += hidden;     // W1t [Cp][hidden]
    o_wd = o; o += 9 * hidden; o_bd = o; o += hidden;      // Wd  [9][hidden]
    o_w2 = o; o += hidden * C; o = round_up(o, 4);         // W2t [hidden][C]
    o_b2 = o; o += c4;
    // tensor-core (bf16) sections, present when hidden splits into chunks of 72
    // (HRFuser-T) or 78 (HRFuser-B) channels: per chunk, fp32 b1[80] | wd[9][80] | bd[80]
    // (zero padded to 80), then b2[NOUT]; then bf16 operand tiles in the
    // chunk-major layout of umma.cuh: W1 [KC/8][80][8] and W2 [80/8][NOUT][8] per chunk.
    tc_KC = round_up(C, 16); tc_NOUT = tc_KC;
    // chunk width: 72 hidden channels (HRFuser-T widths, hidden = 4 * 18k) or 78 (HRFuser-B)
    tc_CH = (hidden % 72 == 0) ? 72 : (hidden % 78 == 0) ? 78 : 0;
    tc_nchunk = tc_CH ? hidden / tc_CH : 0;
    o_tc_f32 = o; o += tc_nchunk * (11 * 80) + (tc_nchunk ? tc_NOUT : 0);
    o = round_up(o, 4);
    o_tc_w1 = o; o += tc_nchunk * 80 * tc_KC / 2;
    o_tc_w2 = o; o += tc_nchunk * tc_NOUT * 80 / 2;
    // second-generation kernel (mixffn_v2.cuh), per 72-channel chunk: bf16 W1 tile [KC/8][80][8]
    // with the LayerNorm affine folded in (row C = folded bias), fp16 W2 tile [10][NOUT][8]
    // (row 72 of chunk 0 = b2), fp16 depthwise table wd[9][80] | bd[80]; every section is
    // pre-multiplied by 0.5 (the GELU behind it takes x / 2)
    const int v2 = (tc_CH == 72) ? tc_nchunk : 0;
    o_v2_w1 = o; o += v2 * 80 * tc_KC / 2;
    o_v2_w2 = o; o += v2 * tc_NOUT * 80 / 2;
    o_v2_cv = o; o += v2 * 400;
    total = o;
    ldx = stride4odd(Cp);
    ldh = stride4odd(HC);
  }
  int tc_KC, tc_NOUT, tc_CH, tc_nchunk, o_tc_f32, o_tc_w1, o_tc_w2, o_v2_w1, o_v2_w2, o_v2_cv;
};

constexpr int kFfnThreads = 256;
constexpr int kFfnTile = 8;                       // 8x8 outputs
constexpr int kFfnHalo = kFfnTile + 2;            // 10x10 inputs
constexpr int kFfnOut = kFfnTile * kFfnTile;      // 64
constexpr int kFfnIn = kFfnHalo * kFfnHalo;       // 100

__host__ __device__ inline size_t ffn_smem_floats(const FfnLayout& L) {
  return (size_t)kFfnIn * L.ldx + (size_t)kFfnIn * L.ldh + (size_t)kFfnOut * L.ldh + kFfnIn + 4;
}

struct FfnParams {
  const void* x;
  const float* blob;
  void* out;
  void* ws;           // fp32 workspace (split tensor-core variant) or nullptr
  int B, H, W, C, hidden;
  float eps;
  FastDiv d_tiles_xy, d_tiles_x;   // tensor-core kernel: tile index -> (image, tile row, tile col)
};

// MAXT: fc2 output tiles (4 rows x CT cols) per thread = ceil(16 * C/CT / 256)
template <typename T, int CT, int MAXT>
__global__ void __launch_bounds__(kFfnThreads) mixffn_kernel(FfnParams p) {
  extern __shared__ __align__(16) float smem[];
  const FfnLayout L(p.C, p.hidden);
  const int C = L.C, HC = L.HC;
  const int warp = threadIdx.x / 32, lane = threadIdx.x & 31, nwarps = blockDim.x / 32;
  float* XN = smem;                                   // [100][ldx]
  float* H1 = XN + (size_t)kFfnIn * L.ldx;            // [100][ldh]
  float* H2 = H1 + (size_t)kFfnIn * L.ldh;            // [64][ldh]
  int* inside = reinterpret_cast<int*>(H2 + (size_t)kFfnOut * L.ldh);  // [100]

  const int tiles_x = ceil_div(p.W, kFfnTile), tiles_y = ceil_div(p.H, kFfnTile);
  const int tile = blockIdx.x;
  const int b = tile / (tiles_x * tiles_y);
  const int ty0 = ((tile / tiles_x) % tiles_y) * kFfnTile, tx0 = (tile % tiles_x) * kFfnTile;
  const float* blob = p.blob;
  const T* x = static_cast<const T*>(p.x);
  T* out = static_cast<T*>(p.out);

  // ---- stage LN(x) for the halo tile ----------------------------------------
  for (int s = warp; s < kFfnIn; s += nwarps) {
    const int h = ty0 - 1 + s / kFfnHalo, w = tx0 - 1 + s % kFfnHalo;
    const bool valid = h >= 0 && h < p.H && w >= 0 && w < p.W;
    if (lane == 0) inside[s] = valid;
    if (valid) {
      const T* src = x + ((size_t)(b * p.H + h) * p.W + w) * C;
      warp_layernorm([&](int c) { return Elem<T>::ld(src + c); }, C, L.ldx, blob + L.o_ln_w,
                     blob + L.o_ln_b, p.eps, XN + (size_t)s * L.ldx);
    } else {
      for (int c = lane; c < L.ldx; c += 32) XN[(size_t)s * L.ldx + c] = 0.f;
    }
  }

  // fc2 accumulators: this thread's tiles are t = threadIdx.x + i*blockDim.x
  const int ntc = C / CT;
  const int ntiles = (kFfnOut / 4) * ntc;
  float acc[MAXT][4][CT];
#pragma unroll
  for (int i = 0; i < MAXT; ++i)
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < CT; ++c) acc[i][r][c] = 0.f;
  __syncthreads();

  for (int c0 = 0; c0 < L.hidden; c0 += HC) {
    // ---- fc1 + BN1 + GELU on the halo (zeros outside the image) -------------
    {
      const float* b1 = blob + L.o_b1 + c0;
      const int ldh = L.ldh;
      // W1t chunk: columns c0..c0+HC of a [Cp][hidden] matrix -> row stride `hidden`
      const float* W1 = blob + L.o_w1 + c0;
      const int hidden = L.hidden;
      // block_gemm assumes row stride == N; use a strided variant inline
      const int ntc1 = HC / 4, ntr1 = kFfnIn / 4;
      for (int t = threadIdx.x; t < ntr1 * ntc1; t += blockDim.x) {
        const int n0 = (t % ntc1) * 4, r0 = (t / ntc1) * 4;
        float a1[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) a1[r][c] = 0.f;
        for (int k = 0; k < L.Cp; k += 4) {
          float4 a[4];
#pragma unroll
          for (int r = 0; r < 4; ++r)
            a[r] = *reinterpret_cast<const float4*>(XN + (size_t)(r0 + r) * L.ldx + k);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const float4 w4 =
                __ldg(reinterpret_cast<const float4*>(W1 + (size_t)(k + kk) * hidden + n0));
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const float av = kk == 0 ? a[r].x : kk == 1 ? a[r].y : kk == 2 ? a[r].z : a[r].w;
              a1[r][0] = fmaf(av, w4.x, a1[r][0]);
              a1[r][1] = fmaf(av, w4.y, a1[r][1]);
              a1[r][2] = fmaf(av, w4.z, a1[r][2]);
              a1[r][3] = fmaf(av, w4.w, a1[r][3]);
            }
          }
        }
        const float4 bb = __ldg(reinterpret_cast<const float4*>(b1 + n0));
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const bool in = inside[r0 + r] != 0;
          float4 o;
          o.x = in ? gelu_erf(a1[r][0] + bb.x) : 0.f;
          o.y = in ? gelu_erf(a1[r][1] + bb.y) : 0.f;
          o.z = in ? gelu_erf(a1[r][2] + bb.z) : 0.f;
          o.w = in ? gelu_erf(a1[r][3] + bb.w) : 0.f;
          *reinterpret_cast<float4*>(H1 + (size_t)(r0 + r) * ldh + n0) = o;
        }
      }
    }
    __syncthreads();

    // ---- depthwise 3x3 + BN2 + GELU ------------------------------------------
    {
      const float* wd = blob + L.o_wd + c0;
      const float* bd = blob + L.o_bd + c0;
      for (int e = threadIdx.x; e < kFfnOut * HC; e += blockDim.x) {
        const int j = e % HC, o = e / HC;
        const int oy = o / kFfnTile, ox = o % kFfnTile;
        float s = __ldg(bd + j);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx)
            s = fmaf(H1[(size_t)((oy + dy) * kFfnHalo + ox + dx) * L.ldh + j],
                     __ldg(wd + (size_t)(dy * 3 + dx) * L.hidden + j), s);
        H2[(size_t)o * L.ldh + j] = gelu_erf(s);
      }
    }
    __syncthreads();

    // ---- partial fc2: acc += H2[:, chunk] * W2t[c0:c0+HC, :] -----------------
    {
      const float* W2 = blob + L.o_w2 + (size_t)c0 * C;
#pragma unroll
      for (int i = 0; i < MAXT; ++i) {
        const int t = threadIdx.x + i * blockDim.x;
        if (t < ntiles) {
          const int n0 = (t % ntc) * CT, r0 = (t / ntc) * 4;
          for (int k = 0; k < HC; k += 4) {
            float4 a[4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
              a[r] = *reinterpret_cast<const float4*>(H2 + (size_t)(r0 + r) * L.ldh + k);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              float w[CT];
              const float* wp = W2 + (size_t)(k + kk) * C + n0;
              if constexpr (CT == 4) {
                const float4 w4 = __ldg(reinterpret_cast<const float4*>(wp));
                w[0] = w4.x; w[1] = w4.y; w[2] = w4.z; w[3] = w4.w;
              } else {
                const float2 w2 = __ldg(reinterpret_cast<const float2*>(wp));
                w[0] = w2.x; w[1] = w2.y;
              }
#pragma unroll
              for (int r = 0; r < 4; ++r) {
                const float av = kk == 0 ? a[r].x : kk == 1 ? a[r].y : kk == 2 ? a[r].z : a[r].w;
#pragma unroll
                for (int c = 0; c < CT; ++c) acc[i][r][c] = fmaf(av, w[c], acc[i][r][c]);
              }
            }
          }
        }
      }
    }
    __syncthreads();
  }

  // ---- epilogue: BN3 (folded) + GELU + residual ------------------------------
  const float* b2 = blob + L.o_b2;
#pragma unroll
  for (int i = 0; i < MAXT; ++i) {
    const int t = threadIdx.x + i * blockDim.x;
    if (t < ntiles) {
      const int n0 = (t % ntc) * CT, r0 = (t / ntc) * 4;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int o = r0 + r;
        const int h = ty0 + o / kFfnTile, w = tx0 + o % kFfnTile;
        if (h < p.H && w < p.W) {
          const size_t off = ((size_t)(b * p.H + h) * p.W + w) * C + n0;
#pragma unroll
          for (int c = 0; c < CT; ++c)
            Elem<T>::st(out + off + c,
                        Elem<T>::ld(x + off + c) + gelu_erf(acc[i][r][c] + __ldg(b2 + n0 + c)));
        }
      }
    }
  }
}

template <typename T>
static int launch_mixffn(const FfnParams& p, cudaStream_t stream) {
  const FfnLayout L(p.C, p.hidden);
  const size_t smem = ffn_smem_floats(L) * sizeof(float);
  const int grid = p.B * ceil_div(p.H, kFfnTile) * ceil_div(p.W, kFfnTile);
  const int CT = (p.C % 4 == 0) ? 4 : 2;
  const int maxt = ceil_div((kFfnOut / 4) * (p.C / CT), kFfnThreads);
  void (*kern)(FfnParams) = nullptr;
#define HRF_FFN_CASE(ct, mt) \
  if (CT == ct && maxt == mt) kern = mixffn_kernel<T, ct, mt>;
  HRF_FFN_CASE(2, 1) HRF_FFN_CASE(2, 2) HRF_FFN_CASE(2, 3)
  HRF_FFN_CASE(4, 1) HRF_FFN_CASE(4, 2) HRF_FFN_CASE(4, 3) HRF_FFN_CASE(4, 4)
#undef HRF_FFN_CASE
  HRF_REQUIRE(kern != nullptr, HRF_EUNSUPPORTED, "mixffn: C=%d needs %d tiles/thread", p.C, maxt);
  HRF_REQUIRE(smem <= 227 * 1024, HRF_EUNSUPPORTED, "mixffn: C=%d needs %zu B smem", p.C, smem);
  HRF_CUDA(ensure_smem((const void*)kern, smem));
  kern<<<grid, kFfnThreads, smem, stream>>>(p);
  count_launch();
  HRF_CUDA(cudaGetLastError());
  return HRF_OK;
}

}  // namespace hrf
