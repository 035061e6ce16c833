// Multi-resolution exchange kernels (HRModule.forward, reference hrnet.py:184-207
// with the fuse layers of hrformer.py:498-561) and boundary layout converters.
// All three are bandwidth kernels on channels-last tokens.
#pragma once
#include "common.cuh"

namespace hrf {

// ---------------------------------------------------------------------------
// 1x1 conv + folded BN (+ReLU):  blob = Wt [Kp][Cout] (k-major, BN scale folded)
// followed by bias [round_up(Cout,4)].
// ---------------------------------------------------------------------------
struct PwLayout {
  int Cin, Cout, Kp, o_w, o_b, total, lda;
  __host__ __device__ PwLayout(int cin, int cout) {
    Cin = cin; Cout = cout; Kp = round_up(cin, 4);
    o_w = 0; o_b = round_up(Kp * cout, 4); total = o_b + round_up(cout, 4);
    lda = stride4odd(Kp);
  }
};
constexpr int kPwThreads = 256;

// two consecutive channels as one 4-byte (bf16) / 8-byte (fp32) access; channel
// counts are even, so rows of any token start pair-aligned
template <typename T> struct Pair;
template <> struct Pair<float> {
  __device__ static float2 ld(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
  __device__ static void st(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
  __device__ static float2 rt(float2 v) { return v; }      // value as stored
};
template <> struct Pair<__nv_bfloat16> {
  __device__ static float2 ld(const __nv_bfloat16* p) {
    const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(p));
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
  }
  __device__ static void st(__nv_bfloat16* p, float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    *reinterpret_cast<uint32_t*>(p) = *reinterpret_cast<const uint32_t*>(&h);
  }
  __device__ static float2 rt(float2 v) { return __bfloat1622float2(__floats2bfloat162_rn(v.x, v.y)); }
};

struct PwParams {
  const void* x; const float* blob; void* out;
  int ntok, Cin, Cout, relu;
  int w_smem;     // stage the k-major weight matrix in shared memory first
};

// one coalesced copy of the weight matrix into shared memory (or pass-through)
__device__ inline const float* stage_weights(const float* gw, float* sw, int n, int enable) {
  if (!enable) return gw;
  for (int e = threadIdx.x * 4; e < n; e += blockDim.x * 4)
    *reinterpret_cast<float4*>(sw + e) = __ldg(reinterpret_cast<const float4*>(gw + e));
  return sw;
}

// TOK tokens per CTA: 64 for full-resolution maps, 16 for the low-resolution branches
// (1 920 tokens at C=144) so that the grid still covers the 148 SMs
template <typename T, int CT, int TOK>
__global__ void __launch_bounds__(kPwThreads) pw_kernel(PwParams p) {
  extern __shared__ __align__(16) float smem[];
  const PwLayout L(p.Cin, p.Cout);
  const T* x = static_cast<const T*>(p.x);
  T* out = static_cast<T*>(p.out);
  const int t0 = blockIdx.x * TOK;
  const int m = min(TOK, p.ntok - t0);
  pdl_launch_dependents();
  const float* Wt = stage_weights(p.blob + L.o_w, smem + TOK * L.lda, L.Kp * p.Cout, p.w_smem);
  pdl_wait();
  // stage the token tile (contiguous in memory) as fp32, zero-padded to lda
  const int hp = L.lda / 2;
  for (int e = threadIdx.x; e < TOK * hp; e += blockDim.x) {
    const int r = e / hp, c = (e - r * hp) * 2;
    float2 v = make_float2(0.f, 0.f);
    if (r < m && c < p.Cin) v = Pair<T>::ld(x + (size_t)(t0 + r) * p.Cin + c);
    *reinterpret_cast<float2*>(smem + r * L.lda + c) = v;
  }
  __syncthreads();
  const float* bias = p.blob + L.o_b;
  const int relu = p.relu, Cout = p.Cout;
  block_gemm<4, CT>(smem, L.lda, m, Wt, L.Kp, Cout, [&](int r, int n, float v) {
    v += __ldg(bias + n);
    if (relu) v = fmaxf(v, 0.f);
    Elem<T>::st(out + (size_t)(t0 + r) * Cout + n, v);
  });
}

template <typename T, int TOK>
static int launch_pw_tok(PwParams p, cudaStream_t stream) {
  const PwLayout L(p.Cin, p.Cout);
  size_t smem = (size_t)TOK * L.lda * sizeof(float);
  p.w_smem = smem + (size_t)L.Kp * p.Cout * sizeof(float) <= 160 * 1024;
  if (p.w_smem) smem += (size_t)L.Kp * p.Cout * sizeof(float);
  HRF_REQUIRE(smem <= 227 * 1024, HRF_EUNSUPPORTED, "pw: Cin=%d too wide", p.Cin);
  auto kern = (p.Cout % 4 == 0) ? pw_kernel<T, 4, TOK> : pw_kernel<T, 2, TOK>;
  HRF_CUDA(ensure_smem((const void*)kern, smem));
  HRF_CUDA(launch_pdl(kern, dim3(ceil_div(p.ntok, TOK)), dim3(kPwThreads), smem, stream, p));
  count_launch();
  HRF_CUDA(cudaGetLastError());
  return HRF_OK;
}

template <typename T>
static int launch_pw(const PwParams& p, cudaStream_t stream) {
  HRF_REQUIRE(p.Cout % 2 == 0 && p.Cin % 2 == 0, HRF_EUNSUPPORTED,
              "pw: Cin=%d / Cout=%d must be even", p.Cin, p.Cout);
  return p.ntok >= 148 * 2 * 64 ? launch_pw_tok<T, 64>(p, stream) : launch_pw_tok<T, 16>(p, stream);
}

// ---------------------------------------------------------------------------
// dense 3x3 conv (pad 1, stride 1 or 2) + folded BN (+ReLU) on channels-last tokens: the
// transition layers between the HRNet stages and towards the fusion blocks
// (reference hrformer.py `_make_transition_layer` :562-607 and hrfuser_hrformer_based.py
// `_make_transition_layer_modality`).  Implicit GEMM: the CTA gathers the 9 taps of TOK output
// tokens into an fp32 [TOK][9*Cin] tile (zero outside the image) and runs the block GEMM of the
// 1x1 kernel on it.  blob = PwLayout(9*Cin, Cout), K index = tap*Cin + c.
// ---------------------------------------------------------------------------
struct Conv3Params {
  const void* x; const float* blob; void* out;
  int B, H, W, Ho, Wo, Cin, Cout, stride, relu;
  int w_smem;
  FastDiv d_cin, d_wo, d_howo;
};

template <typename T, int CT, int TOK>
__global__ void __launch_bounds__(kPwThreads) conv3x3_kernel(Conv3Params p) {
  extern __shared__ __align__(16) float smem[];
  const int K = 9 * p.Cin;
  const PwLayout L(K, p.Cout);
  const T* x = static_cast<const T*>(p.x);
  T* out = static_cast<T*>(p.out);
  const int ntok = p.B * p.Ho * p.Wo;
  const int t0 = blockIdx.x * TOK;
  const int m = min(TOK, ntok - t0);
  pdl_launch_dependents();
  const float* Wt = stage_weights(p.blob + L.o_w, smem + TOK * L.lda, L.Kp * p.Cout, p.w_smem);
  pdl_wait();
  // gather: element (r, k = tap*Cin + c) of the patch matrix, two channels at a time
  const int hk = K / 2, hp = L.lda / 2;
  for (int e = threadIdx.x; e < TOK * hp; e += blockDim.x) {
    const int r = e / hp, k2 = e - r * hp;
    float2 v = make_float2(0.f, 0.f);
    if (r < m && k2 < hk) {
      int tap, c;
      p.d_cin.divmod(2 * k2, tap, c);
      int b, rem, oy, ox;
      p.d_howo.divmod(t0 + r, b, rem);
      p.d_wo.divmod(rem, oy, ox);
      const int iy = oy * p.stride - 1 + tap / 3, ix = ox * p.stride - 1 + tap % 3;
      if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W)
        v = Pair<T>::ld(x + ((size_t)(b * p.H + iy) * p.W + ix) * p.Cin + c);
    }
    *reinterpret_cast<float2*>(smem + r * L.lda + 2 * k2) = v;
  }
  __syncthreads();
  const float* bias = p.blob + L.o_b;
  const int relu = p.relu, Cout = p.Cout;
  block_gemm<4, CT>(smem, L.lda, m, Wt, L.Kp, Cout, [&](int r, int n, float v) {
    v += __ldg(bias + n);
    if (relu) v = fmaxf(v, 0.f);
    Elem<T>::st(out + (size_t)(t0 + r) * Cout + n, v);
  });
}

template <typename T, int TOK>
static int launch_conv3x3_tok(Conv3Params p, cudaStream_t stream) {
  const PwLayout L(9 * p.Cin, p.Cout);
  size_t smem = (size_t)TOK * L.lda * sizeof(float);
  p.w_smem = smem + (size_t)L.Kp * p.Cout * sizeof(float) <= 160 * 1024;
  if (p.w_smem) smem += (size_t)L.Kp * p.Cout * sizeof(float);
  HRF_REQUIRE(smem <= 227 * 1024, HRF_EUNSUPPORTED, "conv3x3: Cin=%d too wide", p.Cin);
  auto kern = (p.Cout % 4 == 0) ? conv3x3_kernel<T, 4, TOK> : conv3x3_kernel<T, 2, TOK>;
  HRF_CUDA(ensure_smem((const void*)kern, smem));
  const int ntok = p.B * p.Ho * p.Wo;
  HRF_CUDA(launch_pdl(kern, dim3(ceil_div(ntok, TOK)), dim3(kPwThreads), smem, stream, p));
  count_launch();
  HRF_CUDA(cudaGetLastError());
  return HRF_OK;
}

// Direct form for the narrow transition convs (K = 9*Cin is small; the op is launch / latency
// bound, so it wants many short threads): a thread owns one output token x NCO output channels,
// a warp covers 32 consecutive tokens of one output-channel group, so the weights (fp32
// [K][NCO] slice of this group in shared memory) are warp-broadcast 8-byte loads and every
// multiply-add is a packed fp32x2 FMA on a channel pair.  Per input-channel pair the nine
// taps are requested together.  grid = (token blocks, Cout / NCO).
template <typename T, int NCO>
__global__ void __launch_bounds__(128) conv3x3_direct_kernel(Conv3Params p) {
  extern __shared__ __align__(16) float smem[];
  const int K = 9 * p.Cin, co0 = blockIdx.y * NCO;
  const PwLayout L(K, p.Cout);
  const T* x = static_cast<const T*>(p.x);
  T* out = static_cast<T*>(p.out);
  pdl_launch_dependents();
  for (int e = threadIdx.x; e < K * NCO; e += blockDim.x) {
    const int k = e / NCO, j = e - k * NCO;
    smem[e] = __ldg(p.blob + L.o_w + (size_t)k * p.Cout + co0 + j);
  }
  pdl_wait();
  __syncthreads();
  const int ntok = p.B * p.Ho * p.Wo;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntok) return;
  int bb, rem, oy, ox;
  p.d_howo.divmod(t, bb, rem);
  p.d_wo.divmod(rem, oy, ox);
  // the nine taps: clamped row pointer + inside flag
  const T* px[9];
  bool in[9];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int iy = oy * p.stride - 1 + tap / 3, ix = ox * p.stride - 1 + tap % 3;
    in[tap] = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
    px[tap] = x + ((size_t)(bb * p.H + (in[tap] ? iy : 0)) * p.W + (in[tap] ? ix : 0)) * p.Cin;
  }
  float2 acc[NCO / 2];
#pragma unroll
  for (int j = 0; j < NCO / 2; ++j)
    acc[j] = __ldg(reinterpret_cast<const float2*>(p.blob + L.o_b + co0) + j);
  const int hc = p.Cin / 2;
#pragma unroll 1
  for (int c2 = 0; c2 < hc; ++c2) {
    float2 xv[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      xv[tap] = Pair<T>::ld(px[tap] + 2 * c2);
      if (!in[tap]) xv[tap] = make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const float2* w0 = reinterpret_cast<const float2*>(smem + (size_t)(tap * p.Cin + 2 * c2) * NCO);
      const float2* w1 = w0 + NCO / 2;
#pragma unroll
      for (int j = 0; j < NCO / 2; ++j) {
        acc[j] = __ffma2_rn(make_float2(xv[tap].x, xv[tap].x), w0[j], acc[j]);
        acc[j] = __ffma2_rn(make_float2(xv[tap].y, xv[tap].y), w1[j], acc[j]);
      }
    }
  }
  T* o = out + (size_t)t * p.Cout + co0;
#pragma unroll
  for (int j = 0; j < NCO / 2; ++j) {
    float2 v = acc[j];
    if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); }
    Pair<T>::st(o + 2 * j, v.x, v.y);
  }
}

template <typename T, int NCO>
static int launch_conv3x3_direct(const Conv3Params& p, cudaStream_t stream) {
  const size_t smem = (size_t)9 * p.Cin * NCO * sizeof(float);
  auto kern = conv3x3_direct_kernel<T, NCO>;
  HRF_CUDA(ensure_smem((const void*)kern, smem));
  const int ntok = p.B * p.Ho * p.Wo;
  HRF_CUDA(launch_pdl(kern, dim3(ceil_div(ntok, 128), p.Cout / NCO), dim3(128), smem, stream, p));
  count_launch();
  HRF_CUDA(cudaGetLastError());
  return HRF_OK;
}

// output size + index helpers; call before any conv3x3 launcher
static int conv3x3_prepare(Conv3Params& p) {
  HRF_REQUIRE(p.Cout % 2 == 0 && p.Cin % 2 == 0, HRF_EUNSUPPORTED,
              "conv3x3: Cin=%d / Cout=%d must be even", p.Cin, p.Cout);
  HRF_REQUIRE(p.stride == 1 || p.stride == 2, HRF_EUNSUPPORTED, "conv3x3: stride %d", p.stride);
  p.Ho = (p.H - 1) / p.stride + 1;
  p.Wo = (p.W - 1) / p.stride + 1;
  p.d_cin = FastDiv(p.Cin);
  p.d_wo = FastDiv(p.Wo);
  p.d_howo = FastDiv(p.Ho * p.Wo);
  return HRF_OK;
}

template <typename T>
static int launch_conv3x3(const Conv3Params& p, cudaStream_t stream) {
  const int ntok = p.B * p.Ho * p.Wo;
  if (p.Cin <= 160 && p.Cout / 6 <= 65535) {   // narrow: the direct kernel
    if (p.Cout % 6 == 0) return launch_conv3x3_direct<T, 6>(p, stream);
    if (p.Cout % 8 == 0) return launch_conv3x3_direct<T, 8>(p, stream);
  }
  return ntok >= 148 * 2 * 64 ? launch_conv3x3_tok<T, 64>(p, stream) : launch_conv3x3_tok<T, 16>(p, stream);
}

// ---------------------------------------------------------------------------
// depthwise 3x3 stride-2 (pad 1) + BN, then 1x1 + BN (+ReLU).
// blob = Wdw [9][Cin] (BN folded), bdw [c4(Cin)], then a PwLayout blob.
// ---------------------------------------------------------------------------
struct DwPwLayout {
  int Cin, Cout, o_wd, o_bd, o_pw, total;
  __host__ __device__ DwPwLayout(int cin, int cout) {
    Cin = cin; Cout = cout;
    o_wd = 0; o_bd = round_up(9 * cin, 4); o_pw = o_bd + round_up(cin, 4);
    total = o_pw + PwLayout(cin, cout).total;
  }
};
struct DwPwParams {
  const void* x; const float* blob; void* out;
  int B, H, W, Ho, Wo, Cin, Cout, relu;
  int w_smem;
  FastDiv d_wo, d_howo;   // filled by launch_dwpw
};

template <typename T, int CT, int TOK>
__global__ void __launch_bounds__(kPwThreads) dwpw_kernel(DwPwParams p) {
  extern __shared__ __align__(16) float smem[];
  const DwPwLayout D(p.Cin, p.Cout);
  const PwLayout L(p.Cin, p.Cout);
  const T* x = static_cast<const T*>(p.x);
  T* out = static_cast<T*>(p.out);
  const int ntok = p.B * p.Ho * p.Wo;
  const int t0 = blockIdx.x * TOK;
  const int m = min(TOK, ntok - t0);
  const float* wd = p.blob + D.o_wd;
  const float* bd = p.blob + D.o_bd;
  const float* pw = p.blob + D.o_pw;
  pdl_launch_dependents();
  const float* Wt = stage_weights(pw + L.o_w, smem + TOK * L.lda, L.Kp * p.Cout, p.w_smem);
  pdl_wait();
  const int hp = L.lda / 2;
  for (int e = threadIdx.x; e < TOK * hp; e += blockDim.x) {
    const int r = e / hp, c = (e - r * hp) * 2;
    float2 s = make_float2(0.f, 0.f);
    if (r < m && c < p.Cin) {
      int b, rem, oy, ox;
      p.d_howo.divmod(t0 + r, b, rem);
      p.d_wo.divmod(rem, oy, ox);
      s = __ldg(reinterpret_cast<const float2*>(bd + c));
      // all nine taps are loaded before the first FMA (clamped address, zero weight
      // outside the image) so the loads overlap instead of serialising on L2 latency
      float2 v[9], w[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const int iy = oy * 2 - 1 + k / 3, ix = ox * 2 - 1 + k % 3;
        const bool in = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
        const int cy = min(max(iy, 0), p.H - 1), cx = min(max(ix, 0), p.W - 1);
        v[k] = Pair<T>::ld(x + ((size_t)(b * p.H + cy) * p.W + cx) * p.Cin + c);
        w[k] = in ? __ldg(reinterpret_cast<const float2*>(wd + k * p.Cin + c)) : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        s.x = fmaf(v[k].x, w[k].x, s.x);
        s.y = fmaf(v[k].y, w[k].y, s.y);
      }
    }
    *reinterpret_cast<float2*>(smem + r * L.lda + c) = s;
  }
  __syncthreads();
  const float* bias = pw + L.o_b;
  const int relu = p.relu, Cout = p.Cout;
  block_gemm<4, CT>(smem, L.lda, m, Wt, L.Kp, Cout, [&](int r, int n, float v) {
    v += __ldg(bias + n);
    if (relu) v = fmaxf(v, 0.f);
    Elem<T>::st(out + (size_t)(t0 + r) * Cout + n, v);
  });
}

template <typename T, int TOK>
static int launch_dwpw_tok(DwPwParams p, cudaStream_t stream) {
  const PwLayout L(p.Cin, p.Cout);
  size_t smem = (size_t)TOK * L.lda * sizeof(float);
  p.w_smem = smem + (size_t)L.Kp * p.Cout * sizeof(float) <= 160 * 1024;
  if (p.w_smem) smem += (size_t)L.Kp * p.Cout * sizeof(float);
  HRF_REQUIRE(smem <= 227 * 1024, HRF_EUNSUPPORTED, "dwpw: Cin=%d too wide", p.Cin);
  auto kern = (p.Cout % 4 == 0) ? dwpw_kernel<T, 4, TOK> : dwpw_kernel<T, 2, TOK>;
  HRF_CUDA(ensure_smem((const void*)kern, smem));
  HRF_CUDA(launch_pdl(kern, dim3(ceil_div(p.B * p.Ho * p.Wo, TOK)), dim3(kPwThreads), smem, stream, p));
  count_launch();
  HRF_CUDA(cudaGetLastError());
  return HRF_OK;
}

template <typename T>
static int launch_dwpw(DwPwParams p, cudaStream_t stream) {
  HRF_REQUIRE(p.Cout % 2 == 0 && p.Cin % 2 == 0, HRF_EUNSUPPORTED,
              "dwpw: Cin=%d / Cout=%d must be even", p.Cin, p.Cout);
  p.d_wo = FastDiv(p.Wo);
  p.d_howo = FastDiv(p.Ho * p.Wo);
  return p.B * p.Ho * p.Wo >= 148 * 2 * 64 ? launch_dwpw_tok<T, 64>(p, stream)
                                           : launch_dwpw_tok<T, 16>(p, stream);
}

// ---------------------------------------------------------------------------
// out = ReLU(x + sum_j bilinear(up_j -> (H,W)) + sum_j same_j), optional second
// copy of the result as contiguous fp32 NCHW (the backbone's output contract).
// Bilinear: align_corners=False, scale = in/out from sizes, source index clamped
// at 0 (ATen upsample_bilinear2d semantics used by hrnet.py:199-203).
// ---------------------------------------------------------------------------
struct FuseParams {
  const void* x;
  const void* up[HRF_MAX_FUSE_TERMS];
  const void* same[HRF_MAX_FUSE_TERMS];
  int up_H[HRF_MAX_FUSE_TERMS], up_W[HRF_MAX_FUSE_TERMS];
  void* out;
  float* out_nchw;
  int B, H, W, C, n_up, n_same, relu;
  float up_sh[HRF_MAX_FUSE_TERMS], up_sw[HRF_MAX_FUSE_TERMS];   // filled by launch_fuse
  FastDiv d_hc, d_w, d_h, d_hw;
  int tokb, tokb_log2;                                           // NCHW variant: tokens per CTA
};

// values of PG consecutive channel pairs of one token: x + same-resolution terms + bilinear
// gathers.  The bilinear coordinates / weights are computed once per token and up-term, and
// the 4 * PG gathers of a term are independent loads.
template <typename T, int PG>
__device__ __forceinline__ void fuse_values(const FuseParams& p, int b, int h, int w, size_t off, int c,
                                            float2* v) {
#pragma unroll
  for (int k = 0; k < PG; ++k) v[k] = Pair<T>::ld(static_cast<const T*>(p.x) + off + 2 * k);
  for (int j = 0; j < p.n_same; ++j) {
#pragma unroll
    for (int k = 0; k < PG; ++k) {
      const float2 a = Pair<T>::ld(static_cast<const T*>(p.same[j]) + off + 2 * k);
      v[k].x += a.x; v[k].y += a.y;
    }
  }
  for (int j = 0; j < p.n_up; ++j) {
    const int ih = p.up_H[j], iw = p.up_W[j];
    const T* u = static_cast<const T*>(p.up[j]);
    // align_corners=False, scale = in / out (computed on the host exactly as float(in)/float(out))
    const float fy = fmaxf(p.up_sh[j] * ((float)h + 0.5f) - 0.5f, 0.f);
    const float fx = fmaxf(p.up_sw[j] * ((float)w + 0.5f) - 0.5f, 0.f);
    const int y0 = min((int)fy, ih - 1), x0 = min((int)fx, iw - 1);
    const int y1 = y0 + (y0 < ih - 1 ? 1 : 0), x1 = x0 + (x0 < iw - 1 ? 1 : 0);
    const float ly = fminf(fmaxf(fy - (float)y0, 0.f), 1.f), lx = fminf(fmaxf(fx - (float)x0, 0.f), 1.f);
    const float hy = 1.f - ly, hx = 1.f - lx;
    const T* ub = u + (size_t)b * ih * iw * p.C + c;
    const T* p00 = ub + (size_t)(y0 * iw + x0) * p.C;
    const T* p01 = ub + (size_t)(y0 * iw + x1) * p.C;
    const T* p10 = ub + (size_t)(y1 * iw + x0) * p.C;
    const T* p11 = ub + (size_t)(y1 * iw + x1) * p.C;
#pragma unroll
    for (int k = 0; k < PG; ++k) {
      const float2 v00 = Pair<T>::ld(p00 + 2 * k), v01 = Pair<T>::ld(p01 + 2 * k);
      const float2 v10 = Pair<T>::ld(p10 + 2 * k), v11 = Pair<T>::ld(p11 + 2 * k);
      v[k].x += hy * (hx * v00.x + lx * v01.x) + ly * (hx * v10.x + lx * v11.x);
      v[k].y += hy * (hx * v00.y + lx * v01.y) + ly * (hx * v10.y + lx * v11.y);
    }
  }
  if (p.relu) {
#pragma unroll
    for (int k = 0; k < PG; ++k) { v[k].x = fmaxf(v[k].x, 0.f); v[k].y = fmaxf(v[k].y, 0.f); }
  }
}

// thread = (token, group of PG channel pairs); 32-bit index math (tensors are < 2^31 elements)
template <typename T, int PG>
__global__ void __launch_bounds__(256) fuse_sum_kernel(FuseParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const int total = p.B * p.H * p.W * (p.C / 2 / PG);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    int t, g, hb, w, b, h;
    p.d_hc.divmod(e, t, g);                          // d_hc = groups per token
    p.d_w.divmod(t, hb, w);
    p.d_h.divmod(hb, b, h);
    const size_t off = (size_t)t * p.C + 2 * PG * g;
    float2 v[PG];
    fuse_values<T, PG>(p, b, h, w, off, 2 * PG * g, v);
#pragma unroll
    for (int k = 0; k < PG; ++k) Pair<T>::st(static_cast<T*>(p.out) + off + 2 * k, v[k].x, v[k].y);
  }
}

// Variant that also writes the fp32 NCHW copy (the backbone's four outputs): a CTA owns
// kFuseTok consecutive tokens, keeps the values (rounded through the storage type, so both
// copies agree) in shared memory [C][tokens] and writes every channel's run of tokens as one
// contiguous segment -- the direct form scatters 4-byte writes H*W*4 bytes apart.
template <typename T, int PG>
__global__ void __launch_bounds__(256) fuse_sum_nchw_kernel(FuseParams p) {
  extern __shared__ float sv[];                      // [C][kFuseTok + 1]
  pdl_launch_dependents();
  pdl_wait();
  const int kFuseTok = p.tokb;                       // tokens per CTA (a power of two)
  const int ng = p.C / 2 / PG, ntok = p.B * p.H * p.W, hw = p.H * p.W;
  const int t0 = blockIdx.x * kFuseTok;
  const int m = min(kFuseTok, ntok - t0);
  for (int e = threadIdx.x; e < m * ng; e += blockDim.x) {
    int r, g, hb, w, b, h;
    p.d_hc.divmod(e, r, g);
    const int t = t0 + r;
    p.d_w.divmod(t, hb, w);
    p.d_h.divmod(hb, b, h);
    const size_t off = (size_t)t * p.C + 2 * PG * g;
    float2 v[PG];
    fuse_values<T, PG>(p, b, h, w, off, 2 * PG * g, v);
#pragma unroll
    for (int k = 0; k < PG; ++k) {
      Pair<T>::st(static_cast<T*>(p.out) + off + 2 * k, v[k].x, v[k].y);
      const float2 vs = Pair<T>::rt(v[k]);           // as stored
      sv[(2 * (PG * g + k)) * (kFuseTok + 1) + r] = vs.x;
      sv[(2 * (PG * g + k) + 1) * (kFuseTok + 1) + r] = vs.y;
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < p.C * kFuseTok; e += blockDim.x) {
    const int c = e >> p.tokb_log2, r = e & (kFuseTok - 1);
    if (r < m) {
      const int t = t0 + r;
      const int b = p.d_hw.div(t), thw = t - b * hw;
      p.out_nchw[((size_t)b * p.C + c) * hw + thw] = sv[c * (kFuseTok + 1) + r];
    }
  }
}

template <typename T, int PG>
static int launch_fuse_pg(FuseParams p, cudaStream_t stream) {
  const int ng = p.C / 2 / PG;
  p.d_hc = FastDiv(ng);
  const size_t total = (size_t)p.B * p.H * p.W * ng;
  if (p.out_nchw) {
    // tokens per CTA: one (token, pair group) work item per thread, fewer for small maps so
    // that the grid still covers the SMs; a power of two >= 16
    const int ntok = p.B * p.H * p.W;
    int tokb = 16, lg = 4;
    while (tokb * 2 * ng <= 256 && ntok / (tokb * 2) >= 148) { tokb *= 2; ++lg; }
    p.tokb = tokb;
    p.tokb_log2 = lg;
    const size_t smem = (size_t)p.C * (tokb + 1) * sizeof(float);
    HRF_REQUIRE(smem <= 200 * 1024, HRF_EUNSUPPORTED, "fuse_sum: C=%d too wide for the NCHW copy", p.C);
    HRF_CUDA(ensure_smem((const void*)fuse_sum_nchw_kernel<T, PG>, smem));
    HRF_CUDA(launch_pdl(fuse_sum_nchw_kernel<T, PG>, dim3(ceil_div(ntok, tokb)), dim3(256), smem, stream, p));
  } else {
    // 128-thread CTAs: few threads in total (one per token and pair group), spread them
    const int grid = (int)((total + 127) / 128 < 148 * 32 ? (total + 127) / 128 : 148 * 32);
    HRF_CUDA(launch_pdl(fuse_sum_kernel<T, PG>, dim3(grid), dim3(128), 0, stream, p));
  }
  count_launch();
  HRF_CUDA(cudaGetLastError());
  return HRF_OK;
}

template <typename T>
static int launch_fuse(FuseParams p, cudaStream_t stream) {
  const size_t total = (size_t)p.B * p.H * p.W * (p.C / 2);
  HRF_REQUIRE(p.C % 2 == 0 && total * 2 < ((size_t)1 << 31), HRF_EUNSUPPORTED,
              "fuse_sum: C=%d must be even and the tensor below 2^31 elements", p.C);
  p.d_w = FastDiv(p.W);
  p.d_h = FastDiv(p.H);
  p.d_hw = FastDiv(p.H * p.W);
  for (int j = 0; j < p.n_up; ++j) {
    p.up_sh[j] = (float)p.up_H[j] / (float)p.H;
    p.up_sw[j] = (float)p.up_W[j] / (float)p.W;
  }
  const int hc = p.C / 2;
  if (hc % 9 == 0) return launch_fuse_pg<T, 9>(p, stream);     // HRFuser-T widths 18k
  if (hc % 13 == 0) return launch_fuse_pg<T, 13>(p, stream);   // HRFuser-B widths 78k
  if (hc % 8 == 0) return launch_fuse_pg<T, 8>(p, stream);
  return launch_fuse_pg<T, 1>(p, stream);
}

// ---------------------------------------------------------------------------
// layout converters (tiled transpose through shared memory)
// ---------------------------------------------------------------------------
template <typename TS, typename TD, bool ToNhwc>
__global__ void __launch_bounds__(256) layout_kernel(const TS* src, TD* dst, int C, int HW) {
  // ToNhwc: src [B][C][HW] -> dst [B][HW][C];   else src [B][HW][C] -> dst [B][C][HW]
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x / 32;  // 32 x 8
  const size_t base = (size_t)b * C * HW;
  if (ToNhwc) {
    for (int i = ty; i < 32; i += 8) {
      const int c = c0 + i, q = p0 + tx;
      tile[i][tx] = (c < C && q < HW) ? Elem<TS>::ld(src + base + (size_t)c * HW + q) : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
      const int q = p0 + i, c = c0 + tx;
      if (c < C && q < HW) Elem<TD>::st(dst + base + (size_t)q * C + c, tile[tx][i]);
    }
  } else {
    for (int i = ty; i < 32; i += 8) {
      const int q = p0 + i, c = c0 + tx;
      tile[i][tx] = (c < C && q < HW) ? Elem<TS>::ld(src + base + (size_t)q * C + c) : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
      const int c = c0 + i, q = p0 + tx;
      if (c < C && q < HW) Elem<TD>::st(dst + base + (size_t)c * HW + q, tile[tx][i]);
    }
  }
}

template <bool ToNhwc>
static int launch_layout(int B, int C, int H, int W, int sdt, const void* src, int ddt, void* dst,
                         cudaStream_t stream) {
  const int HW = H * W;
  dim3 grid(ceil_div(HW, 32), ceil_div(C, 32), B);
  HRF_REQUIRE(grid.y <= 65535 && grid.z <= 65535, HRF_EUNSUPPORTED, "layout: tensor too large");
  using bf = __nv_bfloat16;
  if (sdt == HRF_F32 && ddt == HRF_F32)
    layout_kernel<float, float, ToNhwc><<<grid, 256, 0, stream>>>((const float*)src, (float*)dst, C, HW);
  else if (sdt == HRF_F32 && ddt == HRF_BF16)
    layout_kernel<float, bf, ToNhwc><<<grid, 256, 0, stream>>>((const float*)src, (bf*)dst, C, HW);
  else if (sdt == HRF_BF16 && ddt == HRF_F32)
    layout_kernel<bf, float, ToNhwc><<<grid, 256, 0, stream>>>((const bf*)src, (float*)dst, C, HW);
  else if (sdt == HRF_BF16 && ddt == HRF_BF16)
    layout_kernel<bf, bf, ToNhwc><<<grid, 256, 0, stream>>>((const bf*)src, (bf*)dst, C, HW);
  else
    HRF_REQUIRE(false, HRF_EINVAL, "layout: bad dtype");
  count_launch();
  HRF_CUDA(cudaGetLastError());
  return HRF_OK;
}

}  // namespace hrf

// ---------------------------------------------------------------------------
// Conv epilogue for the cuDNN-side layers (stems / Bottlenecks / transitions):
//   y = act(y + bias[c] (+ residual)),  in place on a channels-last tensor.
// One pass instead of the bias-add, residual-add and ReLU passes torch issues
// after a convolution (the BN affine is already folded into the conv weights).
// ---------------------------------------------------------------------------
namespace hrf {

template <typename T, int VEC>
__global__ void __launch_bounds__(256) bias_act_kernel(T* y, const float* __restrict__ bias,
                                                       const T* __restrict__ res, size_t n_vec,
                                                       int C, int relu) {
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_vec;
       v += (size_t)gridDim.x * blockDim.x) {
    const size_t e0 = v * VEC;
    const int c0 = (int)(e0 % (size_t)C);
    float f[VEC];
    if constexpr (sizeof(T) * VEC == 16) {
      const uint4 u = *reinterpret_cast<const uint4*>(y + e0);
      const T* p = reinterpret_cast<const T*>(&u);
#pragma unroll
      for (int i = 0; i < VEC; ++i) f[i] = (float)p[i];
      if (res) {
        const uint4 r = __ldg(reinterpret_cast<const uint4*>(res + e0));
        const T* q = reinterpret_cast<const T*>(&r);
#pragma unroll
        for (int i = 0; i < VEC; ++i) f[i] += (float)q[i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) f[i] = (float)y[e0 + i] + (res ? (float)res[e0 + i] : 0.f);
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      f[i] += __ldg(bias + c0 + i);
      if (relu) f[i] = fmaxf(f[i], 0.f);
    }
    if constexpr (sizeof(T) * VEC == 16) {
      uint4 u;
      T* p = reinterpret_cast<T*>(&u);
#pragma unroll
      for (int i = 0; i < VEC; ++i) p[i] = (T)f[i];
      *reinterpret_cast<uint4*>(y + e0) = u;
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) y[e0 + i] = (T)f[i];
    }
  }
}

template <typename T>
static int launch_bias_act(void* y, const float* bias, const void* res, size_t n_tok, int C,
                           int relu, cudaStream_t stream) {
  constexpr int V16 = 16 / (int)sizeof(T);
  const size_t total = n_tok * (size_t)C;
  const bool wide = (C % V16 == 0) && ((uintptr_t)y % 16 == 0) && (!res || (uintptr_t)res % 16 == 0);
  const int vec = wide ? V16 : (C % 2 == 0 ? 2 : 1);
  const size_t n_vec = total / vec;
  const int grid = (int)((n_vec + 255) / 256 < 148 * 16 ? (n_vec + 255) / 256 : 148 * 16);
  if (vec == V16)
    bias_act_kernel<T, V16><<<grid, 256, 0, stream>>>((T*)y, bias, (const T*)res, n_vec, C, relu);
  else if (vec == 2)
    bias_act_kernel<T, 2><<<grid, 256, 0, stream>>>((T*)y, bias, (const T*)res, n_vec, C, relu);
  else
    bias_act_kernel<T, 1><<<grid, 256, 0, stream>>>((T*)y, bias, (const T*)res, n_vec, C, relu);
  count_launch();
  HRF_CUDA(cudaGetLastError());
  return HRF_OK;
}

}  // namespace hrf
