// Train-mode BatchNorm / SyncBatchNorm on NCHW planes: the per-channel
// batch statistics the reference gets from nn.BatchNorm2d / nn.SyncBatchNorm
// (mmcv build_norm_layer; hrnet.py:338-358, hrformer.py:267-282, resnet.py:161-206) and the
// per-channel affine passes of its forward and backward.
//
//   bn_reduce_kernel<T, VEC, BWD>   one CTA per (plane, chunk of 256 * 4 * VEC elements):
//        BWD = false: chunk count / mean / M2 (the chunk stays in registers between the sum and
//                     the centred second moment, so E[x^2] - mean^2 cancellation never happens
//                     in fp32)
//        BWD = true : sum(dy), sum(dy * (x - mean) * invstd)
//        -> one float2 partial per CTA, [b][chunk][c]
//   bn_finalize_kernel<BWD>         one warp per channel: fixed-order combination of the
//        partials in fp64 (Chan's update for the forward) -> sums[0..C) , sums[C..2C) (fp64,
//        additive across ranks: sum x | sum x^2, or sum dy | sum dy * xhat)
//   bn_affine_kernel<T, VEC, TWO, MODE>  out = a[c] * x + c0[c]             (forward normalise)
//                                   out = a[c] * dy + b[c] * x + c0[c] (backward dx)
//        with a / b / c0 given, or derived inside the kernel from the reduced statistics
//        (COEF_FWD also saves mean / invstd and updates the running statistics), so that a BN
//        forward is 3 launches and nothing else on the host, a backward 3 more
//
// Fused activation (`act`: ReLU or exact GELU, the nn.ReLU / nn.GELU that follows the norm layer
// at hrformer.py:267-282, hrnet.py:338-358, resnet.py:161-206): the forward applies it to the
// normalised value in the same pass; the backward recomputes z = w * xhat + b from x and the
// saved statistics and multiplies dy by act'(z) on the fly in both of its passes, so neither
// the pre-activation tensor nor a separate activation pass ever exists in HBM.
//
// All three are pure streaming kernels: 16-byte loads, four in flight per thread, no reuse.
// Deterministic: no atomics; the partial order is fixed by (b, chunk).
#pragma once
#include "common.cuh"

namespace hrf {

constexpr int kBnThreads = 256;
constexpr int kBnVecPerThread = 4;

enum { BN_ACT_NONE = 0, BN_ACT_RELU = 1, BN_ACT_GELU = 2 };
__device__ __forceinline__ float bn_act(float z, int act) {
  if (act == BN_ACT_RELU) return fmaxf(z, 0.f);
  if (act == BN_ACT_GELU) return 0.5f * z * (1.f + erff(z * 0.70710678118654752f));
  return z;
}
// d act(z) / dz
__device__ __forceinline__ float bn_act_grad(float z, int act) {
  if (act == BN_ACT_RELU) return z > 0.f ? 1.f : 0.f;
  if (act == BN_ACT_GELU)
    return 0.5f * (1.f + erff(z * 0.70710678118654752f)) + z * 0.39894228040143268f * __expf(-0.5f * z * z);
  return 1.f;
}

template <typename T>
__device__ __forceinline__ float bn_to_float(T v) { return (float)v; }

template <typename T, int VEC>
__device__ __forceinline__ void bn_load(const T* p, float (&f)[VEC]) {
  if constexpr (VEC == 1) {
    f[0] = bn_to_float(__ldg(p));
  } else {
    static_assert(sizeof(T) * VEC == 16, "16-byte vectors");
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const T* q = reinterpret_cast<const T*>(&u);
#pragma unroll
    for (int i = 0; i < VEC; ++i) f[i] = bn_to_float(q[i]);
  }
}
template <typename T, int VEC>
__device__ __forceinline__ void bn_store(T* p, const float (&f)[VEC]) {
  if constexpr (VEC == 1) {
    p[0] = (T)f[0];
  } else {
    uint4 u;
    T* q = reinterpret_cast<T*>(&u);
#pragma unroll
    for (int i = 0; i < VEC; ++i) q[i] = (T)f[i];
    *reinterpret_cast<uint4*>(p) = u;
  }
}

// block-wide sum of two floats; result valid in every thread
__device__ __forceinline__ float2 bn_block_sum2(float a, float b, float2* red /*[8]*/) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int w = threadIdx.x >> 5;
  __syncthreads();                      // red may still be read from a previous call
  if ((threadIdx.x & 31) == 0) red[w] = make_float2(a, b);
  __syncthreads();
  float2 s = red[0];
#pragma unroll
  for (int i = 1; i < kBnThreads / 32; ++i) {
    s.x += red[i].x;
    s.y += red[i].y;
  }
  return s;
}

template <typename T, int VEC, bool BWD>
__global__ void __launch_bounds__(kBnThreads)
bn_reduce_kernel(const T* __restrict__ x, const T* __restrict__ dy, const float* __restrict__ mean,
                 const float* __restrict__ invstd, const float* __restrict__ weight,
                 const float* __restrict__ bias, int act, float2* __restrict__ part, int C, int HW,
                 FastDiv chunks_div) {
  constexpr int NV = kBnVecPerThread;
  constexpr int CH = kBnThreads * NV * VEC;
  __shared__ float2 red[kBnThreads / 32];
  const int chunks = (int)chunks_div.d;
  const int plane = chunks_div.div((int)blockIdx.x);
  const int chunk = (int)blockIdx.x - plane * chunks;
  const int b = plane / C, c = plane - b * C;
  const int e0 = chunk * CH;
  const int n = min(CH, HW - e0);
  const size_t base = (size_t)plane * HW + e0;

  float v[NV][VEC];
  float s0 = 0.f, s1 = 0.f;
  if constexpr (!BWD) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int o = (i * kBnThreads + (int)threadIdx.x) * VEC;
      if (o < n) {
        bn_load<T, VEC>(x + base + o, v[i]);
#pragma unroll
        for (int j = 0; j < VEC; ++j) s0 += v[i][j];
      }
    }
    const float m = bn_block_sum2(s0, 0.f, red).x / (float)n;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int o = (i * kBnThreads + (int)threadIdx.x) * VEC;
      if (o < n) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const float d = v[i][j] - m;
          s1 = fmaf(d, d, s1);
        }
      }
    }
    const float m2 = bn_block_sum2(s1, 0.f, red).x;
    if (threadIdx.x == 0) part[((size_t)b * chunks + chunk) * C + c] = make_float2(m, m2);
  } else {
    const float mu = __ldg(mean + c), is = __ldg(invstd + c);
    const float za = (weight ? __ldg(weight + c) : 1.f) * is, zb = bias ? __ldg(bias + c) : 0.f;
    float g[NV][VEC];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int o = (i * kBnThreads + (int)threadIdx.x) * VEC;
      if (o < n) {
        bn_load<T, VEC>(x + base + o, v[i]);
        bn_load<T, VEC>(dy + base + o, g[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int o = (i * kBnThreads + (int)threadIdx.x) * VEC;
      if (o < n) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const float xc = v[i][j] - mu;
          const float ge = act ? g[i][j] * bn_act_grad(fmaf(za, xc, zb), act) : g[i][j];
          s0 += ge;
          s1 = fmaf(ge, xc * is, s1);
        }
      }
    }
    const float2 s = bn_block_sum2(s0, s1, red);
    if (threadIdx.x == 0) part[((size_t)b * chunks + chunk) * C + c] = s;
  }
}

// One warp per channel.  Lane l combines partials l, l + 32, ... in order, then the 32 lane
// results are combined by a fixed shuffle tree: the result depends only on the shapes.
template <bool BWD>
__global__ void __launch_bounds__(128)
bn_finalize_kernel(const float2* __restrict__ part, double* __restrict__ sums, int C, int HW,
                   int n_part /* B * chunks */, int chunks, int chunk_elems,
                   float* __restrict__ dweight, float* __restrict__ dbias) {
  const int c = (int)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
  if (c >= C) return;
  const int lane = threadIdx.x & 31;
  if constexpr (BWD) {
    double a = 0.0, b = 0.0;
    for (int p = lane; p < n_part; p += 32) {
      const float2 v = part[(size_t)p * C + c];
      a += (double)v.x;
      b += (double)v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) {
      sums[c] = a;
      sums[C + c] = b;
      if (dbias) dbias[c] = (float)a;        // rank-local parameter gradients
      if (dweight) dweight[c] = (float)b;
    }
  } else {
    double n = 0.0, mean = 0.0, m2 = 0.0;
    for (int p = lane; p < n_part; p += 32) {
      const int chunk = p % chunks;
      const double ni = (double)min(chunk_elems, HW - chunk * chunk_elems);
      const float2 v = part[(size_t)p * C + c];
      const double d = (double)v.x - mean, nn = n + ni;
      mean += d * ni / nn;
      m2 += (double)v.y + d * d * n * ni / nn;
      n = nn;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double n2 = __shfl_xor_sync(0xffffffffu, n, o);
      const double mean2 = __shfl_xor_sync(0xffffffffu, mean, o);
      const double m22 = __shfl_xor_sync(0xffffffffu, m2, o);
      const double nn = n + n2;
      if (nn > 0.0) {
        // symmetric in (this, other): both lanes of a pair compute the same value
        const double d = mean2 - mean;
        const double mnew = (n * mean + n2 * mean2) / nn;
        m2 = m2 + m22 + d * d * n * n2 / nn;
        mean = mnew;
        n = nn;
      }
    }
    if (lane == 0) {
      sums[c] = n * mean;
      sums[C + c] = m2 + n * mean * mean;
      if (c == 0) sums[2 * C] = n;           // element count: the third part of the SyncBN message
    }
  }
}

// Where the per-channel coefficients of bn_affine_kernel come from.
//   COEF_GIVEN: a / b / c0 arrays.
//   COEF_FWD  : the (all-reduced) statistics message [sum x | sum x^2 | count] (fp64) + weight,
//               bias, eps: y = (x - mean) * invstd * w + b.  The CTA of plane c, chunk 0 also
//               stores mean / invstd for the backward and updates the running statistics.
//   COEF_BWD  : the (all-reduced) [sum dy | sum dy * xhat] (fp64) + count, weight, mean, invstd:
//               dx = w * invstd * (dy - mean(dy) - xhat * mean(dy * xhat)).
// One thread per CTA evaluates the handful of fp64 operations; the rest wait at a barrier with
// their loads already in flight.
enum { COEF_GIVEN = 0, COEF_FWD = 1, COEF_BWD = 2 };
struct BnCoef {
  const float* a;
  const float* b;
  const float* c0;
  const double* sums;      // fwd: 2C + 1 (count last); bwd: 2C
  const double* count;     // bwd: element count (device scalar)
  const float* weight;     // may be NULL (1)
  const float* bias;       // may be NULL (0)
  const float* mean;       // bwd: saved
  const float* invstd;     // bwd: saved
  float* save_mean;        // fwd: out
  float* save_invstd;      // fwd: out
  float* running_mean;     // fwd: in / out, may be NULL
  float* running_var;
  float eps, momentum;
};

template <typename T, int VEC, bool TWO, int MODE>
__global__ void __launch_bounds__(kBnThreads)
bn_affine_kernel(const T* __restrict__ x, const T* __restrict__ dy, const BnCoef k,
                 T* __restrict__ out, int C, int HW, FastDiv chunks_div, int act) {
  static_assert(MODE != COEF_BWD || TWO, "the backward reads x and dy");
  static_assert(MODE != COEF_FWD || !TWO, "the forward reads x only");
  constexpr int NV = kBnVecPerThread;
  constexpr int CH = kBnThreads * NV * VEC;
  const int chunks = (int)chunks_div.d;
  const int plane = chunks_div.div((int)blockIdx.x);
  const int chunk = (int)blockIdx.x - plane * chunks;
  const int c = plane % C;
  const int e0 = chunk * CH;
  const int n = min(CH, HW - e0);
  const size_t base = (size_t)plane * HW + e0;
  float v[NV][VEC], g[NV][VEC];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int o = (i * kBnThreads + (int)threadIdx.x) * VEC;
    if (o < n) {
      bn_load<T, VEC>(x + base + o, v[i]);
      if constexpr (TWO) bn_load<T, VEC>(dy + base + o, g[i]);
    }
  }
  // out = ka * (TWO ? dy : x - km) + kb * (x - km) + kc: centring x before the multiply keeps
  // the result exact to fp32 rounding when |mean| >> std
  float ka, kb = 0.f, kc, km = 0.f, kz = 0.f;
  if constexpr (MODE == COEF_GIVEN) {
    ka = __ldg(k.a + c);
    kc = __ldg(k.c0 + c);
    if constexpr (TWO) kb = __ldg(k.b + c);
  } else {
    __shared__ float coef[5];
    if (threadIdx.x == 0) {
      const double w = k.weight ? (double)__ldg(k.weight + c) : 1.0;
      if constexpr (MODE == COEF_FWD) {
        const double cnt = k.sums[2 * C];
        const double mean = k.sums[c] / cnt;
        const double var = fmax(k.sums[C + c] / cnt - mean * mean, 0.0);
        const double invstd = 1.0 / sqrt(var + (double)k.eps);
        const double a = w * invstd;
        coef[0] = (float)a;
        coef[1] = 0.f;
        coef[2] = k.bias ? __ldg(k.bias + c) : 0.f;
        coef[3] = (float)mean;
        if (plane == c && chunk == 0) {            // b == 0: once per channel
          k.save_mean[c] = (float)mean;
          k.save_invstd[c] = (float)invstd;
          if (k.running_mean) {
            const double m = (double)k.momentum;
            const double unbiased = var * (cnt / fmax(cnt - 1.0, 1.0));
            k.running_mean[c] = (float)((1.0 - m) * (double)k.running_mean[c] + m * mean);
            k.running_var[c] = (float)((1.0 - m) * (double)k.running_var[c] + m * unbiased);
          }
        }
      } else {
        const double cnt = *k.count;
        const double mean = (double)__ldg(k.mean + c), invstd = (double)__ldg(k.invstd + c);
        const double mdy = k.sums[c] / cnt, mdyx = k.sums[C + c] / cnt;
        const double gs = w * invstd;
        coef[0] = (float)gs;
        coef[1] = (float)(-gs * invstd * mdyx);
        coef[2] = (float)(-gs * mdy);
        coef[3] = (float)mean;
        coef[4] = k.bias ? __ldg(k.bias + c) : 0.f;     // z = gs * (x - mean) + bias
      }
    }
    __syncthreads();
    ka = coef[0];
    kb = coef[1];
    kc = coef[2];
    km = coef[3];
    if constexpr (MODE == COEF_BWD) kz = coef[4];
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int o = (i * kBnThreads + (int)threadIdx.x) * VEC;
    if (o < n) {
      float r[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const float xc = v[i][j] - km;
        if constexpr (MODE == COEF_BWD) {
          const float ge = act ? g[i][j] * bn_act_grad(fmaf(ka, xc, kz), act) : g[i][j];
          r[j] = fmaf(ka, ge, fmaf(kb, xc, kc));
        } else {
          if constexpr (TWO) r[j] = fmaf(ka, g[i][j], fmaf(kb, xc, kc));
          else r[j] = fmaf(ka, xc, kc);
          if (act) r[j] = bn_act(r[j], act);
        }
      }
      bn_store<T, VEC>(out + base + o, r);
    }
  }
}

// ---- host side ---------------------------------------------------------------------------
struct BnGeom {
  int vec, chunk_elems, chunks;
  long long blocks;
};
template <typename T>
static BnGeom bn_geom(int B, int C, int HW, const void* p0, const void* p1, const void* p2) {
  constexpr int V16 = 16 / (int)sizeof(T);
  auto al = [](const void* p) { return !p || (uintptr_t)p % 16 == 0; };
  BnGeom g;
  g.vec = (HW % V16 == 0 && al(p0) && al(p1) && al(p2)) ? V16 : 1;
  g.chunk_elems = kBnThreads * kBnVecPerThread * g.vec;
  g.chunks = ceil_div(HW, g.chunk_elems);
  g.blocks = (long long)B * C * g.chunks;
  return g;
}
// partials are laid out for the widest chunking (vec = 1 gives the most chunks)
static size_t bn_workspace_bytes(int B, int C, int HW) {
  const int chunks = ceil_div(HW, kBnThreads * kBnVecPerThread);
  return (size_t)B * chunks * C * sizeof(float2);
}

template <typename T, bool BWD>
static int launch_bn_reduce(int B, int C, int HW, const void* x, const void* dy, const float* mean,
                            const float* invstd, const float* weight, const float* bias, int act,
                            double* sums, float* dweight, float* dbias, void* ws, cudaStream_t stream) {
  const BnGeom g = bn_geom<T>(B, C, HW, x, dy, nullptr);
  HRF_REQUIRE(g.blocks < (1ll << 31), HRF_EUNSUPPORTED, "bn: %lld blocks", g.blocks);
  const FastDiv cd(g.chunks);
  float2* part = reinterpret_cast<float2*>(ws);
  if (g.vec > 1)
    bn_reduce_kernel<T, 16 / (int)sizeof(T), BWD><<<(unsigned)g.blocks, kBnThreads, 0, stream>>>(
        (const T*)x, (const T*)dy, mean, invstd, weight, bias, act, part, C, HW, cd);
  else
    bn_reduce_kernel<T, 1, BWD><<<(unsigned)g.blocks, kBnThreads, 0, stream>>>(
        (const T*)x, (const T*)dy, mean, invstd, weight, bias, act, part, C, HW, cd);
  HRF_CUDA(cudaGetLastError());
  bn_finalize_kernel<BWD><<<ceil_div(C, 4), 128, 0, stream>>>(part, sums, C, HW, B * g.chunks,
                                                               g.chunks, g.chunk_elems, dweight, dbias);
  count_launch(2);
  HRF_CUDA(cudaGetLastError());
  return HRF_OK;
}

template <typename T, int MODE>
static int launch_bn_affine(int B, int C, int HW, const void* x, const void* dy, const BnCoef& k,
                            int act, void* out, cudaStream_t stream) {
  const BnGeom g = bn_geom<T>(B, C, HW, x, dy, out);
  HRF_REQUIRE(g.blocks < (1ll << 31), HRF_EUNSUPPORTED, "bn: %lld blocks", g.blocks);
  const FastDiv cd(g.chunks);
  constexpr int V16 = 16 / (int)sizeof(T);
  const unsigned grid = (unsigned)g.blocks;
  const T* xp = (const T*)x;
  const T* dp = (const T*)dy;
  T* op = (T*)out;
  if constexpr (MODE != COEF_FWD) {
    if (dy) {
      if (g.vec > 1)
        bn_affine_kernel<T, V16, true, MODE><<<grid, kBnThreads, 0, stream>>>(xp, dp, k, op, C, HW, cd, act);
      else
        bn_affine_kernel<T, 1, true, MODE><<<grid, kBnThreads, 0, stream>>>(xp, dp, k, op, C, HW, cd, act);
      count_launch();
      HRF_CUDA(cudaGetLastError());
      return HRF_OK;
    }
  }
  if constexpr (MODE != COEF_BWD) {
    if (g.vec > 1)
      bn_affine_kernel<T, V16, false, MODE><<<grid, kBnThreads, 0, stream>>>(xp, nullptr, k, op, C, HW, cd, act);
    else
      bn_affine_kernel<T, 1, false, MODE><<<grid, kBnThreads, 0, stream>>>(xp, nullptr, k, op, C, HW, cd, act);
    count_launch();
    HRF_CUDA(cudaGetLastError());
  }
  return HRF_OK;
}

}  // namespace hrf
