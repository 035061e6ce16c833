// Shared device/host helpers for the hrfuser_b200 kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hrfuser_b200.h"

namespace hrf {

// -DHRF_KERNEL_PROFILE (tools/gpu_phases.sh): thread 0 of every CTA accumulates clock64 deltas
// per phase of a kernel's tile loop into g_prof[cta][16] (slot 15 = tiles processed), read back
// by hrf_debug_prof().  Compiled out of the product library.
#ifdef HRF_KERNEL_PROFILE
__device__ unsigned long long g_prof[2048 * 16];
__device__ unsigned long long g_prof_cta[2048 * 4];      // per CTA: SM id, globaltimer start / end
__device__ __forceinline__ unsigned long long prof_gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void prof_cta_begin() {
  if (threadIdx.x == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    g_prof_cta[(blockIdx.x & 2047) * 4 + 0] = smid;
    g_prof_cta[(blockIdx.x & 2047) * 4 + 1] = prof_gtime();
  }
}
__device__ __forceinline__ void prof_cta_end() {
  if (threadIdx.x == 0) g_prof_cta[(blockIdx.x & 2047) * 4 + 2] = prof_gtime();
}
#define HRF_PROF_DECL long long pt_ = clock64(); prof_cta_begin();
#define HRF_PROF_END prof_cta_end();
#define HRF_PROF(k)                                                       \
  if (threadIdx.x == 0) {                                                 \
    const long long now_ = clock64();                                     \
    g_prof[(blockIdx.x & 2047) * 16 + (k)] += (unsigned long long)(now_ - pt_); \
    pt_ = now_;                                                           \
  }
#define HRF_PROF_TILE \
  if (threadIdx.x == 0) g_prof[(blockIdx.x & 2047) * 16 + 15] += 1;
#else
#define HRF_PROF_DECL
#define HRF_PROF_END
#define HRF_PROF(k)
#define HRF_PROF_TILE
#endif


constexpr int kWarp = 32;

// ---- error plumbing (thread-local message, never throws) --------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void count_launch(int n = 1);
bool tc_disabled();
bool pdl_enabled();
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per kernel (kept out of
// CUDA-graph capture after the first, warm-up, call)
cudaError_t ensure_smem(const void* kern, size_t bytes);

// ---- programmatic dependent launch ---------------------------------------------------
// Every hot kernel is launched with cudaLaunchAttributeProgrammaticStreamSerialization: it
// may start while its predecessor in the stream is still draining, runs its prologue (weights
// -> shared memory, TMEM allocation, barrier init: nothing that depends on the predecessor),
// and only then executes griddepcontrol.wait, which returns once the predecessor has completed
// and flushed its memory.  launch_dependents is issued first thing, so the successor's CTAs
// fill SM slots as soon as this grid's CTAs retire.  Inside a CUDA graph the edges become
// programmatic dependencies.  A kernel launched this way MUST call pdl_wait() before it touches
// anything an earlier kernel wrote.
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#define HRF_REQUIRE(cond, code, ...)          \
  do {                                        \
    if (!(cond)) {                            \
      ::hrf::set_error(__VA_ARGS__);          \
      return (code);                          \
    }                                         \
  } while (0)

#define HRF_CUDA(call)                                         \
  do {                                                         \
    cudaError_t e_ = (call);                                   \
    if (e_ != cudaSuccess) return ::hrf::cuda_fail(e_, #call); \
  } while (0)

// ---- small math ---------------------------------------------------------------
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int round_up(int a, int b) { return ceil_div(a, b) * b; }
// smallest 4*odd >= n: row strides of this form make 128-bit shared loads that
// walk rows conflict-free (bank group = row * odd mod 8).
__host__ __device__ inline int stride4odd(int n) {
  int q = ceil_div(n, 4);
  if ((q & 1) == 0) ++q;
  return q * 4;
}

// Division of a non-negative int (< 2^31) by a launch-constant divisor without the
// ~30-instruction integer-division sequence: q = umulhi(n, m) >> s (round-up method).
struct FastDiv {
  unsigned m, s, d;
  FastDiv() : m(0), s(0), d(1) {}
  explicit FastDiv(int div) : d((unsigned)div) {
    s = 0;
    while ((1u << s) < d) ++s;
    // m = ceil(2^(32+s) / d) - 2^32  (33-bit magic, "add" form)
    const unsigned long long t = (((1ull << s) - d) << 32) / d + 1;
    m = (unsigned)t;
  }
  __device__ __forceinline__ int div(int n) const {
    const unsigned q = __umulhi((unsigned)n, m);
    return (int)((q + (unsigned)n) >> s);      // n < 2^31: q + n cannot overflow
  }
  __device__ __forceinline__ void divmod(int n, int& q, int& r) const {
    q = div(n);
    r = n - q * (int)d;
  }
};

// ---- activation element access (storage type T, math in fp32) -----------------
template <typename T> struct Elem;
template <> struct Elem<float> {
  __device__ static float ld(const float* p) { return __ldg(p); }
  __device__ static void st(float* p, float v) { *p = v; }
};
template <> struct Elem<__nv_bfloat16> {
  __device__ static float ld(const __nv_bfloat16* p) {
    return __bfloat162float(*p);
  }
  __device__ static void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

__device__ inline float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ inline float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// exact (erf) GELU, as nn.GELU() (hrformer.py:270,279,282 via build_activation_layer)
__device__ inline float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

// LayerNorm of one token by one warp: `ld(c)` returns channel c.  Writes
// dst[c] = (x-mean)*rstd*gamma[c]+beta[c] for c < C and zeros for C <= c < Cpad.
// Two-pass variance like torch's native_layer_norm (biased, eps inside sqrt).
template <typename LoadFn>
__device__ inline void warp_layernorm(LoadFn ld, int C, int Cpad, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, float eps, float* dst) {
  const int lane = threadIdx.x & 31;
  constexpr int kMax = 8;  // supports C <= 256
  float v[kMax];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kMax; ++i) {
    int c = lane + i * 32;
    v[i] = (c < C) ? ld(c) : 0.f;
    s += v[i];
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kMax; ++i) {
    int c = lane + i * 32;
    float d = (c < C) ? v[i] - mean : 0.f;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
  for (int i = 0; i < kMax; ++i) {
    int c = lane + i * 32;
    if (c < C)
      dst[c] = (v[i] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
    else if (c < Cpad)
      dst[c] = 0.f;
  }
}

// ---- block-level register-tiled GEMM on a shared-memory A operand -------------
// out(r, n) = sum_{k<K} A[r*lda + k] * Wt[k*N + n],  r < M, n < N.
// A: fp32 in shared memory, K a multiple of 4 (caller zero-pads), rows 16-byte
// aligned.  Wt: fp32, k-major, in shared memory (staged by the caller: one coalesced
// copy instead of K/4 serialised L2 round trips) or in global memory (read through L1).  Each thread owns RT x CT output tiles; consecutive
// threads take consecutive column tiles so Wt loads coalesce and A loads
// broadcast.  epi(r, n, value) is called for every valid output element.
template <int RT, int CT, typename Epi>
__device__ inline void block_gemm(const float* A, int lda, int M, const float* Wt,
                                  int K, int N, Epi epi) {
  static_assert(CT == 2 || CT == 4, "CT");
  const int ntc = N / CT;
  const int ntr = ceil_div(M, RT);
  for (int t = threadIdx.x; t < ntr * ntc; t += blockDim.x) {
    const int n0 = (t % ntc) * CT;
    const int r0 = (t / ntc) * RT;
    float acc[RT][CT];
#pragma unroll
    for (int r = 0; r < RT; ++r)
#pragma unroll
      for (int c = 0; c < CT; ++c) acc[r][c] = 0.f;
    const float* arow[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) arow[r] = A + (size_t)min(r0 + r, M - 1) * lda;
    for (int k = 0; k < K; k += 4) {
      float4 a[RT];
#pragma unroll
      for (int r = 0; r < RT; ++r) a[r] = *reinterpret_cast<const float4*>(arow[r] + k);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float w[CT];
        const float* wp = Wt + (size_t)(k + kk) * N + n0;
        if constexpr (CT == 4) {
          float4 w4 = *reinterpret_cast<const float4*>(wp);     // generic: global or shared
          w[0] = w4.x; w[1] = w4.y; w[2] = w4.z; w[3] = w4.w;
        } else {
          float2 w2 = *reinterpret_cast<const float2*>(wp);
          w[0] = w2.x; w[1] = w2.y;
        }
#pragma unroll
        for (int r = 0; r < RT; ++r) {
          const float av = kk == 0 ? a[r].x : kk == 1 ? a[r].y : kk == 2 ? a[r].z : a[r].w;
#pragma unroll
          for (int c = 0; c < CT; ++c) acc[r][c] = fmaf(av, w[c], acc[r][c]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RT; ++r)
      if (r0 + r < M) {
#pragma unroll
        for (int c = 0; c < CT; ++c) epi(r0 + r, n0 + c, acc[r][c]);
      }
  }
}

}  // namespace hrf
