// C-ABI of libhrfuser_b200.so (see include/hrfuser_b200.h).
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "hrfuse.cuh"
#include "mixffn.cuh"
#include "umma_selftest.cuh"
#include "window_attn_tc.cuh"
#include "window_attn_v3.cuh"
#include "mixffn_tc.cuh"
#include "mixffn_v2.cuh"
#include "conv3x3_tc.cuh"
#include "conv_gemm_tc.cuh"
#include "ln_train.cuh"
#include "attn_train.cuh"
#include "dwconv_train.cuh"
#include "pool.cuh"
#include "stem_conv_tc.cuh"
#include "generic.cuh"
#include "window_attn.cuh"
#include "bn_train.cuh"
#include "input_prologue.cuh"

namespace hrf {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return HRF_ECUDA;
}
cudaError_t ensure_smem(const void* kern, size_t bytes) {
  static std::mutex mu;
  // cudaFuncSetAttribute applies to the CURRENT device: key the cache on (device, kernel)
  static std::map<std::pair<int, const void*>, size_t> done;
  std::lock_guard<std::mutex> lk(mu);
  int dev = 0;
  if (cudaError_t de = cudaGetDevice(&dev); de != cudaSuccess) return de;
  const std::pair<int, const void*> key(dev, kern);
  auto it = done.find(key);
  if (it != done.end() && it->second >= bytes) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  // ask for the largest shared-memory carve-out: the driver's default split sizes it for ONE
  // CTA of the kernel, which can leave room for fewer co-resident CTAs than the kernel was
  // designed for (HRF_CARVEOUT=0 keeps the driver default, for A/B runs)
  static const bool carve = [] { const char* v = std::getenv("HRF_CARVEOUT"); return !(v && v[0] == '0'); }();
  if (e == cudaSuccess && carve)
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e == cudaSuccess) done[key] = bytes;
  // HRF_DEBUG_OCC=<threads>: print the resident CTAs per SM the runtime predicts for this kernel
  static const int occ_threads = [] { const char* v = std::getenv("HRF_DEBUG_OCC"); return v ? std::atoi(v) : 0; }();
  if (e == cudaSuccess && occ_threads > 0) {
    int nb = -1;
    cudaFuncAttributes fa{};
    cudaFuncGetAttributes(&fa, kern);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, fa.maxThreadsPerBlock < occ_threads ? fa.maxThreadsPerBlock : occ_threads, bytes);
    std::fprintf(stderr, "[hrf occ] kernel %p regs %d static smem %zu dyn smem %zu maxThreads %d -> %d CTAs/SM (at %d threads)\n",
                 kern, fa.numRegs, fa.sharedSizeBytes, bytes, fa.maxThreadsPerBlock, nb,
                 fa.maxThreadsPerBlock < occ_threads ? fa.maxThreadsPerBlock : occ_threads);
  }
  return e;
}
// HRF_DISABLE_TC=1 routes bf16 problems to the SIMT kernels (A/B comparisons)
bool tc_disabled() {
  static const bool off = [] {
    const char* e = std::getenv("HRF_DISABLE_TC");
    return e && e[0] == '1';
  }();
  return off;
}
// Programmatic dependent launch: off unless HRF_PDL=1 or hrf_set_pdl(1).  Measured on B200: a
// single lsa -> mixffn chain in a graph gains 5 %, but the engine's multi-stream graph loses
// ~2 % (early-launched dependents hold SM slots and TMEM columns that kernels of the other
// streams could use), so the engine leaves it off.
static std::atomic<int> g_pdl{-1};
bool pdl_enabled() {
  int v = g_pdl.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = std::getenv("HRF_PDL");
    v = (e && std::atoi(e) != 0) ? 1 : 0;
    g_pdl.store(v, std::memory_order_relaxed);
  }
  return v != 0;
}
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

// round-to-nearest-even fp32 -> bf16 bits (finite inputs)
static inline uint16_t f32_to_bf16(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

// round-to-nearest-even fp32 -> fp16 bits, saturating to the largest finite value
static inline uint16_t f32_to_f16(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  const uint32_t sign = (u >> 16) & 0x8000u;
  u &= 0x7FFFFFFFu;
  if (u >= 0x477FF000u) return (uint16_t)(sign | 0x7BFFu);           // >= 65520 (or inf / nan) -> 65504
  if (u < 0x38800000u) {                                             // subnormal half or zero
    if (u < 0x33000000u) return (uint16_t)sign;
    const int shift = 126 - (int)(u >> 23);                          // 14 .. 24
    const uint32_t mant = (u & 0x7FFFFFu) | 0x800000u;
    const uint32_t half = mant >> shift, rem = mant & ((1u << shift) - 1u), mid = 1u << (shift - 1);
    return (uint16_t)(sign | (half + ((rem > mid || (rem == mid && (half & 1u))) ? 1u : 0u)));
  }
  const uint32_t r = u + 0xFFFu + ((u >> 13) & 1u);                  // round to nearest even at bit 13
  return (uint16_t)(sign | ((r - 0x38000000u) >> 13));
}

// eval-mode BatchNorm as y = x*scale + shift
static void bn_affine(const float* const bn[4], int n, float eps, std::vector<float>& scale,
                      std::vector<float>& shift) {
  scale.assign(n, 1.f);
  shift.assign(n, 0.f);
  if (!bn) return;
  for (int i = 0; i < n; ++i) {
    const float s = bn[0][i] / std::sqrt(bn[3][i] + eps);
    scale[i] = s;
    shift[i] = bn[1][i] - bn[2][i] * s;
  }
}

static void pack_pw(const PwLayout& L, const float* w, const float* bias, const float* const bn[4],
                    float eps, float* blob) {
  std::vector<float> sc, sh;
  bn_affine(bn, L.Cout, eps, sc, sh);
  std::memset(blob, 0, sizeof(float) * L.total);
  for (int n = 0; n < L.Cout; ++n) {
    for (int k = 0; k < L.Cin; ++k) blob[L.o_w + (size_t)k * L.Cout + n] = w[(size_t)n * L.Cin + k] * sc[n];
    blob[L.o_b + n] = (bias ? bias[n] : 0.f) * sc[n] + sh[n];
  }
}

}  // namespace hrf

using namespace hrf;

extern "C" {

int hrf_abi_version(void) { return HRF_ABI_VERSION; }
const char* hrf_last_error(void) { return g_err; }
int hrf_set_pdl(int32_t enable) {
  const int prev = hrf::pdl_enabled() ? 1 : 0;
  hrf::g_pdl.store(enable ? 1 : 0, std::memory_order_relaxed);
  return prev;
}

unsigned long long hrf_launch_count(void) { return g_launches.load(); }

int hrf_device_check(void) {
  int dev = 0;
  HRF_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  HRF_CUDA(cudaGetDeviceProperties(&prop, dev));
  HRF_REQUIRE(prop.major == 10, HRF_EDEVICE, "device %s is sm_%d%d; this library is sm_100a only",
              prop.name, prop.major, prop.minor);
  return HRF_OK;
}

// ------------------------------------------------------------------ attention
static int check_attn(const HrfAttnDesc* d) {
  HRF_REQUIRE(d != nullptr, HRF_EINVAL, "attn: null descriptor");
  HRF_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0 && d->C > 0 && d->heads > 0, HRF_EINVAL,
              "attn: non-positive dimension");
  HRF_REQUIRE(d->C % d->heads == 0, HRF_EINVAL, "attn: C=%d not divisible by heads=%d", d->C, d->heads);
  HRF_REQUIRE(d->C % 2 == 0, HRF_EUNSUPPORTED, "attn: C=%d must be even", d->C);
  HRF_REQUIRE(d->win >= 1 && d->win * d->win <= 256, HRF_EUNSUPPORTED, "attn: window %d", d->win);
  HRF_REQUIRE(d->n_kv >= 0 && d->n_kv <= 8, HRF_EINVAL, "attn: n_kv=%d", d->n_kv);
  HRF_REQUIRE(d->dtype == HRF_F32 || d->dtype == HRF_BF16, HRF_EINVAL, "attn: dtype");
  return HRF_OK;
}

// HRF_ATTN_V3=0 keeps the second-generation kernel for C = 18 (A/B runs)
static bool attn_v3_enabled() {
  const char* e = std::getenv("HRF_ATTN_V3");
  return !(e && e[0] == '0');
}
// which implementation serves a window-attention problem
enum { PATH_TC = 0, PATH_FUSED_SIMT = 1, PATH_GENERIC = 2 };
static int attn_path(const HrfAttnDesc* d) {
  AttnParams p{};
  p.C = d->C; p.heads = d->heads; p.win = d->win;
  if (d->dtype == HRF_BF16 && attn_tc_supported(p) && !tc_disabled()) return PATH_TC;
  // bf16 mode, windows other than 7: un-fused path with every contraction on tcgen05
  // (projection GEMMs + attn_core_tc_kernel) instead of the FFMA fused kernel
  if (d->dtype == HRF_BF16 && d->win != 7 && !tc_disabled() &&
      core_tc_geom(d->win, d->C / d->heads).smem <= 200 * 1024)
    return PATH_GENERIC;
  const AttnLayout L(d->C, d->heads, d->win);
  const size_t smem = L.smem_floats(d->n_kv > 0, kAttnThreads / 32) * sizeof(float);
  return (L.ldx <= 256 && smem <= 227 * 1024) ? PATH_FUSED_SIMT : PATH_GENERIC;
}

size_t hrf_attn_blob_floats(const HrfAttnDesc* d) {
  if (!d || d->heads <= 0 || d->C <= 0) return 0;
  return (size_t)AttnLayout(d->C, d->heads, d->win).total;
}

int hrf_attn_pack(const HrfAttnDesc* d, const float* ln_q_w, const float* ln_q_b,
                  const float* ln_kv_w, const float* ln_kv_b, const float* wq, const float* bq,
                  const float* wk, const float* bk, const float* wv, const float* bv,
                  const float* wo, const float* bo, const float* rpb_table, float* blob) {
  int rc = check_attn(d);
  if (rc) return rc;
  HRF_REQUIRE(ln_q_w && ln_q_b && ln_kv_w && ln_kv_b && wq && wk && wv && wo && blob, HRF_EINVAL,
              "attn_pack: null pointer");
  const AttnLayout L(d->C, d->heads, d->win);
  const int C = L.C;
  std::memset(blob, 0, sizeof(float) * L.total);
  const float scale = 1.0f / std::sqrt((float)L.hd);
  for (int c = 0; c < C; ++c) {
    blob[L.o_lnq_w + c] = ln_q_w[c];
    blob[L.o_lnq_b + c] = ln_q_b[c];
    blob[L.o_lnkv_w + c] = ln_kv_w[c];
    blob[L.o_lnkv_b + c] = ln_kv_b[c];
    blob[L.o_bq + c] = (bq ? bq[c] : 0.f) * scale;
    blob[L.o_bk + c] = bk ? bk[c] : 0.f;
    blob[L.o_bv + c] = bv ? bv[c] : 0.f;
    blob[L.o_bo + c] = bo ? bo[c] : 0.f;
  }
  for (int n = 0; n < C; ++n)
    for (int k = 0; k < C; ++k) {
      blob[L.o_wq + (size_t)k * C + n] = wq[(size_t)n * C + k] * scale;
      blob[L.o_wk + (size_t)k * C + n] = wk[(size_t)n * C + k];
      blob[L.o_wv + (size_t)k * C + n] = wv[(size_t)n * C + k];
      // out_proj consumes the head-padded O layout: input channel k = h*hd+dd
      // lives at row h*hdp+dd
      const int h = k / L.hd, dd = k - h * L.hd;
      blob[L.o_wo + (size_t)(h * L.hdp + dd) * C + n] = wo[(size_t)n * C + k];
    }
  if (rpb_table)
    for (int h = 0; h < L.heads; ++h)
      for (int t = 0; t < L.T; ++t) blob[L.o_rpb + (size_t)h * L.T + t] = rpb_table[(size_t)t * L.heads + h];
  // ---- tensor-core sections (bf16 operand tiles, head-padded) -------------------
  {
    const int HDP = L.tc_HDP, KC = L.tc_KC, NQ = L.tc_NQ, NOUT = L.tc_NOUT;
    float* bias = blob + L.o_tc_bias;
    uint16_t* tq = reinterpret_cast<uint16_t*>(blob + L.o_tc_wq);
    uint16_t* tk = reinterpret_cast<uint16_t*>(blob + L.o_tc_wk);
    uint16_t* tv = reinterpret_cast<uint16_t*>(blob + L.o_tc_wv);
    uint16_t* to = reinterpret_cast<uint16_t*>(blob + L.o_tc_wo);
    for (int h = 0; h < L.heads; ++h)
      for (int dd = 0; dd < L.hd; ++dd) {
        const int n = h * L.hd + dd, np = h * HDP + dd;      // feature, head-padded feature
        bias[np] = (bq ? bq[n] : 0.f) * scale;
        bias[NQ + np] = bk ? bk[n] : 0.f;
        bias[2 * NQ + np] = bv ? bv[n] : 0.f;
        for (int k = 0; k < C; ++k) {
          const size_t e = umma::tile_off(np, k, NQ) / 2;
          tq[e] = f32_to_bf16(wq[(size_t)n * C + k] * scale);
          tk[e] = f32_to_bf16(wk[(size_t)n * C + k]);
          tv[e] = f32_to_bf16(wv[(size_t)n * C + k]);
        }
        if (KC > C) {   // bias row: the kernels feed a constant 1 in column C of LN(x) / LN(z)
          const size_t e = umma::tile_off(np, C, NQ) / 2;
          tq[e] = f32_to_bf16(bias[np]);
          tk[e] = f32_to_bf16(bias[NQ + np]);
          tv[e] = f32_to_bf16(bias[2 * NQ + np]);
        }
        for (int n2 = 0; n2 < C; ++n2)                       // out_proj: rows = out feature, K = padded O
          to[umma::tile_off(n2, np, NOUT) / 2] = f32_to_bf16(wo[(size_t)n2 * C + n]);
      }
    for (int c = 0; c < C; ++c) bias[3 * NQ + c] = bo ? bo[c] : 0.f;
  }
  // ---- third-generation sections (window_attn_v3.cuh): LayerNorm affine, softmax scale, log2 e
  // and the output projection folded into the operand tiles -------------------------------------
  if (L.o_v3 >= 0) {
    using A = AttnV3;
    static_assert(A::SEC_SELF + A::SEC_CROSS == 4784 + 5808, "AttnLayout::o_v3 size");
    constexpr int Cc = A::C;
    const double l2e = 1.4426950408889634, sc = (double)scale * l2e;
    // rows of the three projections over the K index: 0..17 channel (x gamma), 18 bias, 19 beta row
    double rq[Cc][20], rk[Cc][20], rv[Cc + 1][20];
    for (int n = 0; n < Cc; ++n) {
      double bq_b = 0, bk_b = 0;
      for (int k = 0; k < Cc; ++k) {
        rq[n][k] = (double)ln_q_w[k] * wq[(size_t)n * Cc + k] * sc;
        rk[n][k] = (double)ln_kv_w[k] * wk[(size_t)n * Cc + k];
        bq_b += (double)ln_q_b[k] * wq[(size_t)n * Cc + k];
        bk_b += (double)ln_kv_b[k] * wk[(size_t)n * Cc + k];
      }
      rq[n][18] = (bq ? bq[n] : 0.0) * sc; rq[n][19] = bq_b * sc;
      rk[n][18] = bk ? bk[n] : 0.0;        rk[n][19] = bk_b;
      // V' = LN(z) (Wo Wv)^T: one head, so the output projection commutes with the softmax average
      double b18 = bo ? bo[n] : 0.0, b19 = 0;
      for (int k = 0; k < Cc; ++k) rv[n][k] = 0;
      for (int dd = 0; dd < Cc; ++dd) {
        const double wod = wo[(size_t)n * Cc + dd];
        double vb = 0;
        for (int k = 0; k < Cc; ++k) {
          rv[n][k] += wod * wv[(size_t)dd * Cc + k] * ln_kv_w[k];
          vb += (double)ln_kv_b[k] * wv[(size_t)dd * Cc + k];
        }
        b18 += wod * (bv ? bv[dd] : 0.0);
        b19 += wod * vb;
      }
      rv[n][18] = b18; rv[n][19] = b19;
    }
    for (int k = 0; k < 20; ++k) rv[Cc][k] = k == 18 ? 1.0 : 0.0;     // the constant-1 column of V'
    unsigned char* base = reinterpret_cast<unsigned char*>(blob + L.o_v3);
    uint16_t* w_self = reinterpret_cast<uint16_t*>(base);
    uint16_t* w_q = reinterpret_cast<uint16_t*>(base + A::SEC_SELF);
    uint16_t* w_kv = reinterpret_cast<uint16_t*>(base + A::SEC_SELF + A::W_Q_B);
    for (int k = 0; k < 20; ++k) {
      for (int n = 0; n < Cc; ++n) {
        w_self[umma::tile_off(n, k, 64) / 2] = f32_to_bf16((float)rq[n][k]);
        w_self[umma::tile_off(18 + n, k, 64) / 2] = f32_to_bf16((float)rk[n][k]);
        w_q[umma::tile_off(n, k, 32) / 2] = f32_to_bf16((float)rq[n][k]);
        w_kv[umma::tile_off(n, k, 48) / 2] = f32_to_bf16((float)rk[n][k]);
      }
      for (int n = 0; n <= Cc; ++n) {
        w_self[umma::tile_off(36 + n, k, 64) / 2] = f32_to_bf16((float)rv[n][k]);
        w_kv[umma::tile_off(18 + n, k, 48) / 2] = f32_to_bf16((float)rv[n][k]);
      }
    }
    float* t_self = reinterpret_cast<float*>(base + A::W_SELF_B);
    float* t_cross = reinterpret_cast<float*>(base + A::SEC_SELF + A::W_Q_B + A::W_KV_B);
    for (int t = 0; t < 169; ++t) {
      const float v = rpb_table ? (float)(rpb_table[(size_t)t] * l2e) : 0.f;      // heads == 1
      t_self[t] = v;
      t_cross[t] = v;
    }
  }
  return HRF_OK;
}

size_t hrf_attn_workspace_bytes(const HrfAttnDesc* d) {
  if (!d || d->heads <= 0 || d->C <= 0 || d->C % d->heads != 0 || d->win <= 0) return 0;
  switch (attn_path(d)) {
    case PATH_TC: return attn_tc_workspace_bytes(d->B, d->H, d->W, d->C, d->heads);
    case PATH_GENERIC:
      return sizeof(float) * attn_generic_ws_floats(d->B, d->H, d->W, d->C, d->heads, d->win, d->n_kv > 0);
  }
  return 0;
}

int hrf_window_attn_fwd(const HrfAttnDesc* d, const void* x, const void* const* kv,
                        const float* const* blobs, void* out, void* workspace,
                        size_t workspace_bytes, void* stream) {
  int rc = check_attn(d);
  if (rc) return rc;
  HRF_REQUIRE(workspace_bytes >= hrf_attn_workspace_bytes(d) &&
                  (workspace != nullptr || hrf_attn_workspace_bytes(d) == 0),
              HRF_EINVAL, "attn_fwd: workspace of %zu bytes required, %zu given",
              hrf_attn_workspace_bytes(d), workspace_bytes);
  HRF_REQUIRE(x && blobs && out, HRF_EINVAL, "attn_fwd: null pointer");
  HRF_REQUIRE(x != out, HRF_EINVAL, "attn_fwd: out must not alias x");
  HRF_REQUIRE(d->n_kv == 0 || kv, HRF_EINVAL, "attn_fwd: kv list missing");
  cudaStream_t st = (cudaStream_t)stream;
  const int passes = d->n_kv > 0 ? d->n_kv : 1;
  if (attn_path(d) == PATH_TC && AttnV3::applies(d->C, d->heads, d->win) && passes <= AttnV3::MAXMOD && attn_v3_enabled()) {
    // C = 18: one launch for the block, every modality inside it
    AttnV3Params p{};
    p.x = x; p.out = out; p.n_mod = passes;
    p.xs[0] = x; p.outs[0] = out; p.n_prob = 1;
    p.v3_off = AttnLayout(d->C, d->heads, d->win).o_v3;
    for (int k = 0; k < passes; ++k) {
      p.z[k] = d->n_kv > 0 ? kv[k] : x;
      p.blob[k] = blobs[k];
      HRF_REQUIRE(p.z[k] && p.blob[k], HRF_EINVAL, "attn_fwd: null kv/blob %d", k);
      HRF_REQUIRE(p.z[k] != out, HRF_EINVAL, "attn_fwd: out must not alias kv %d", k);
    }
    p.B = d->B; p.H = d->H; p.W = d->W; p.pad_mask = d->with_pad_mask; p.eps = d->ln_eps;
    return launch_window_attn_v3(p, d->n_kv > 0, st);
  }
  for (int k = 0; k < passes; ++k) {
    AttnParams p;
    p.xq = x;
    p.resid = k == 0 ? x : out;   // later modalities accumulate onto out
    p.z = d->n_kv > 0 ? kv[k] : x;
    p.blob = blobs[k];
    p.out = out;
    p.ws = workspace;
    HRF_REQUIRE(p.z && p.blob, HRF_EINVAL, "attn_fwd: null kv/blob %d", k);
    p.B = d->B; p.H = d->H; p.W = d->W; p.C = d->C; p.heads = d->heads; p.win = d->win;
    p.cross = d->n_kv > 0; p.pad_mask = d->with_pad_mask; p.eps = d->ln_eps;
    switch (attn_path(d)) {
      case PATH_TC: rc = launch_window_attn_tc(p, st); break;
      case PATH_FUSED_SIMT:
        rc = d->dtype == HRF_F32 ? launch_window_attn<float>(p, st)
                                 : launch_window_attn<__nv_bfloat16>(p, st);
        break;
      default:
        rc = d->dtype == HRF_F32 ? launch_window_attn_generic<float>(p, st)
                                 : launch_window_attn_generic<__nv_bfloat16>(p, st);
    }
    if (rc) return rc;
  }
  return HRF_OK;
}

// Self-attention (n_kv == 0) of several tensors of ONE shape, each with its own blob, in one launch
// where the kernel supports it (C = 18: window_attn_v3); otherwise one launch per tensor.
int hrf_window_attn_grouped_fwd(const HrfAttnDesc* d, int32_t n, const void* const* xs,
                                const float* const* blobs, void* const* outs, void* workspace,
                                size_t workspace_bytes, void* stream) {
  int rc = check_attn(d);
  if (rc) return rc;
  HRF_REQUIRE(d->n_kv == 0, HRF_EINVAL, "attn_grouped_fwd: self-attention only (n_kv = %d)", d->n_kv);
  HRF_REQUIRE(n >= 1 && xs && blobs && outs, HRF_EINVAL, "attn_grouped_fwd: arguments");
  for (int q = 0; q < n; ++q) {
    HRF_REQUIRE(xs[q] && blobs[q] && outs[q] && xs[q] != outs[q], HRF_EINVAL, "attn_grouped_fwd: pointers of problem %d", q);
  }
  if (n <= AttnV3::MAXMOD && attn_path(d) == PATH_TC && AttnV3::applies(d->C, d->heads, d->win) && attn_v3_enabled()) {
    AttnV3Params p{};
    p.n_mod = 1; p.n_prob = n;
    p.v3_off = AttnLayout(d->C, d->heads, d->win).o_v3;
    for (int q = 0; q < n; ++q) { p.xs[q] = xs[q]; p.outs[q] = outs[q]; p.blob[q] = blobs[q]; }
    p.x = xs[0]; p.out = outs[0];
    p.B = d->B; p.H = d->H; p.W = d->W; p.pad_mask = d->with_pad_mask; p.eps = d->ln_eps;
    return launch_window_attn_v3(p, false, (cudaStream_t)stream);
  }
  for (int q = 0; q < n; ++q) {
    rc = hrf_window_attn_fwd(d, xs[q], nullptr, &blobs[q], outs[q], workspace, workspace_bytes, stream);
    if (rc) return rc;
  }
  return HRF_OK;
}

// ------------------------------------------------------------------ MixFFN
static int check_ffn(const HrfFfnDesc* d) {
  HRF_REQUIRE(d != nullptr, HRF_EINVAL, "ffn: null descriptor");
  HRF_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0 && d->C > 0 && d->hidden > 0, HRF_EINVAL,
              "ffn: non-positive dimension");
  HRF_REQUIRE(d->C % 2 == 0 && d->hidden % 4 == 0, HRF_EUNSUPPORTED,
              "ffn: C=%d must be even and hidden=%d a multiple of 4", d->C, d->hidden);
  HRF_REQUIRE(d->dtype == HRF_F32 || d->dtype == HRF_BF16, HRF_EINVAL, "ffn: dtype");
  return HRF_OK;
}

static int ffn_path(const HrfFfnDesc* d) {
  FfnParams p{};
  p.C = d->C; p.hidden = d->hidden;
  if (d->dtype == HRF_BF16 && ffn_tc_supported(p) && !tc_disabled()) return PATH_TC;
  const FfnLayout L(d->C, d->hidden);
  const int CT = d->C % 4 == 0 ? 4 : 2;
  const int maxt = ceil_div((kFfnOut / 4) * (d->C / CT), kFfnThreads);
  const bool fits = L.ldx <= 256 && maxt <= (CT == 4 ? 4 : 3) &&
                    ffn_smem_floats(L) * sizeof(float) <= 227 * 1024;
  return fits ? PATH_FUSED_SIMT : PATH_GENERIC;
}

size_t hrf_ffn_blob_floats(const HrfFfnDesc* d) {
  if (!d || d->C <= 0 || d->hidden <= 0) return 0;
  return (size_t)FfnLayout(d->C, d->hidden).total;
}

int hrf_ffn_pack(const HrfFfnDesc* d, const float* ln_w, const float* ln_b, const float* w1,
                 const float* b1, const float* const bn1[4], const float* wd, const float* bd,
                 const float* const bn2[4], const float* w2, const float* b2,
                 const float* const bn3[4], float bn_eps, float* blob) {
  int rc = check_ffn(d);
  if (rc) return rc;
  HRF_REQUIRE(ln_w && ln_b && w1 && wd && w2 && blob, HRF_EINVAL, "ffn_pack: null pointer");
  const FfnLayout L(d->C, d->hidden);
  const int C = L.C, Hd = L.hidden;
  std::memset(blob, 0, sizeof(float) * L.total);
  std::vector<float> s1, t1, s2, t2, s3, t3;
  bn_affine(bn1, Hd, bn_eps, s1, t1);
  bn_affine(bn2, Hd, bn_eps, s2, t2);
  bn_affine(bn3, C, bn_eps, s3, t3);
  for (int c = 0; c < C; ++c) {
    blob[L.o_ln_w + c] = ln_w[c];
    blob[L.o_ln_b + c] = ln_b[c];
    blob[L.o_b2 + c] = (b2 ? b2[c] : 0.f) * s3[c] + t3[c];
  }
  for (int j = 0; j < Hd; ++j) {
    for (int k = 0; k < C; ++k) blob[L.o_w1 + (size_t)k * Hd + j] = w1[(size_t)j * C + k] * s1[j];
    blob[L.o_b1 + j] = (b1 ? b1[j] : 0.f) * s1[j] + t1[j];
    for (int t = 0; t < 9; ++t) blob[L.o_wd + (size_t)t * Hd + j] = wd[(size_t)j * 9 + t] * s2[j];
    blob[L.o_bd + j] = (bd ? bd[j] : 0.f) * s2[j] + t2[j];
    for (int n = 0; n < C; ++n) blob[L.o_w2 + (size_t)j * C + n] = w2[(size_t)n * Hd + j] * s3[n];
  }
  // ---- tensor-core sections ------------------------------------------------------
  if (L.tc_nchunk > 0) {
    const int KC = L.tc_KC, NOUT = L.tc_NOUT;
    float* f = blob + L.o_tc_f32;
    uint16_t* tile1 = reinterpret_cast<uint16_t*>(blob + L.o_tc_w1);
    uint16_t* tile2 = reinterpret_cast<uint16_t*>(blob + L.o_tc_w2);
    for (int ch = 0; ch < L.tc_nchunk; ++ch) {
      float* fb = f + (size_t)ch * 880;                       // b1[80] | wd[9][80] | bd[80]
      uint16_t* w1t = tile1 + (size_t)ch * 80 * KC;           // B of fc1: rows = hidden ch, K = C
      uint16_t* w2t = tile2 + (size_t)ch * NOUT * 80;         // B of fc2: rows = out ch, K = hidden ch
      for (int jj = 0; jj < L.tc_CH; ++jj) {
        const int j = ch * L.tc_CH + jj;
        fb[jj] = (b1 ? b1[j] : 0.f) * s1[j] + t1[j];
        for (int t = 0; t < 9; ++t) fb[80 + t * 80 + jj] = wd[(size_t)j * 9 + t] * s2[j];
        fb[800 + jj] = (bd ? bd[j] : 0.f) * s2[j] + t2[j];
        for (int k = 0; k < C; ++k)
          w1t[umma::tile_off(jj, k, 80) / 2] = f32_to_bf16(w1[(size_t)j * C + k] * s1[j]);
        // bias row: the kernels that have a spare K column (KC > C) feed a constant 1 in
        // column C of the activation tile, so b1 rides through the MMA
        if (KC > C) w1t[umma::tile_off(jj, C, 80) / 2] = f32_to_bf16(fb[jj]);
        for (int n = 0; n < C; ++n)
          w2t[umma::tile_off(n, jj, NOUT) / 2] = f32_to_bf16(w2[(size_t)n * Hd + j] * s3[n]);
      }
    }
    float* b2p = f + (size_t)L.tc_nchunk * 880;
    for (int c = 0; c < C; ++c) {
      b2p[c] = (b2 ? b2[c] : 0.f) * s3[c] + t3[c];
      // fc2 bias row: K index CH (first padding channel of the chunk) of the FIRST chunk's W2
      // tile; the kernels keep a constant 1 in that column of the hidden activations
      tile2[umma::tile_off(c, L.tc_CH, NOUT) / 2] = f32_to_bf16(b2p[c]);
    }
    // second-generation kernel (mixffn_v2.cuh): LN affine folded into W1 / b1, every section
    // pre-multiplied by 0.5, hidden path in fp16
    if (L.tc_CH == 72) {
      uint16_t* v1 = reinterpret_cast<uint16_t*>(blob + L.o_v2_w1);
      uint16_t* v2 = reinterpret_cast<uint16_t*>(blob + L.o_v2_w2);
      uint16_t* cv = reinterpret_cast<uint16_t*>(blob + L.o_v2_cv);
      for (int ch = 0; ch < L.tc_nchunk; ++ch) {
        uint16_t* w1t = v1 + (size_t)ch * 80 * KC;
        uint16_t* w2t = v2 + (size_t)ch * NOUT * 80;
        uint16_t* cvt = cv + (size_t)ch * 800;                // wd[9][80] | bd[80]
        for (int jj = 0; jj < 72; ++jj) {
          const int j = ch * 72 + jj;
          double bfold = (double)((b1 ? b1[j] : 0.f) * s1[j] + t1[j]);
          for (int k = 0; k < C; ++k) {
            const float w = w1[(size_t)j * C + k] * s1[j];
            w1t[umma::tile_off(jj, k, 80) / 2] = f32_to_bf16(0.5f * w * ln_w[k]);
            bfold += (double)w * (double)ln_b[k];
          }
          w1t[umma::tile_off(jj, C, 80) / 2] = f32_to_bf16(0.5f * (float)bfold);
          for (int t = 0; t < 9; ++t) cvt[t * 80 + jj] = f32_to_f16(0.5f * wd[(size_t)j * 9 + t] * s2[j]);
          cvt[720 + jj] = f32_to_f16(0.5f * ((bd ? bd[j] : 0.f) * s2[j] + t2[j]));
          for (int n = 0; n < C; ++n)
            w2t[umma::tile_off(n, jj, NOUT) / 2] = f32_to_f16(0.5f * w2[(size_t)n * Hd + j] * s3[n]);
        }
      }
      for (int c = 0; c < C; ++c) v2[umma::tile_off(c, 72, NOUT) / 2] = f32_to_f16(0.5f * b2p[c]);
    }
  }
  return HRF_OK;
}

size_t hrf_ffn_workspace_bytes(const HrfFfnDesc* d) {
  if (!d || d->C <= 0 || d->hidden <= 0) return 0;
  switch (ffn_path(d)) {
    case PATH_TC: return ffn_tc_workspace_bytes(d->B, d->H, d->W, d->C, d->hidden);
    case PATH_GENERIC: return sizeof(float) * ffn_generic_ws_floats(d->B, d->H, d->W, d->C, d->hidden);
  }
  return 0;
}

int hrf_mixffn_fwd(const HrfFfnDesc* d, const void* x, const float* blob, void* out,
                   void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_ffn(d);
  if (rc) return rc;
  HRF_REQUIRE(x && blob && out, HRF_EINVAL, "ffn_fwd: null pointer");
  HRF_REQUIRE(x != out, HRF_EINVAL, "ffn_fwd: out must not alias x (3x3 halo reads)");
  HRF_REQUIRE(workspace_bytes >= hrf_ffn_workspace_bytes(d) &&
                  (workspace != nullptr || hrf_ffn_workspace_bytes(d) == 0),
              HRF_EINVAL, "ffn_fwd: workspace of %zu bytes required, %zu given",
              hrf_ffn_workspace_bytes(d), workspace_bytes);
  FfnParams p{x, blob, out, workspace, d->B, d->H, d->W, d->C, d->hidden, d->ln_eps, FastDiv(), FastDiv()};
  cudaStream_t st = (cudaStream_t)stream;
  switch (ffn_path(d)) {
    case PATH_TC: return ffn_v2_supported(p) ? launch_mixffn_v2(p, st) : launch_mixffn_tc(p, st);
    case PATH_FUSED_SIMT:
      return d->dtype == HRF_F32 ? launch_mixffn<float>(p, st) : launch_mixffn<__nv_bfloat16>(p, st);
  }
  return d->dtype == HRF_F32 ? launch_mixffn_generic<float>(p, st)
                             : launch_mixffn_generic<__nv_bfloat16>(p, st);
}

// MixFFN of several tensors of ONE shape, each with its own blob, in one launch where the kernel
// supports it (C = 18: mixffn_v2); otherwise one launch per tensor.
int hrf_mixffn_grouped_fwd(const HrfFfnDesc* d, int32_t n, const void* const* xs, const float* const* blobs,
                           void* const* outs, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_ffn(d);
  if (rc) return rc;
  HRF_REQUIRE(n >= 1 && xs && blobs && outs, HRF_EINVAL, "ffn_grouped_fwd: arguments");
  bool v2 = n <= kFfnMaxProb && ffn_path(d) == PATH_TC;
  for (int q = 0; q < n; ++q) {
    HRF_REQUIRE(xs[q] && blobs[q] && outs[q] && xs[q] != outs[q], HRF_EINVAL, "ffn_grouped_fwd: pointers of problem %d", q);
    FfnParams pq{xs[q], blobs[q], outs[q], nullptr, d->B, d->H, d->W, d->C, d->hidden, d->ln_eps, FastDiv(), FastDiv()};
    v2 = v2 && ffn_v2_supported(pq);
  }
  if (v2) {
    FfnParams p{xs[0], blobs[0], outs[0], nullptr, d->B, d->H, d->W, d->C, d->hidden, d->ln_eps, FastDiv(), FastDiv()};
    FfnV2Problems pb;
    pb.n = n;
    for (int q = 0; q < n; ++q) { pb.x[q] = xs[q]; pb.blob[q] = blobs[q]; pb.out[q] = outs[q]; }
    return launch_mixffn_v2(p, (cudaStream_t)stream, &pb);
  }
  for (int q = 0; q < n; ++q) {
    rc = hrf_mixffn_fwd(d, xs[q], blobs[q], outs[q], workspace, workspace_bytes, stream);
    if (rc) return rc;
  }
  return HRF_OK;
}

// ------------------------------------------------------------------ exchange
size_t hrf_pw_blob_floats(const HrfPwDesc* d) {
  if (!d || d->Cin <= 0 || d->Cout <= 0) return 0;
  return (size_t)PwLayout(d->Cin, d->Cout).total;
}
int hrf_pw_pack(const HrfPwDesc* d, const float* w, const float* bias, const float* const bn[4],
                float bn_eps, float* blob) {
  HRF_REQUIRE(d && w && blob, HRF_EINVAL, "pw_pack: null pointer");
  pack_pw(PwLayout(d->Cin, d->Cout), w, bias, bn, bn_eps, blob);
  return HRF_OK;
}
int hrf_pw_fwd(const HrfPwDesc* d, const void* x, const float* blob, void* out, void* stream) {
  HRF_REQUIRE(d && x && blob && out, HRF_EINVAL, "pw_fwd: null pointer");
  HRF_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, HRF_EINVAL, "pw_fwd: dims");
  PwParams p{x, blob, out, d->B * d->H * d->W, d->Cin, d->Cout, d->relu, 0};
  if (d->dtype == HRF_F32) return launch_pw<float>(p, (cudaStream_t)stream);
  if (d->dtype == HRF_BF16) return launch_pw<__nv_bfloat16>(p, (cudaStream_t)stream);
  HRF_REQUIRE(false, HRF_EINVAL, "pw_fwd: dtype");
}

size_t hrf_conv3x3_blob_floats(const HrfConvDesc* d) {
  if (!d || d->Cin <= 0 || d->Cout <= 0) return 0;
  return (size_t)ConvTcLayout(d->Cin, d->Cout).total;
}
int hrf_conv3x3_pack(const HrfConvDesc* d, const float* w, const float* bias, const float* const bn[4],
                     float bn_eps, float* blob) {
  HRF_REQUIRE(d && w && blob, HRF_EINVAL, "conv3x3_pack: null pointer");
  // (Cout, Cin, 3, 3) -> (Cout, tap, Cin): the K index of the implicit GEMM is tap*Cin + c
  const int Cin = d->Cin, Cout = d->Cout;
  std::vector<float> wp((size_t)Cout * 9 * Cin);
  for (int n = 0; n < Cout; ++n)
    for (int c = 0; c < Cin; ++c)
      for (int t = 0; t < 9; ++t) wp[((size_t)n * 9 + t) * Cin + c] = w[((size_t)n * Cin + c) * 9 + t];
  const PwLayout P(9 * Cin, Cout);
  const ConvTcLayout T(Cin, Cout);
  std::memset(blob, 0, sizeof(float) * T.total);
  pack_pw(P, wp.data(), bias, bn, bn_eps, blob);
  if (T.total > P.total) {
    // tensor-core section: per N split and tap a bf16 B tile [KC/8][NOUT][8]; row k = Cin of the
    // centre tap holds the folded bias (it multiplies the constant-1 column of the activations)
    uint16_t* wt = reinterpret_cast<uint16_t*>(blob + T.o_w);
    const int cper = Cout / T.NSPLIT;
    for (int sp = 0; sp < T.NSPLIT; ++sp)
      for (int t = 0; t < 9; ++t)
        for (int nn = 0; nn < cper; ++nn) {
          const int n = sp * cper + nn;
          uint16_t* tile = wt + ((size_t)sp * 9 + t) * T.NOUT * T.KC;
          for (int c = 0; c < Cin; ++c)    // folded weight from the fp32 section: Wt[k][n]
            tile[umma::tile_off(nn, c, T.NOUT) / 2] = f32_to_bf16(blob[P.o_w + (size_t)(t * Cin + c) * Cout + n]);
          if (t == 4) tile[umma::tile_off(nn, Cin, T.NOUT) / 2] = f32_to_bf16(blob[P.o_b + n]);
        }
  }
  return HRF_OK;
}
int hrf_conv3x3_fwd(const HrfConvDesc* d, const void* x, const float* blob, void* out, void* stream) {
  HRF_REQUIRE(d && x && blob && out, HRF_EINVAL, "conv3x3_fwd: null pointer");
  HRF_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, HRF_EINVAL, "conv3x3_fwd: dims");
  Conv3Params p{};
  p.x = x; p.blob = blob; p.out = out;
  p.B = d->B; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.Cout = d->Cout;
  p.stride = d->stride; p.relu = d->relu;
  const int rc = conv3x3_prepare(p);
  if (rc) return rc;
  if (d->dtype == HRF_F32) return launch_conv3x3<float>(p, (cudaStream_t)stream);
  if (d->dtype == HRF_BF16)
    return conv3x3_tc_supported(p) ? launch_conv3x3_tc(p, (cudaStream_t)stream)
                                   : launch_conv3x3<__nv_bfloat16>(p, (cudaStream_t)stream);
  HRF_REQUIRE(false, HRF_EINVAL, "conv3x3_fwd: dtype");
}

// ------------------------------------------------------------------ dense conv as TMA + tcgen05 GEMM
int hrf_convgemm_supported(const HrfConvGemmDesc* d) {
  HRF_REQUIRE(d != nullptr, HRF_EINVAL, "convgemm: null descriptor");
  HRF_REQUIRE(d->B > 0 && conv_gemm_supported(d->Cin, d->Cout, d->ksize, d->stride, d->H, d->W), HRF_EUNSUPPORTED,
              "convgemm: %dx%d conv %d -> %d stride %d is outside what the kernel covers", d->ksize, d->ksize,
              d->Cin, d->Cout, d->stride);
  return HRF_OK;
}
size_t hrf_convgemm_blob_floats(const HrfConvGemmDesc* d) {
  if (!d || hrf_convgemm_supported(d) != HRF_OK) return 0;
  return (size_t)ConvGemmLayout(d->Cin, d->Cout, d->ksize * d->ksize).total;
}
int hrf_convgemm_pack(const HrfConvGemmDesc* d, const float* w, const float* bias, const float* const bn[4],
                      float bn_eps, const float* extra_bias, float* blob) {
  int rc = hrf_convgemm_supported(d);
  if (rc) return rc;
  HRF_REQUIRE(w && blob, HRF_EINVAL, "convgemm_pack: null pointer");
  const int taps = d->ksize * d->ksize, Cin = d->Cin, Cout = d->Cout;
  const ConvGemmLayout L(Cin, Cout, taps);
  std::vector<float> sc, sh;
  bn_affine(bn, Cout, bn_eps, sc, sh);
  std::memset(blob, 0, sizeof(float) * L.total);
  uint16_t* wt = reinterpret_cast<uint16_t*>(blob + L.o_w);
  for (int n = 0; n < Cout; ++n) {
    blob[L.o_bias + n] = (bias ? bias[n] : 0.f) * sc[n] + sh[n] + (extra_bias ? extra_bias[n] : 0.f);
    for (int c = 0; c < Cin; ++c)
      for (int t = 0; t < taps; ++t)
        wt[((size_t)t * L.NPAD + n) * Cin + c] = f32_to_bf16(w[((size_t)n * Cin + c) * taps + t] * sc[n]);
  }
  return HRF_OK;
}
static int convgemm_launch(const HrfConvGemmDesc* d, int32_t n, const void* const* xs, const void* const* xs2,
                           int32_t cin1, const void* const* resids, const float* const* blobs, void* const* outs,
                           void* stream) {
  int rc = hrf_convgemm_supported(d);
  if (rc) return rc;
  if (xs2) {
    HRF_REQUIRE(d->ksize == 1 && d->stride == 1 && !resids, HRF_EUNSUPPORTED,
                "convgemm_cat_fwd: 1x1, stride 1, no residual");
    HRF_REQUIRE(cin1 > 0 && cin1 < d->Cin && cin1 % 64 == 0, HRF_EUNSUPPORTED,
                "convgemm_cat_fwd: cin1=%d of Cin=%d (both parts multiples of 64)", cin1, d->Cin);
    for (int q = 0; q < n; ++q)
      HRF_REQUIRE(xs2[q] && xs2[q] != outs[q] && (reinterpret_cast<uintptr_t>(xs2[q]) & 15) == 0, HRF_EINVAL,
                  "convgemm_cat_fwd: x2 null, aliasing out or not 16-byte aligned (problem %d)", q);
  }
  HRF_REQUIRE(n >= 1 && n <= kMaxProb, HRF_EINVAL, "convgemm_fwd: 1..%d problems per launch, %d given", kMaxProb, n);
  HRF_REQUIRE(xs && blobs && outs, HRF_EINVAL, "convgemm_fwd: null pointer");
  ConvGemmParams p{};
  for (int q = 0; q < n; ++q) {
    const void* r = resids ? resids[q] : nullptr;
    HRF_REQUIRE(xs[q] && blobs[q] && outs[q], HRF_EINVAL, "convgemm_fwd: null pointer (problem %d)", q);
    HRF_REQUIRE(xs[q] != outs[q], HRF_EINVAL, "convgemm_fwd: out must not alias x");
    HRF_REQUIRE((r != nullptr) == (resids != nullptr && resids[0] != nullptr), HRF_EINVAL,
                "convgemm_fwd: residuals for all problems or for none");
    HRF_REQUIRE(((reinterpret_cast<uintptr_t>(xs[q]) | reinterpret_cast<uintptr_t>(outs[q]) |
                  reinterpret_cast<uintptr_t>(blobs[q]) | reinterpret_cast<uintptr_t>(r)) & 15) == 0,
                HRF_EINVAL, "convgemm_fwd: pointers must be 16-byte aligned");
    p.blob[q] = blobs[q]; p.resid[q] = r; p.out[q] = outs[q];
  }
  p.n_prob = n;
  p.B = d->B; p.Cin = d->Cin; p.Cout = d->Cout; p.taps = d->ksize * d->ksize; p.relu = d->relu; p.stride = d->stride;
  p.Ho = (d->H + d->stride - 1) / d->stride;
  p.Wo = (d->W + d->stride - 1) / d->stride;
  p.tiles_w = ceil_div(p.Wo, cg::TM_W);
  p.tiles_h = ceil_div(p.Ho, cg::TM_H);
  p.n_tiles = p.B * p.tiles_w * p.tiles_h;
  p.d_tiles_prob = FastDiv(p.n_tiles);
  p.d_tiles_img = FastDiv(p.tiles_w * p.tiles_h);
  p.d_tiles_w = FastDiv(p.tiles_w);
  cudaStream_t st = (cudaStream_t)stream;
  switch (ConvGemmLayout(d->Cin, d->Cout, p.taps).NPAD) {
    case 32: return launch_conv_gemm_n<32>(p, xs, xs2, cin1, d->H, d->W, d->stride, st);
    case 64: return launch_conv_gemm_n<64>(p, xs, xs2, cin1, d->H, d->W, d->stride, st);
    case 128: return launch_conv_gemm_n<128>(p, xs, xs2, cin1, d->H, d->W, d->stride, st);
    default: return launch_conv_gemm_n<256>(p, xs, xs2, cin1, d->H, d->W, d->stride, st);
  }
}
int hrf_convgemm_grouped_fwd(const HrfConvGemmDesc* d, int32_t n, const void* const* xs, const void* const* resids,
                             const float* const* blobs, void* const* outs, void* stream) {
  return convgemm_launch(d, n, xs, nullptr, 0, resids, blobs, outs, stream);
}
int hrf_convgemm_grouped_cat_fwd(const HrfConvGemmDesc* d, int32_t n, const void* const* xs, const void* const* xs2,
                                 int32_t cin1, const float* const* blobs, void* const* outs, void* stream) {
  HRF_REQUIRE(xs2, HRF_EINVAL, "convgemm_cat_fwd: null pointer");
  return convgemm_launch(d, n, xs, xs2, cin1, nullptr, blobs, outs, stream);
}
int hrf_convgemm_fwd(const HrfConvGemmDesc* d, const void* x, const void* resid, const float* blob, void* out,
                     void* stream) {
  return hrf_convgemm_grouped_fwd(d, 1, &x, resid ? &resid : nullptr, &blob, &out, stream);
}

size_t hrf_dwpw_blob_floats(const HrfDwPwDesc* d) {
  if (!d || d->Cin <= 0 || d->Cout <= 0) return 0;
  return (size_t)DwPwLayout(d->Cin, d->Cout).total;
}
int hrf_dwpw_pack(const HrfDwPwDesc* d, const float* wdw, const float* const bn_dw[4],
                  const float* wpw, const float* const bn_pw[4], float bn_eps, float* blob) {
  HRF_REQUIRE(d && wdw && wpw && blob, HRF_EINVAL, "dwpw_pack: null pointer");
  const DwPwLayout D(d->Cin, d->Cout);
  std::memset(blob, 0, sizeof(float) * D.total);
  std::vector<float> s, t;
  bn_affine(bn_dw, d->Cin, bn_eps, s, t);
  for (int c = 0; c < d->Cin; ++c) {
    for (int k = 0; k < 9; ++k) blob[D.o_wd + (size_t)k * d->Cin + c] = wdw[(size_t)c * 9 + k] * s[c];
    blob[D.o_bd + c] = t[c];
  }
  pack_pw(PwLayout(d->Cin, d->Cout), wpw, nullptr, bn_pw, bn_eps, blob + D.o_pw);
  return HRF_OK;
}
int hrf_dwpw_fwd(const HrfDwPwDesc* d, const void* x, const float* blob, void* out, void* stream) {
  HRF_REQUIRE(d && x && blob && out, HRF_EINVAL, "dwpw_fwd: null pointer");
  HRF_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, HRF_EINVAL, "dwpw_fwd: dims");
  DwPwParams p{x, blob, out, d->B, d->H, d->W, (d->H + 1) / 2, (d->W + 1) / 2, d->Cin, d->Cout, d->relu, 0};
  if (d->dtype == HRF_F32) return launch_dwpw<float>(p, (cudaStream_t)stream);
  if (d->dtype == HRF_BF16) return launch_dwpw<__nv_bfloat16>(p, (cudaStream_t)stream);
  HRF_REQUIRE(false, HRF_EINVAL, "dwpw_fwd: dtype");
}

int hrf_fuse_sum_fwd(const HrfFuseDesc* d, const void* x, const void* const* up,
                     const void* const* same, void* out, float* out_nchw_f32, void* stream) {
  HRF_REQUIRE(d && x && out, HRF_EINVAL, "fuse_fwd: null pointer");
  HRF_REQUIRE(d->n_up >= 0 && d->n_up <= HRF_MAX_FUSE_TERMS && d->n_same >= 0 &&
                  d->n_same <= HRF_MAX_FUSE_TERMS, HRF_EINVAL, "fuse_fwd: term count");
  HRF_REQUIRE((d->n_up == 0 || up) && (d->n_same == 0 || same), HRF_EINVAL, "fuse_fwd: term list");
  FuseParams p{};
  p.x = x; p.out = out; p.out_nchw = out_nchw_f32;
  p.B = d->B; p.H = d->H; p.W = d->W; p.C = d->C; p.n_up = d->n_up; p.n_same = d->n_same;
  p.relu = d->relu;
  for (int j = 0; j < d->n_up; ++j) {
    HRF_REQUIRE(up[j] && d->up_H[j] > 0 && d->up_W[j] > 0, HRF_EINVAL, "fuse_fwd: up term %d", j);
    p.up[j] = up[j]; p.up_H[j] = d->up_H[j]; p.up_W[j] = d->up_W[j];
  }
  for (int j = 0; j < d->n_same; ++j) {
    HRF_REQUIRE(same[j], HRF_EINVAL, "fuse_fwd: same term %d", j);
    p.same[j] = same[j];
  }
  if (d->dtype == HRF_F32) return launch_fuse<float>(p, (cudaStream_t)stream);
  if (d->dtype == HRF_BF16) return launch_fuse<__nv_bfloat16>(p, (cudaStream_t)stream);
  HRF_REQUIRE(false, HRF_EINVAL, "fuse_fwd: dtype");
}

size_t hrf_stem_blob_floats(const HrfStemDesc* d) {
  if (!d || d->Cin <= 0 || d->Cout <= 0) return 0;
  return (size_t)StemLayout(d->Cin, d->Cout).total;
}

int hrf_stem_pack(const HrfStemDesc* d, const float* w, const float* const bn[4], float bn_eps,
                  float* blob) {
  HRF_REQUIRE(d && w && blob, HRF_EINVAL, "stem_pack: null pointer");
  HRF_REQUIRE(d->Cin >= 1 && d->Cin <= 3 && d->Cout % 16 == 0, HRF_EUNSUPPORTED,
              "stem_pack: Cin=%d Cout=%d", d->Cin, d->Cout);
  const StemLayout L(d->Cin, d->Cout);
  std::memset(blob, 0, sizeof(float) * L.total);
  std::vector<float> sc, sh;
  bn_affine(bn, d->Cout, bn_eps, sc, sh);
  uint16_t* wt = reinterpret_cast<uint16_t*>(blob + L.o_w);
  for (int n = 0; n < d->Cout; ++n) {
    blob[L.o_bias + n] = sh[n];
    for (int k = 0; k < d->Cin * 9; ++k)       // (Cout, Cin, 3, 3) row-major: k = (ci*3+ky)*3+kx
      wt[umma::tile_off(n, k, d->Cout) / 2] = f32_to_bf16(w[(size_t)n * d->Cin * 9 + k] * sc[n]);
  }
  return HRF_OK;
}

int hrf_stem_conv_fwd(const HrfStemDesc* d, const float* x_nchw, const float* blob,
                      void* out_nhwc_bf16, void* stream) {
  HRF_REQUIRE(d && x_nchw && blob && out_nhwc_bf16, HRF_EINVAL, "stem_conv: null pointer");
  HRF_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0, HRF_EINVAL, "stem_conv: dims");
  StemParams p{x_nchw, blob, static_cast<__nv_bfloat16*>(out_nhwc_bf16), d->B, d->Cin, d->H, d->W,
               (d->H + 1) / 2, (d->W + 1) / 2, d->Cout, d->relu};
  return launch_stem_conv(p, (cudaStream_t)stream);
}

int hrf_bias_act_fwd(int64_t n_tokens, int32_t C, int32_t dtype, int32_t relu, void* y,
                     const float* bias, const void* residual, void* stream) {
  HRF_REQUIRE(y && bias && n_tokens > 0 && C > 0, HRF_EINVAL, "bias_act: args");
  if (dtype == HRF_F32)
    return launch_bias_act<float>(y, bias, residual, (size_t)n_tokens, C, relu, (cudaStream_t)stream);
  if (dtype == HRF_BF16)
    return launch_bias_act<__nv_bfloat16>(y, bias, residual, (size_t)n_tokens, C, relu,
                                          (cudaStream_t)stream);
  HRF_REQUIRE(false, HRF_EINVAL, "bias_act: dtype");
}

static int bn_check(const HrfBnDesc* d) {
  HRF_REQUIRE(d != nullptr, HRF_EINVAL, "bn: null descriptor");
  HRF_REQUIRE(d->B > 0 && d->C > 0 && d->HW > 0, HRF_EINVAL, "bn: B=%d C=%d HW=%d", d->B, d->C, d->HW);
  HRF_REQUIRE(d->dtype == HRF_F32 || d->dtype == HRF_BF16, HRF_EINVAL, "bn: dtype");
  return HRF_OK;
}
size_t hrf_bn_workspace_bytes(const HrfBnDesc* d) {
  if (bn_check(d) != HRF_OK) return 0;
  return bn_workspace_bytes(d->B, d->C, d->HW);
}
int hrf_bn_stats(const HrfBnDesc* d, const void* x, double* sums, void* ws, size_t ws_bytes,
                 void* stream) {
  if (int rc = bn_check(d)) return rc;
  HRF_REQUIRE(x && sums && ws, HRF_EINVAL, "bn_stats: null pointer");
  HRF_REQUIRE(ws_bytes >= bn_workspace_bytes(d->B, d->C, d->HW), HRF_EINVAL, "bn_stats: workspace too small");
  if (d->dtype == HRF_F32)
    return launch_bn_reduce<float, false>(d->B, d->C, d->HW, x, nullptr, nullptr, nullptr, nullptr, nullptr, 0, sums, nullptr, nullptr, ws, (cudaStream_t)stream);
  return launch_bn_reduce<__nv_bfloat16, false>(d->B, d->C, d->HW, x, nullptr, nullptr, nullptr, nullptr, nullptr, 0, sums, nullptr, nullptr, ws, (cudaStream_t)stream);
}
static int bn_act_check(int32_t act) {
  HRF_REQUIRE(act >= BN_ACT_NONE && act <= BN_ACT_GELU, HRF_EINVAL, "bn: act %d (0 none, 1 ReLU, 2 GELU)", act);
  return HRF_OK;
}
int hrf_bn_bwd_stats(const HrfBnDesc* d, const void* x, const void* dy, const float* mean,
                     const float* invstd, const float* weight, const float* bias, int32_t act,
                     double* sums, float* dweight, float* dbias, void* ws, size_t ws_bytes,
                     void* stream) {
  if (int rc = bn_check(d)) return rc;
  if (int rc = bn_act_check(act)) return rc;
  HRF_REQUIRE(x && dy && mean && invstd && sums && ws, HRF_EINVAL, "bn_bwd_stats: null pointer");
  HRF_REQUIRE(ws_bytes >= bn_workspace_bytes(d->B, d->C, d->HW), HRF_EINVAL, "bn_bwd_stats: workspace too small");
  if (d->dtype == HRF_F32)
    return launch_bn_reduce<float, true>(d->B, d->C, d->HW, x, dy, mean, invstd, weight, bias, act, sums, dweight, dbias, ws, (cudaStream_t)stream);
  return launch_bn_reduce<__nv_bfloat16, true>(d->B, d->C, d->HW, x, dy, mean, invstd, weight, bias, act, sums, dweight, dbias, ws, (cudaStream_t)stream);
}
// ------------------------------------------------------------------ training-mode attention core
static int attn_core_check(int nWin, int N, int C, int heads) {
  HRF_REQUIRE(nWin > 0 && N > 0 && C > 0 && heads > 0 && C % heads == 0, HRF_EINVAL, "attn_core_train: dimensions");
  HRF_REQUIRE(N <= kAtMaxN && C / heads <= kAtMaxHd, HRF_EUNSUPPORTED,
              "attn_core_train: N=%d (<= %d) head_dim=%d (<= %d)", N, kAtMaxN, C / heads, kAtMaxHd);
  return HRF_OK;
}
int hrf_attn_core_train_fwd(int32_t nWin, int32_t N, int32_t C, int32_t heads, float scale, const float* q,
                            const float* k, const float* v, const float* table, const int32_t* rpi, float* o,
                            float* P, void* stream) {
  if (int rc = attn_core_check(nWin, N, C, heads)) return rc;
  HRF_REQUIRE(q && k && v && o && P && ((table == nullptr) == (rpi == nullptr)), HRF_EINVAL, "attn_core_train_fwd: pointers");
  AttnCoreTrain p{q, k, v, table, rpi, o, P, nWin, N, C, heads, C / heads, scale};
  return launch_attn_core_train_fwd(p, (cudaStream_t)stream);
}
size_t hrf_attn_core_train_ws_floats(int32_t nWin, int32_t N, int32_t heads) {
  if (nWin <= 0 || N <= 0 || heads <= 0) return 0;
  return ((size_t)attn_core_chunks(nWin, heads) + 1) * heads * N * N;      // per-CTA partials + their sum
}
int hrf_attn_core_train_bwd(int32_t nWin, int32_t N, int32_t C, int32_t heads, float scale, const float* q,
                            const float* k, const float* v, const float* P, const float* dout, float* dq,
                            float* dk, float* dv, const int32_t* rpi, int32_t T, float* dtable, float* workspace,
                            size_t workspace_floats, void* stream) {
  if (int rc = attn_core_check(nWin, N, C, heads)) return rc;
  HRF_REQUIRE(q && k && v && P && dout && dq && dk && dv, HRF_EINVAL, "attn_core_train_bwd: null pointer");
  if (dtable) {
    HRF_REQUIRE(rpi && workspace && T > 0, HRF_EINVAL, "attn_core_train_bwd: dtable needs rpi, T and a workspace");
    HRF_REQUIRE(workspace_floats >= hrf_attn_core_train_ws_floats(nWin, N, heads), HRF_EINVAL,
                "attn_core_train_bwd: workspace too small");
  }
  AttnCoreTrainBwd p{q, k, v, P, dout, dq, dk, dv, dtable ? workspace : nullptr, nWin, N, C, heads, C / heads, 0, scale};
  return launch_attn_core_train_bwd(p, rpi, T, dtable, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ training-mode depthwise conv
static int dw_train_check(int B, int C, int H, int W, int stride) {
  HRF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, HRF_EINVAL, "dwconv_train: dimensions");
  HRF_REQUIRE(stride == 1 || stride == 2, HRF_EUNSUPPORTED, "dwconv_train: stride %d", stride);
  HRF_REQUIRE((long long)B * C * H * ((W + 3) / 4) < (1ll << 31), HRF_EUNSUPPORTED, "dwconv_train: tensor too large");
  return HRF_OK;
}
static DwTrainParams dw_train_params(int B, int C, int H, int W, int stride) {
  DwTrainParams p{};
  p.B = B; p.C = C; p.H = H; p.W = W; p.Ho = (H - 1) / stride + 1; p.Wo = (W - 1) / stride + 1;
  return p;
}
size_t hrf_dwconv_train_ws_floats(int32_t B, int32_t C, int32_t H, int32_t W, int32_t stride) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || (stride != 1 && stride != 2)) return 0;
  return dw_train_ws_floats(B, C, (H - 1) / stride + 1, (W - 1) / stride + 1);
}
int hrf_dwconv_train_fwd(int32_t B, int32_t C, int32_t H, int32_t W, int32_t stride, const float* x, const float* w,
                         const float* bias, float* y, void* stream) {
  if (int rc = dw_train_check(B, C, H, W, stride)) return rc;
  HRF_REQUIRE(x && w && y, HRF_EINVAL, "dwconv_train_fwd: null pointer");
  DwTrainParams p = dw_train_params(B, C, H, W, stride);
  p.x = x; p.w = w; p.bias = bias; p.out = y;
  return stride == 1 ? launch_dw_train<1>(0, p, nullptr, nullptr, (cudaStream_t)stream)
                     : launch_dw_train<2>(0, p, nullptr, nullptr, (cudaStream_t)stream);
}
int hrf_dwconv_train_dgrad(int32_t B, int32_t C, int32_t H, int32_t W, int32_t stride, const float* g, const float* w,
                           float* dx, void* stream) {
  if (int rc = dw_train_check(B, C, H, W, stride)) return rc;
  HRF_REQUIRE(g && w && dx, HRF_EINVAL, "dwconv_train_dgrad: null pointer");
  DwTrainParams p = dw_train_params(B, C, H, W, stride);
  p.g = g; p.w = w; p.out = dx;
  return stride == 1 ? launch_dw_train<1>(1, p, nullptr, nullptr, (cudaStream_t)stream)
                     : launch_dw_train<2>(1, p, nullptr, nullptr, (cudaStream_t)stream);
}
int hrf_dwconv_train_wgrad(int32_t B, int32_t C, int32_t H, int32_t W, int32_t stride, const float* x, const float* g,
                           float* dw, float* dbias, float* workspace, size_t workspace_floats, void* stream) {
  if (int rc = dw_train_check(B, C, H, W, stride)) return rc;
  HRF_REQUIRE(x && g && dw && workspace, HRF_EINVAL, "dwconv_train_wgrad: null pointer");
  HRF_REQUIRE(workspace_floats >= hrf_dwconv_train_ws_floats(B, C, H, W, stride), HRF_EINVAL,
              "dwconv_train_wgrad: workspace too small");
  DwTrainParams p = dw_train_params(B, C, H, W, stride);
  p.x = x; p.g = g; p.part = workspace;
  return stride == 1 ? launch_dw_train<1>(2, p, dw, dbias, (cudaStream_t)stream)
                     : launch_dw_train<2>(2, p, dw, dbias, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ train-mode LayerNorm
size_t hrf_ln_bwd_workspace_floats(int32_t rows, int32_t C) {
  return rows > 0 && C > 0 ? ln_bwd_workspace_floats(rows, C) : 0;
}
int hrf_ln_fwd(int32_t rows, int32_t C, float eps, const float* x, const float* gamma, const float* beta,
               float* y, float* mean, float* rstd, void* stream) {
  HRF_REQUIRE(rows > 0 && C > 0 && C <= 1024, HRF_EUNSUPPORTED, "ln_fwd: rows=%d C=%d (C <= 1024)", rows, C);
  HRF_REQUIRE(x && gamma && beta && y && mean && rstd, HRF_EINVAL, "ln_fwd: null pointer");
  return launch_ln_fwd(x, gamma, beta, y, mean, rstd, rows, C, eps, (cudaStream_t)stream);
}
int hrf_ln_bwd(int32_t rows, int32_t C, const float* x, const float* dy, const float* mean, const float* rstd,
               const float* gamma, float* dx, float* dgamma, float* dbeta, float* workspace,
               size_t workspace_floats, void* stream) {
  HRF_REQUIRE(rows > 0 && C > 0 && C <= 1024, HRF_EUNSUPPORTED, "ln_bwd: rows=%d C=%d (C <= 1024)", rows, C);
  HRF_REQUIRE(x && dy && mean && rstd && gamma && dgamma && dbeta && workspace, HRF_EINVAL, "ln_bwd: null pointer");
  HRF_REQUIRE(workspace_floats >= ln_bwd_workspace_floats(rows, C), HRF_EINVAL, "ln_bwd: workspace too small");
  return launch_ln_bwd(x, dy, mean, rstd, gamma, dx, workspace, dgamma, dbeta, rows, C, (cudaStream_t)stream);
}

int hrf_bn_affine(const HrfBnDesc* d, const void* x, const void* dy, const float* a, const float* b,
                  const float* c0, int32_t relu, void* out, void* stream) {
  if (int rc = bn_check(d)) return rc;
  HRF_REQUIRE(x && a && c0 && out, HRF_EINVAL, "bn_affine: null pointer");
  HRF_REQUIRE(!dy || b, HRF_EINVAL, "bn_affine: b is required with dy");
  BnCoef k{};
  k.a = a;
  k.b = b;
  k.c0 = c0;
  if (d->dtype == HRF_F32)
    return launch_bn_affine<float, COEF_GIVEN>(d->B, d->C, d->HW, x, dy, k, relu ? BN_ACT_RELU : BN_ACT_NONE, out, (cudaStream_t)stream);
  return launch_bn_affine<__nv_bfloat16, COEF_GIVEN>(d->B, d->C, d->HW, x, dy, k, relu ? BN_ACT_RELU : BN_ACT_NONE, out, (cudaStream_t)stream);
}
int hrf_bn_normalize(const HrfBnDesc* d, const void* x, const double* stats, const float* weight,
                     const float* bias, float eps, float momentum, float* running_mean,
                     float* running_var, float* save_mean, float* save_invstd, int32_t act,
                     void* y, void* stream) {
  if (int rc = bn_check(d)) return rc;
  if (int rc = bn_act_check(act)) return rc;
  HRF_REQUIRE(x && stats && save_mean && save_invstd && y, HRF_EINVAL, "bn_normalize: null pointer");
  HRF_REQUIRE(!running_mean == !running_var, HRF_EINVAL, "bn_normalize: running_mean and running_var go together");
  HRF_REQUIRE(eps >= 0.f, HRF_EINVAL, "bn_normalize: eps");
  BnCoef k{};
  k.sums = stats;
  k.weight = weight;
  k.bias = bias;
  k.eps = eps;
  k.momentum = momentum;
  k.running_mean = running_mean;
  k.running_var = running_var;
  k.save_mean = save_mean;
  k.save_invstd = save_invstd;
  if (d->dtype == HRF_F32)
    return launch_bn_affine<float, COEF_FWD>(d->B, d->C, d->HW, x, nullptr, k, act, y, (cudaStream_t)stream);
  return launch_bn_affine<__nv_bfloat16, COEF_FWD>(d->B, d->C, d->HW, x, nullptr, k, act, y, (cudaStream_t)stream);
}
int hrf_bn_bwd_dx(const HrfBnDesc* d, const void* x, const void* dy, const double* sums,
                  const double* count, const float* weight, const float* bias, int32_t act,
                  const float* mean, const float* invstd, void* dx, void* stream) {
  if (int rc = bn_check(d)) return rc;
  if (int rc = bn_act_check(act)) return rc;
  HRF_REQUIRE(x && dy && sums && count && mean && invstd && dx, HRF_EINVAL, "bn_bwd_dx: null pointer");
  BnCoef k{};
  k.sums = sums;
  k.count = count;
  k.weight = weight;
  k.bias = bias;
  k.mean = mean;
  k.invstd = invstd;
  if (d->dtype == HRF_F32)
    return launch_bn_affine<float, COEF_BWD>(d->B, d->C, d->HW, x, dy, k, act, dx, (cudaStream_t)stream);
  return launch_bn_affine<__nv_bfloat16, COEF_BWD>(d->B, d->C, d->HW, x, dy, k, act, dx, (cudaStream_t)stream);
}

int hrf_selftest_umma(const void* A, const void* B, float* D, int32_t N, int32_t K, int32_t b_mn_major,
                      void* stream) {
  HRF_REQUIRE(A && B && D, HRF_EINVAL, "selftest: null pointer");
  return launch_umma_selftest(A, B, D, N, K, b_mn_major, (cudaStream_t)stream);
}

int hrf_input_prologue_fwd(const HrfInputDesc* d, const void* src, const float* mean,
                           const float* std, float* dst, void* stream) {
  HRF_REQUIRE(d && src && mean && std && dst, HRF_EINVAL, "input_prologue: null pointer");
  HRF_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0, HRF_EINVAL, "input_prologue: empty source");
  HRF_REQUIRE(d->C >= 1 && d->C <= INPUT_MAX_C, HRF_EUNSUPPORTED, "input_prologue: 1..4 channels, got %d", d->C);
  HRF_REQUIRE(d->Hp >= d->H && d->Wp >= d->W, HRF_EINVAL, "input_prologue: padded size %dx%d smaller than the source %dx%d",
              d->Hp, d->Wp, d->H, d->W);
  HRF_REQUIRE(d->Wp % 4 == 0, HRF_EUNSUPPORTED, "input_prologue: padded width %d is not a multiple of 4", d->Wp);
  HRF_REQUIRE(!d->to_rgb || d->C == 3, HRF_EINVAL, "input_prologue: to_rgb needs a 3-channel image");
  HRF_REQUIRE(d->src_dtype == HRF_U8 || d->src_dtype == HRF_F32, HRF_EINVAL, "input_prologue: source dtype");
  InputNorm nm{};
  for (int c = 0; c < d->C; ++c) {
    HRF_REQUIRE(std[c] != 0.f, HRF_EINVAL, "input_prologue: std[%d] is zero", c);
    nm.mean[c] = mean[c];
    nm.stdinv[c] = 1.0 / (double)std[c];   // mmcv.imnormalize: 1 / np.float64(std); cv2.multiply keeps it in fp64
  }
  if (d->src_dtype == HRF_U8)
    return launch_input_prologue_t<uint8_t>(d->B, d->H, d->W, d->C, d->Hp, d->Wp, src, nm, d->to_rgb, d->pad_val,
                                            dst, (cudaStream_t)stream);
  return launch_input_prologue_t<float>(d->B, d->H, d->W, d->C, d->Hp, d->Wp, src, nm, d->to_rgb, d->pad_val, dst,
                                        (cudaStream_t)stream);
}

int hrf_nchw_to_nhwc(int32_t B, int32_t C, int32_t H, int32_t W, int32_t sdt, const void* src,
                     int32_t ddt, void* dst, void* stream) {
  HRF_REQUIRE(src && dst && B > 0 && C > 0 && H > 0 && W > 0, HRF_EINVAL, "nchw_to_nhwc: args");
  return launch_layout<true>(B, C, H, W, sdt, src, ddt, dst, (cudaStream_t)stream);
}
int hrf_pool_fwd(int32_t B, int32_t H, int32_t W, int32_t C, int32_t k, int32_t is_max, const void* x, void* out,
                 void* stream) {
  HRF_REQUIRE(x && out && B > 0 && H > 0 && W > 0 && C > 0 && k > 0, HRF_EINVAL, "pool: args");
  HRF_REQUIRE(C % 8 == 0 && H % k == 0 && W % k == 0, HRF_EUNSUPPORTED,
              "pool: C=%d must be a multiple of 8 and %dx%d a multiple of the window %d", C, H, W, k);
  HRF_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, HRF_EINVAL,
              "pool: pointers must be 16-byte aligned");
  return launch_pool_tokens(x, out, B, H, W, C, k, is_max != 0, (cudaStream_t)stream);
}
int hrf_nhwc_to_nchw(int32_t B, int32_t C, int32_t H, int32_t W, int32_t sdt, const void* src,
                     int32_t ddt, void* dst, void* stream) {
  HRF_REQUIRE(src && dst && B > 0 && C > 0 && H > 0 && W > 0, HRF_EINVAL, "nhwc_to_nchw: args");
  return launch_layout<false>(B, C, H, W, sdt, src, ddt, dst, (cudaStream_t)stream);
}

}  // extern "C"

#ifdef HRF_KERNEL_PROFILE
// instrumented build only: phase cycle counters (see common.cuh)
extern "C" int hrf_debug_prof_cta(unsigned long long* out, int n) {
  using namespace hrf;
  HRF_CUDA(cudaMemcpyFromSymbol(out, g_prof_cta, sizeof(unsigned long long) * n));
  return HRF_OK;
}
extern "C" int hrf_debug_prof(unsigned long long* out, int n, int reset) {
  using namespace hrf;
  if (out) HRF_CUDA(cudaMemcpyFromSymbol(out, g_prof, sizeof(unsigned long long) * n));
  if (reset) {
    static unsigned long long zeros[2048 * 16];
    HRF_CUDA(cudaMemcpyToSymbol(g_prof, zeros, sizeof(zeros)));
  }
  return HRF_OK;
}
#endif
