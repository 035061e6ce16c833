// Probe: 3-D tiled TMA load / store of a token box through a (W*C/2 words, H, B) map.
// nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_probe tools/probes/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../hrfuser_b200/csrc/umma.cuh"
using namespace hrf::umma;

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tm_p, const __grid_constant__ CUtensorMap tmo_p, int c0, int c1,
                      int c2, int bytes, uint32_t* dump, int mode, const CUtensorMap* gmaps) {
  const CUtensorMap& tm = gmaps ? gmaps[0] : tm_p;
  const CUtensorMap& tmo = gmaps ? gmaps[1] : tmo_p;
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) uint64_t bar;
  const int warp = warp_idx_uniform();
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (mode == 0 || mode >= 2) {
    if (threadIdx.x == 0) {
      mbar_expect_tx(&bar, bytes);
      tma_load_3d(sm, &tm, c0, c1, c2, &bar);
    }
  } else {
    if (warp == 0 && elect_one()) {
      mbar_expect_tx(&bar, bytes);
      tma_load_3d(sm, &tm, c0, c1, c2, &bar);
    }
  }
  if (mode >= 2) {                       // no blocking wait: sleep, then report the barrier state
    for (int i = 0; i < 2000; ++i) __nanosleep(1000);
    if (threadIdx.x == 0) printf("mbarrier complete after 2 ms: %d\n", (int)mbar_test(&bar, 0));
    if (mode == 3) return;
  } else {
    mbar_wait(&bar, 0);
  }
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) dump[i] = reinterpret_cast<uint32_t*>(sm)[i];
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    tma_store_3d(&tmo, c0, c1, c2, sm);
    tma_store_commit();
    tma_store_wait_all();
  }
}

int main(int argc, char** argv) {
  const int B = 2, H = 24, W = 32, C = 18;
  const int boxw = argc > 1 ? atoi(argv[1]) : 20, boxh = argc > 2 ? atoi(argv[2]) : 14;
  const int mode = argc > 3 ? atoi(argv[3]) : 0;
  const int dtype_sel = argc > 4 ? atoi(argv[4]) : 0;     // 0: u32 words, 1: bf16 elements
  std::vector<uint16_t> h((size_t)B * H * W * C);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint16_t)(i & 0xffff);
  uint16_t *d, *o;
  cudaMalloc(&d, h.size() * 2);
  cudaMalloc(&o, h.size() * 2);
  cudaMemset(o, 0, h.size() * 2);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  void* f = nullptr;
  cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr);
  printf("entry point %p qr %d\n", f, (int)qr);
  PFN_enc enc = (PFN_enc)f;
  CUtensorMap tm, tmo;
  const int es = dtype_sel ? 2 : 4, per_tok = C * 2 / es;
  cuuint64_t gdim[3] = {(cuuint64_t)W * per_tok, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gstr[2] = {(cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[3] = {(cuuint32_t)(boxw * per_tok), (cuuint32_t)boxh, 1};
  cuuint32_t est[3] = {1, 1, 1};
  const CUtensorMapDataType dt = dtype_sel ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT32;
  CUresult r = enc(&tm, dt, 3, d, gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = enc(&tmo, dt, 3, o, gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode box %dx%d dtype %d -> %d %d\n", boxw, boxh, dtype_sel, (int)r, (int)r2);
  const int bytes = boxw * C * 2 * boxh;
  uint32_t* dump;
  cudaMalloc(&dump, bytes);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  CUtensorMap* gmaps = nullptr;
  if (argc > 5 && atoi(argv[5])) {
    cudaMalloc(&gmaps, 2 * sizeof(CUtensorMap));
    cudaMemcpy(gmaps, &tm, sizeof(tm), cudaMemcpyHostToDevice);
    cudaMemcpy(gmaps + 1, &tmo, sizeof(tm), cudaMemcpyHostToDevice);
  }
  printf("descriptor words: ");
  for (int i = 0; i < 16; ++i) printf("%016llx ", (unsigned long long)reinterpret_cast<uint64_t*>(&tm)[i]);
  printf("\n");
  const int c0w = argc > 6 ? atoi(argv[6]) : -per_tok, c1r = argc > 7 ? atoi(argv[7]) : -1;
  printf("coords %d %d\n", c0w, c1r);
  probe<<<1, 128, 65536>>>(tm, tmo, c0w, c1r, 1, bytes, dump, mode, gmaps);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<uint32_t> hd(bytes / 4);
  cudaMemcpy(hd.data(), dump, bytes, cudaMemcpyDeviceToHost);
  // expected: row r, token t of the box = image 1, h = r - 1, w = t - 1 (zeros outside)
  int bad = 0;
  for (int r_ = 0; r_ < boxh; ++r_)
    for (int t = 0; t < boxw; ++t)
      for (int c = 0; c < C; ++c) {
        const int hh = r_ - 1, ww = t - 1;
        uint16_t want = 0;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) want = h[(((size_t)1 * H + hh) * W + ww) * C + c];
        const uint16_t got = reinterpret_cast<uint16_t*>(hd.data())[((size_t)r_ * boxw + t) * C + c];
        if (got != want && bad++ < 5) printf("mismatch r %d t %d c %d: got %u want %u\n", r_, t, c, got, want);
      }
  printf("load mismatches: %d\n", bad);
  std::vector<uint16_t> ho(h.size());
  cudaMemcpy(ho.data(), o, h.size() * 2, cudaMemcpyDeviceToHost);
  int bad2 = 0;
  for (size_t i = 0; i < h.size(); ++i) {
    const int c = i % C; const int w = (i / C) % W, hh = (i / C / W) % H, b = i / C / W / H; (void)c;
    const bool in = b == 1 && hh >= -1 && hh < boxh - 1 && w >= -1 && w < boxw - 1;
    const uint16_t want = in ? h[i] : 0;
    if (ho[i] != want && bad2++ < 5) printf("store mismatch at %zu: got %u want %u\n", i, ho[i], want);
  }
  printf("store mismatches: %d\n", bad2);
  return 0;
}
