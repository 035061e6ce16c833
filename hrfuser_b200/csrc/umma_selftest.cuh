// Diagnostic: one 128 x N x K bf16 GEMM on tcgen05 with the exact descriptor /
// layout helpers the production kernels use (umma.cuh).  Exposed through the
// C-ABI as hrf_selftest_umma so the GPU test-suite can pin the descriptor
// encodings against torch.matmul before any fused kernel depends on them.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace hrf {

__global__ void __launch_bounds__(128) umma_selftest_kernel(const __nv_bfloat16* A,
                                                            const __nv_bfloat16* B, float* D,
                                                            int N, int K, int b_mn) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  unsigned char* At = sm;                       // [K/8][128][8] bf16
  unsigned char* Bt = sm + (size_t)K * 128 * 2; // K-major: [K/8][N][8]; MN-major: [N/8][K][8]
  const int tid = threadIdx.x, warp = tid / 32;
  for (int e = tid; e < 128 * K; e += 128) {
    const int r = e / K, c = e % K;
    *reinterpret_cast<__nv_bfloat16*>(At + umma::tile_off(r, c, 128)) = A[e];
  }
  for (int e = tid; e < N * K; e += 128) {
    if (b_mn) {  // B given as [K][N]
      const int k = e / N, n = e % N;
      *reinterpret_cast<__nv_bfloat16*>(Bt + umma::tile_off(k, n, K)) = B[e];
    } else {     // B given as [N][K]
      const int n = e / K, k = e % K;
      *reinterpret_cast<__nv_bfloat16*>(Bt + umma::tile_off(n, k, N)) = B[e];
    }
  }
  uint32_t ncols = 32;
  while ((int)ncols < N) ncols <<= 1;
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, ncols);
  if (tid == 0) {
    umma::mbar_init(&bar, 1);
    umma::fence_mbar_init();
  }
  umma::fence_proxy_async();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = umma::idesc_bf16(128, N, false, b_mn != 0);
    const uint32_t a_addr = umma::smem_u32(At), b_addr = umma::smem_u32(Bt);
    for (int s = 0; s < K / 16; ++s) {
      const uint64_t da = umma::desc_kmajor(a_addr, 128, s);
      const uint64_t db = b_mn ? umma::desc_mnmajor(b_addr, K, s) : umma::desc_kmajor(b_addr, N, s);
      umma::mma_bf16(tmem, da, db, idesc, s > 0);
    }
    umma::mma_commit(&bar);
  }
  umma::mbar_wait(&bar, 0);
  umma::tc_fence_after();
  const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
  for (int c = 0; c < N; c += 8) {
    float v[8];
    umma::tmem_ld8(lane_addr + c, v);
    umma::tmem_ld_wait();
    for (int i = 0; i < 8; ++i) D[(size_t)tid * N + c + i] = v[i];
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, ncols);
}

static int launch_umma_selftest(const void* A, const void* B, float* D, int N, int K, int b_mn,
                                cudaStream_t stream) {
  HRF_REQUIRE(N % 16 == 0 && N >= 16 && N <= 256 && K % 16 == 0 && K >= 16 && K <= 256,
              HRF_EINVAL, "selftest: N=%d K=%d", N, K);
  const size_t smem = (size_t)K * 128 * 2 + (size_t)N * K * 2;
  HRF_CUDA(ensure_smem((const void*)umma_selftest_kernel, smem));
  umma_selftest_kernel<<<1, 128, smem, stream>>>((const __nv_bfloat16*)A, (const __nv_bfloat16*)B,
                                                 D, N, K, b_mn);
  count_launch();
  HRF_CUDA(cudaGetLastError());
  return HRF_OK;
}

}  // namespace hrf
