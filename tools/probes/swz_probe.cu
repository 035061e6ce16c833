// Probe: does a 128-byte-swizzled K-major UMMA operand descriptor accept (a) a start address that is
// a multiple of 128 bytes but not of 1024 (a token shift inside the swizzle atom) and (b) a stride
// between 8-row groups that is not a multiple of 1024 (a halo row pitch of 10 tokens = 1280 bytes)?
// The A region is written the way TMA writes a SWIZZLE_128B box: 16-byte chunk index XOR (address bits 7..9).
// nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/probes/swz_probe.bin tools/probes/swz_probe.cu
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../hrfuser_b200/csrc/umma.cuh"
using namespace hrf::umma;

constexpr int TOK = 192, N = 32;

__host__ __device__ inline int a_val(int t, int c) { return ((t * 7 + c * 3) % 17) - 8; }
__host__ __device__ inline int b_val(int n, int k) { return ((n * 5 + k) % 13) - 6; }

__device__ __forceinline__ uint64_t desc(uint32_t saddr, int ks, uint32_t sbo, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)(((saddr + (uint32_t)ks * 32u) & 0x3FFFF) >> 4);
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(128) probe(int shift, int sbo, int bo_mode, float* out) {
  extern __shared__ unsigned char raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_s;
  unsigned char* sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  unsigned char* A = sm;                       // TOK rows of 128 bytes
  unsigned char* B = sm + TOK * 128;           // 32 rows of 128 bytes (TOK * 128 is a multiple of 1024)
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < TOK * 64; e += 128) {
    const int t = e / 64, c = e % 64;
    const uint32_t row_addr = smem_u32(A) + t * 128;
    const uint32_t off = t * 128 + (((c >> 3) ^ ((row_addr >> 7) & 7)) << 4) + (c & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(A + off) = __float2bfloat16((float)a_val(t, c));
  }
  for (int e = tid; e < N * 64; e += 128) {
    const int n = e / 64, k = e % 64;
    const uint32_t off = n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(B + off) = __float2bfloat16((float)b_val(n, k));
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_s, 32);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_s;
  if (warp == 0 && elect_one()) {
    const uint32_t a0 = smem_u32(A) + shift * 128;
    const uint32_t bo = bo_mode ? (a0 >> 7) & 7 : 0;
    constexpr uint32_t idesc = idesc_bf16(128, N, false, false);
    for (int ks = 0; ks < 4; ++ks)
      mma_bf16(tmem, desc(a0, ks, sbo, bo), desc(smem_u32(B), ks, 1024, 0), idesc, ks != 0);
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  float v[32];
  tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld_wait();
  for (int n = 0; n < N; ++n) out[tid * N + n] = v[n];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

int main() {
  float* d;
  cudaMalloc(&d, 128 * N * 4);
  const int smem = TOK * 128 + N * 128 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> h(128 * N);
  const int cases[][3] = {{0, 1024, 0}, {1, 1024, 0}, {1, 1024, 1}, {3, 1024, 0}, {3, 1024, 1}, {8, 1024, 0},
                          {0, 1280, 0}, {1, 1280, 0}, {1, 1280, 1}, {2, 1280, 0}, {2, 1280, 1}, {11, 1280, 0},
                          {11, 1280, 1}, {0, 2304, 0}, {1, 2304, 0}, {1, 2304, 1}};
  for (auto& cs : cases) {
    const int shift = cs[0], sbo = cs[1], bo = cs[2];
    probe<<<1, 128, smem>>>(shift, sbo, bo, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("shift %d sbo %d bo %d: %s\n", shift, sbo, bo, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0, bad_first_group = 0;
    for (int m = 0; m < 128; ++m) {
      const int t = shift + (m / 8) * (sbo / 128) + m % 8;      // the token a linear-address reading gives row m
      for (int n = 0; n < N; ++n) {
        float ref = 0;
        for (int k = 0; k < 64; ++k) ref += (float)(a_val(t, k) * b_val(n, k));
        if (h[m * N + n] != ref) { ++bad; if (m < 8) ++bad_first_group; }
      }
    }
    printf("shift %2d  sbo %4d  base_offset %s : %4d of %d wrong (%d in the first 8 rows)\n", shift, sbo,
           bo ? "(start>>7)&7" : "0           ", bad, 128 * N, bad_first_group);
  }
  return 0;
}
