// tcgen05 / TMEM / mbarrier primitives (inline PTX, sm_100a) used by the bf16
// tensor-core kernels.  One CTA per SM-slice, cta_group::1 everywhere.
//
// Operand tiles live in shared memory in the no-swizzle canonical layout of the
// UMMA shared-memory descriptor, organised as "chunk-major":
//
//     element (row, col) of an [R rows x Ccols] bf16 tile  ->
//         byte offset  (col/8) * (R*16)  +  row*16  +  (col%8)*2
//
// i.e. for every 8-element chunk of the contiguous dimension, all R rows of 16
// bytes, so an 8-row x 16-byte core matrix is 128 contiguous bytes.
//   * used as a K-major operand (rows = M or N index, contiguous = K):
//        LBO (next K chunk) = R*16,  SBO (next 8 rows) = 128
//   * used as an MN-major operand (rows = K index, contiguous = M or N):
//        SBO (next MN chunk) = R*16, LBO (next 8 K rows) = 128
// One thread writing its own row's 16-byte chunks is bank-conflict free.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

namespace hrf {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// blocking probe: the hardware suspends the thread until the phase completes or an
// implementation-defined time limit expires (returns false on the limit)
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a lost commit must become an error (trap), never a hung GPU.
// -DHRF_DEBUG_WAIT (debug library only): a short limit, and the timed-out wait names itself.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag = -1) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
#ifdef HRF_DEBUG_WAIT
    if (clock64() - t0 > 200000000LL) {
      printf("[hrf] mbarrier wait timed out: tag %d parity %u block %d thread %d\n", tag, parity,
             (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
#else
    if (clock64() - t0 > 4000000000LL) __trap();
#endif
  }
}
// Whole-CTA wait for an MMA commit: only warp 0 watches the mbarrier (suspended in
// try_wait); every other warp sleeps in the hardware barrier instead of spinning on
// test_wait and stealing issue slots from the co-resident CTAs.
__device__ __forceinline__ void cta_wait(uint64_t* bar, uint32_t parity) {
  if ((threadIdx.x >> 5) == 0) mbar_wait(bar, parity);
  __syncthreads();
}

// ---- bulk async copy global -> shared (TMA, 1-D), completion on an mbarrier ---------------
// `bytes` multiple of 16, both addresses 16-byte aligned.  One thread: expect_tx(total) once,
// then any number of copies; waiters use mbar_wait(bar, phase).
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---- TMA tensor-tile copies (cp.async.bulk.tensor, 3-D tiled maps built by abi.cu) ---------
// The token tensors [B][H][W][C] are described to the TMA unit as 3-D tensors
// (W*C/2 32-bit words, H, B): a box of `boxw` tokens x `boxh` rows lands in shared memory as
// dense rows, and everything outside the image (negative or too large coordinates) arrives
// as zeros -- the halo / window padding needs no address arithmetic and no predicates.
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst_smem, const void* tmap, int c0, int c1, int c2,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst_smem)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void* tmap, int c0, int c1, int c2,
                                             const void* src_smem) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tmap),
               "r"(smem_u32(src_smem)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the shared-memory source of every committed store has been read (the buffer may be rewritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// every committed store is complete (globally visible at kernel end anyway; used before exit)
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- single-thread issue -----------------------------------------------------------
// tcgen05.mma / commit are issued by one thread.  Guarding them with `threadIdx.x == 0` makes
// the branch thread-divergent for the compiler, which then wraps every UTCHMMA in an
// ELECT/BRA loop over the active lanes (~65 cycles per MMA).  A warp-uniform warp index plus
// elect.sync keeps the region uniform: one UTCHMMA, no loop.
__device__ __forceinline__ int warp_idx_uniform() {
  return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
}
__device__ __forceinline__ bool elect_one() {     // all 32 lanes of the warp must be converged
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- fences --------------------------------------------------------------------
// generic-proxy shared-memory writes -> visible to the async proxy (UMMA operand reads)
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---- TMEM allocation (one warp, power-of-two columns >= 32) --------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}

// ---- descriptors ---------------------------------------------------------------
// shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (Blackwell)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes,
                                              uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor for kind::f16 with bf16 inputs and fp32 accumulation
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) |
         ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// the same with fp16 inputs (the MixFFN hidden path: H2 x W2)
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                         uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// arrive on an mbarrier when all MMAs issued so far by this thread have completed
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// ---- TMEM -> registers: 32 lanes x 32 bit, N consecutive columns per thread ------
// thread t of warp w reads lane 32*(w%4)+t; taddr = base | lane<<16 | column
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
                 "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

// ---- chunk-major operand tiles ---------------------------------------------------
// byte offset of element (row, col) in an R-row tile
__host__ __device__ constexpr uint32_t tile_off(int row, int col, int R) {
  return (uint32_t)((col >> 3) * (R * 16) + row * 16 + (col & 7) * 2);
}
// store 8 consecutive fp32 values of one row as one 16-byte bf16 chunk
__device__ __forceinline__ void st_chunk(unsigned char* tile, int row, int chunk, int R,
                                         const float* v) {
  uint4 u;
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  u.z = *reinterpret_cast<uint32_t*>(&c);
  u.w = *reinterpret_cast<uint32_t*>(&d);
  *reinterpret_cast<uint4*>(tile + (size_t)chunk * (R * 16) + row * 16) = u;
}
// descriptor of the K=16 step `s` of a K-major operand tile (R rows)
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile_saddr, int R, int kstep) {
  return smem_desc(tile_saddr + (uint32_t)kstep * 2u * (uint32_t)(R * 16), (uint32_t)(R * 16), 128u);
}
// descriptor of the K=16 step `s` of an MN-major operand tile whose rows are the K index
// (R = number of K rows in the tile)
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile_saddr, int R, int kstep) {
  return smem_desc(tile_saddr + (uint32_t)kstep * 256u, 128u, (uint32_t)(R * 16));
}

}  // namespace umma
}  // namespace hrf
