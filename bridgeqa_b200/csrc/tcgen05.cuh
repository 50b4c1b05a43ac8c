// tcgen05.cuh -- thin inline-PTX wrappers for Blackwell 5th-gen tensor cores (sm_100a):
// TMEM allocation, UMMA shared-memory / instruction descriptors, tcgen05.mma issue,
// commit -> mbarrier, and TMEM -> register loads for epilogues.
//
// Shared-memory operand layout used throughout (K-major, no swizzle, "chunk-major"):
// a tile of R rows x K bf16 is stored as [K/8][R] 16-byte vectors, i.e. element (r,k)
// lives at byte ((k/8)*R + r)*16 + (k%8)*2.  In UMMA terms the 8x16B core matrices are
// 128 B contiguous, neighbouring core matrices along M/N are 128 B apart (SBO) and
// neighbouring core matrices along K are R*16 B apart (LBO).  A thread that owns one row
// writes 16-byte vectors that are contiguous across the warp (no bank conflicts).
#pragma once
#include <stdint.h>

namespace bqa {
namespace umma {

// ---- descriptors ------------------------------------------------------------------
// 64-bit shared-memory matrix descriptor (PTX ISA "tcgen05 matrix descriptor"):
//  [0,14) start address >> 4, [16,30) leading-dim byte offset >> 4,
//  [32,46) stride-dim byte offset >> 4, [46,48) version = 1 on sm_100, [61,64) swizzle = 0.
__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr, uint32_t lbo_bytes,
                                              uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}

// K-major operand in the 128-byte-swizzle layout (row pitch 128 B, 16-byte chunk index ^ (row & 7), 8-row groups
// `sbo_bytes` apart, buffer 1024-byte aligned): layout type 2 in [61,64); the leading offset is unused (1).
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t smem_addr, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3fffu) | (1ull << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46) | (2ull << 61);
}

// 32-bit instruction descriptor for kind::f16 with 16-bit A/B (both K-major) and FP32 D:
//  [4,6) D format = 1 (f32), [7,10) A format, [10,13) B format (0 = f16, 1 = bf16),
//  bit 15 / 16 = A / B major (0 = K), [17,23) N >> 3, [24,29) M >> 4.
__host__ __device__ constexpr uint32_t instr_desc_16b_f32(int m, int n, int is_bf16) {
  return (1u << 4) | ((uint32_t)is_bf16 << 7) | ((uint32_t)is_bf16 << 10) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---- TMEM allocation (one full warp executes these) --------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_result_addr, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
               ::"r"(smem_result_addr), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}

// ---- ordering -----------------------------------------------------------------------
__device__ __forceinline__ void fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// make generic-proxy shared-memory writes (st.shared) visible to the async proxy that
// tcgen05.mma reads operands through
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- MMA issue (ONE thread) -----------------------------------------------------------
// D[tmem] (+)= A[smem] * B[smem]^T ; accumulate = 0 overwrites D.
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every previously issued tcgen05.mma of this thread is done
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void commit(uint32_t mbar_smem_addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               ::"r"(mbar_smem_addr) : "memory");
}

// ---- TMEM -> registers -----------------------------------------------------------------
// 32 lanes x 32 columns of 32-bit: thread t of the warp receives columns [col, col+32) of
// TMEM lane (lane_base + t); taddr = tmem_base + (lane_base << 16) + col.  A warp may only
// touch lanes [32*(warp_id % 4), +32).
__device__ __forceinline__ void ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

}  // namespace umma
}  // namespace bqa
