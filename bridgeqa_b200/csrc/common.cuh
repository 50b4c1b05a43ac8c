// common.cuh -- shared host/device helpers for libbqa_pointnet2.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bqa_pointnet2.h"

namespace bqa {

// ---- host-side error plumbing (api.cu owns the storage) ------------------------
int set_error(int code, const char *fmt, ...);
void count_launch(int n = 1);
int check_launch(const char *what);  // cudaGetLastError -> BQA_ERR_CUDA

#define BQA_REQUIRE(cond, ...)                                            \
  do {                                                                    \
    if (!(cond)) return ::bqa::set_error(BQA_ERR_INVALID_ARG, __VA_ARGS__); \
  } while (0)

#define BQA_CUDA(call)                                                                 \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess)                                                            \
      return ::bqa::set_error(BQA_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__)); \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// cuda_utils.h:15-19 of the reference (opt_n_threads): the block size the reference
// would have launched with; FPS needs it because the reduction-tree tie-break depends
// on it.
int ref_opt_n_threads(int work_size);

// ---- device helpers ------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n"
               "barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// shared::cta address -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void fence_mbar_init_cluster() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// 16-byte / 4-byte stores into another CTA's shared memory that complete a transaction
// on that CTA's mbarrier (data and the "it has landed" signal travel together).
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, uint32_t a, uint32_t b,
                                            uint32_t c, uint32_t d, uint32_t remote_bar) {
  asm volatile(
      "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
      ::"r"(remote_addr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(remote_bar)
      : "memory");
}

__device__ __forceinline__ void st_async_b32(uint32_t remote_addr, uint32_t a,
                                             uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
               ::"r"(remote_addr), "r"(a), "r"(remote_bar)
               : "memory");
}

// the reference's distance: nvcc 12.9 contracts dx*dx + dy*dy + dz*dz to
// fma(dz,dz, fma(dx,dx, dy*dy)) -- the middle product is the lone FMUL (reference SASS)
__device__ __forceinline__ float sqdist3(float ax, float ay, float az, float bx, float by,
                                         float bz) {
  const float dx = ax - bx, dy = ay - by, dz = az - bz;
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

#endif  // __CUDACC__

}  // namespace bqa
