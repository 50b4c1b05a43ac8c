// ubench.cu -- developer tool: latencies (SM cycles) of the primitives the FPS argmax chain
// is built from, measured on the target GPU.  Not part of the product library.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/bin/ubench tools/ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../bridgeqa_b200/csrc/common.cuh"
namespace bqa { int set_error(int c, const char *, ...) { return c; } void count_launch(int) {} int check_launch(const char *) { return 0; } int ref_opt_n_threads(int) { return 512; } }
using namespace bqa;

constexpr int kIters = 256;

__global__ void k_chain(unsigned long long *out, int mode) {
  __shared__ unsigned long long s64[32];
  __shared__ uint32_t s32[64];
  const int lane = threadIdx.x & 31;
  uint32_t v = threadIdx.x * 2654435761u;
  s32[lane] = v; s32[lane + 32] = v;
  s64[lane] = v;
  __syncthreads();
  long long t0 = clock64();
  if (mode == 0) {
    for (int i = 0; i < kIters; ++i) v = __reduce_max_sync(0xffffffffu, v + i) ^ lane;
  } else if (mode == 1) {
    for (int i = 0; i < kIters; ++i) v = max(v, __shfl_xor_sync(0xffffffffu, v, 16)) + i;
  } else if (mode == 2) {
    for (int i = 0; i < kIters; ++i) v = __ffs(__ballot_sync(0xffffffffu, (v + i) & 1)) + v;
  } else if (mode == 3) {
    for (int i = 0; i < kIters; ++i) v = s32[(v + i) & 63];
  } else if (mode == 4) {   // 64-bit smem atomic max, one lane per warp
    for (int i = 0; i < kIters; ++i) {
      if (lane == 0) v += (uint32_t)atomicMax(&s64[0], (unsigned long long)(v + i));
    }
  } else if (mode == 5) {   // 64-bit butterfly argmax (5 steps, 2 shfl each)
    unsigned long long p = ((unsigned long long)v << 32) | lane;
    for (int i = 0; i < kIters / 8; ++i) {
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        unsigned long long q = __shfl_xor_sync(0xffffffffu, p, o);
        p = q > p ? q : p;
      }
      p += i;
    }
    v = (uint32_t)p;
  } else if (mode == 6) {   // match_any style: redux then ballot+shfl
    for (int i = 0; i < kIters; ++i) {
      uint32_t m = __reduce_max_sync(0xffffffffu, v + i);
      int src = __ffs(__ballot_sync(0xffffffffu, v + i == m)) - 1;
      v = __shfl_sync(0xffffffffu, v ^ lane, src);
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = v; }
}

__global__ void k_bar(unsigned long long *out) {
  long long t0 = clock64();
  for (int i = 0; i < kIters; ++i) __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
}

// named-barrier pair: producer warps arrive, consumer warp syncs (no full-CTA barrier)
__global__ void k_smem_flag(unsigned long long *out) {
  __shared__ volatile uint32_t flag[2];
  if (threadIdx.x == 0) { flag[0] = 0; flag[1] = 0; }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long t0 = clock64();
  if (warp == 0) {
    for (int i = 1; i <= kIters; ++i) {
      if (lane == 0) { flag[0] = i; while (flag[1] != (uint32_t)i) {} }
      __syncwarp();
    }
  } else if (warp == 1) {
    for (int i = 1; i <= kIters; ++i) {
      if (lane == 0) { while (flag[0] != (uint32_t)i) {} flag[1] = i; }
      __syncwarp();
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
}

// cluster ping-pong with st.async + mbarrier: round trip between CTA 0 and CTA (cs-1)
__global__ void k_pingpong(unsigned long long *out, int cs) {
  __shared__ __align__(16) uint32_t slot[8];
  __shared__ __align__(8) uint64_t bars[2];
  const uint32_t rank = cluster_ctarank();
  const uint32_t bar = smem_u32(&bars[0]);
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 8, 1); fence_mbar_init_cluster(); }
  cluster_sync_all();
  const uint32_t peer = rank == 0 ? cs - 1 : 0;
  long long t0 = clock64();
  if (threadIdx.x == 0 && (rank == 0 || rank == (uint32_t)cs - 1)) {
    for (int i = 0; i < kIters; ++i) {
      const uint32_t b = bar + 8 * (i & 1);
      mbar_arrive_expect_tx(b, 16);
      if (rank == 0) {
        st_async_v4(mapa_shared(smem_u32(slot), peer), i, 1, 2, 3, mapa_shared(b, peer));
        mbar_wait(b, (i >> 1) & 1);
      } else {
        mbar_wait(b, (i >> 1) & 1);
        st_async_v4(mapa_shared(smem_u32(slot), peer), i, 1, 2, 3, mapa_shared(b, peer));
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && rank == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  cluster_sync_all();
}

// cluster ping-pong with plain remote stores (st.shared::cluster.u64 carrying a sequence tag)
// and local volatile polling: no mbarrier
__global__ void k_pingpong_poll(unsigned long long *out, int cs) {
  __shared__ __align__(16) volatile unsigned long long slot[4];
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x < 4) slot[threadIdx.x] = 0;
  cluster_sync_all();
  const uint32_t peer = rank == 0 ? cs - 1 : 0;
  const uint32_t remote = mapa_shared(smem_u32((const void *)&slot[0]), peer);
  long long t0 = clock64();
  if (threadIdx.x == 0 && (rank == 0 || rank == (uint32_t)cs - 1)) {
    for (unsigned long long i = 1; i <= kIters; ++i) {
      if (rank == 0) {
        asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(remote), "l"(i) : "memory");
        while (slot[0] != i) {}
      } else {
        while (slot[0] != i) {}
        asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(remote), "l"(i) : "memory");
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && rank == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = slot[0]; }
  cluster_sync_all();
}

static void run_cluster(void (*kern)(unsigned long long *, int), const char *name, int cs, unsigned long long *d) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cs); cfg.blockDim = dim3(128);
  cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension;
  a[0].val.clusterDim.x = cs; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
  cfg.attrs = a; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kern, d, cs);
  unsigned long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-44s cs=%d  %7.1f cycles/iter  (%s) val=%llx\n", name, cs, (double)h[0] / kIters, cudaGetErrorString(cudaGetLastError()), h[1]);
}

int main() {
  unsigned long long *d, h[2];
  cudaMalloc(&d, 16);
  const char *names[] = {"redux.sync max u32 (dependent)", "shfl_xor + max (dependent)", "ballot + ffs (dependent)",
                         "LDS (dependent)", "ATOMS max.u64 lane0", "64-bit butterfly argmax /8", "redux+ballot+shfl"};
  for (int mode = 0; mode < 7; ++mode) {
    k_chain<<<1, 32>>>(d, mode);
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%-44s %7.1f cycles/iter\n", names[mode], (double)h[0] / (mode == 5 ? kIters / 8 : kIters));
  }
  for (int warps : {1, 4, 8, 16, 32}) {
    k_bar<<<1, warps * 32>>>(d);
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("__syncthreads %2d warps                       %7.1f cycles/iter\n", warps, (double)h[0] / kIters);
  }
  k_smem_flag<<<1, 64>>>(d);
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("smem flag ping-pong between 2 warps (round trip) %7.1f cycles\n", (double)h[0] / kIters);
  for (int cs : {2, 4, 8}) run_cluster(k_pingpong, "st.async+mbarrier ping-pong (round trip)", cs, d);
  for (int cs : {2, 4, 8}) run_cluster(k_pingpong_poll, "remote st.u64 + volatile poll (round trip)", cs, d);
  int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
  printf("SMs=%d\n", p.multiProcessorCount);
  return 0;
}
