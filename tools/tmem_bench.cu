// tmem_bench.cu -- developer tool: throughput / latency of the tcgen05 primitives the fused SA / FP
// kernels are built from, measured on the target GPU (SM cycles).  Not part of the product library.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/bin/tmem_bench tools/tmem_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../bridgeqa_b200/csrc/common.cuh"
#include "../bridgeqa_b200/csrc/tcgen05.cuh"
namespace bqa { int set_error(int c, const char *, ...) { return c; } void count_launch(int) {} int check_launch(const char *) { return 0; } int ref_opt_n_threads(int) { return 512; } }
using namespace bqa;

// ---- 1. tcgen05.ld throughput: W warps, each streams its lane quarter ------------------
__global__ void k_ldtm(unsigned long long *out, int iters, int two) {
  __shared__ uint32_t s_tmem;
  if (threadIdx.x < 32) umma::tmem_alloc(smem_u32(&s_tmem), 512);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = s_tmem;
  const int warp = (threadIdx.x >> 5) & 3;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t va[32], vb[32];
    const uint32_t col = (uint32_t)((i * 64) & 511);
    umma::ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + col, va);
    if (two) umma::ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + ((col + 32) & 511), vb);
    umma::wait_ld();
#pragma unroll
    for (int e = 0; e < 32; ++e) acc ^= va[e] ^ (two ? vb[e] : 0u);
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = (unsigned long long)(t1 - t0); }
  if (acc == 0x12345678u) out[1] = acc;
  umma::fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) umma::tmem_dealloc(tmem, 512);
}

// ---- 2. MMA issue -> commit -> wait round trip ------------------------------------------
template <int N>
__global__ void k_mma(unsigned long long *out, int nmma, int reps) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  for (int i = threadIdx.x; i < (64 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init_cluster(); }
  if (threadIdx.x < 32) umma::tmem_alloc(smem_u32(&s_tmem), 512);
  umma::fence_proxy_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = s_tmem;
  const uint32_t a_addr = smem_u32(smem), b_addr = a_addr + 32 * 1024;
  const uint32_t idesc = umma::instr_desc_16b_f32(128, N, 0);
  uint32_t phase = 0;
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int k = 0; k < nmma; ++k) {
        const uint64_t ad = umma::smem_desc(a_addr + (uint32_t)(2 * (k & 7)) * 128 * 16, 128 * 16, 128);
        const uint64_t bd = umma::smem_desc(b_addr + (uint32_t)(2 * (k & 3)) * N * 16, N * 16, 128);
        umma::mma_bf16_ss(tmem, ad, bd, idesc, k != 0);
      }
      umma::commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), phase);
      phase ^= 1;
      umma::fence_after_sync();
    }
    const long long t1 = clock64();
    out[0] = (unsigned long long)(t1 - t0);
  }
  umma::fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) umma::tmem_dealloc(tmem, 512);
}

// ---- 3. epilogue chain of ONE warp: ld -> wait -> cvt/pack -> st.shared -> proxy fence ---
__global__ void k_epi(unsigned long long *out, int iters, int nwarps_active) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint32_t s_tmem;
  if (threadIdx.x < 32) umma::tmem_alloc(smem_u32(&s_tmem), 512);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = s_tmem;
  const int warp = (threadIdx.x >> 5) & 3;
  uint4 *x = reinterpret_cast<uint4 *>(smem) + (threadIdx.x >> 7) * 2048;
  const int row = threadIdx.x & 127;
  __syncthreads();
  const long long t0 = clock64();
  if ((threadIdx.x >> 5) < nwarps_active) {
    for (int i = 0; i < iters; ++i) {
      uint32_t va[32], vb[32];
      const uint32_t col = (uint32_t)((i * 64) & 511);
      umma::ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + col, va);
      umma::ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + ((col + 32) & 511), vb);
      umma::wait_ld();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t p[4], r[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(p[e]) : "f"(__uint_as_float(va[q * 8 + 2 * e + 1]) + 1.f), "f"(__uint_as_float(va[q * 8 + 2 * e]) + 1.f));
          asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r[e]) : "f"(__uint_as_float(vb[q * 8 + 2 * e + 1]) + 1.f), "f"(__uint_as_float(vb[q * 8 + 2 * e]) + 1.f));
        }
        x[q * 128 + row] = make_uint4(p[0], p[1], p[2], p[3]);
        x[(4 + q) * 128 + row] = make_uint4(r[0], r[1], r[2], r[3]);
      }
    }
    umma::fence_proxy_async_smem();
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
  umma::fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) umma::tmem_dealloc(tmem, 512);
}

// ---- 4. dependent global gather latency: idx -> row (cp.async 16 B x chunks) -------------
__global__ void k_gather(unsigned long long *out, const int *idx, const uint4 *rows, int chunks, int iters) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint4 *a = reinterpret_cast<uint4 *>(smem);
  const int row = threadIdx.x & 127;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    const int r = __ldg(idx + (size_t)(blockIdx.x * iters + i) * 128 + row);
    const uint4 *src = rows + (size_t)r * chunks;
    for (int q = 0; q < chunks; ++q) {
      const uint32_t dst = smem_u32(a + q * 128 + row);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + q) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
}

int main() {
  unsigned long long *d, h[2];
  cudaMalloc(&d, 64);
  printf("== tcgen05.ld 32x32b.x32 throughput (one CTA on one SM)\n");
  for (int two = 0; two <= 1; ++two)
    for (int warps : {1, 2, 4, 8, 16}) {
      const int iters = 2048;
      k_ldtm<<<1, warps * 32>>>(d, iters, two);
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      const double bytes = (double)warps * iters * (two ? 2 : 1) * 4096.0;
      printf("  warps=%2d loads/wait=%d : %8llu cycles, %.1f B/cycle/SM, %.1f cycles per wait\n", warps, two + 1, h[0],
             bytes / (double)h[0], (double)h[0] / iters);
    }
  printf("== MMA (M=128, K=16) issue -> commit -> wait, per round (single thread)\n");
  cudaFuncSetAttribute(k_mma<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(k_mma<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(k_mma<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int n : {64, 128, 256})
    for (int nm : {1, 2, 4, 8, 16, 32}) {
      const int reps = 200;
      if (n == 64) k_mma<64><<<1, 128, 64 * 1024>>>(d, nm, reps);
      if (n == 128) k_mma<128><<<1, 128, 64 * 1024>>>(d, nm, reps);
      if (n == 256) k_mma<256><<<1, 128, 64 * 1024>>>(d, nm, reps);
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("  N=%3d mmas=%2d : %.0f cycles per round (floor %d)\n", n, nm, (double)h[0] / reps, nm * 128 * n / 256);
    }
  printf("== epilogue chain (2 x ld x32 -> wait -> 64 add+cvt -> 8 st.shared.v4), per 64 columns\n");
  cudaFuncSetAttribute(k_epi, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  for (int w : {1, 4, 8, 16}) {
    const int iters = 1024;
    k_epi<<<1, 512, 128 * 1024>>>(d, iters, w);
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("  active warps=%2d : %.1f cycles per 64-column step per warp, %.1f fp32/cycle/SM\n", w, (double)h[0] / iters,
           (double)w * iters * 32 * 64 / (double)h[0]);
  }
  printf("== gather: idx -> cp.async 16 B x chunks per row, 128 rows per CTA round (L2-resident table of 32768 rows)\n");
  {
    const int rows = 32768, iters = 64, ctas = 148;
    for (int chunks : {2, 16, 32}) {
      int *idx; uint4 *tab;
      cudaMalloc(&idx, (size_t)ctas * iters * 128 * 4);
      cudaMalloc(&tab, (size_t)rows * chunks * 16);
      cudaMemset(tab, 0, (size_t)rows * chunks * 16);
      int *hidx = new int[ctas * iters * 128];
      uint32_t s = 12345;
      for (int i = 0; i < ctas * iters * 128; ++i) { s = s * 1664525u + 1013904223u; hidx[i] = (s >> 8) % rows; }
      cudaMemcpy(idx, hidx, (size_t)ctas * iters * 128 * 4, cudaMemcpyHostToDevice);
      cudaFuncSetAttribute(k_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
      for (int c : {1, 148}) {
        k_gather<<<c, 128, 64 * 1024>>>(d, idx, tab, chunks, iters);
        k_gather<<<c, 128, 64 * 1024>>>(d, idx, tab, chunks, iters);
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("  chunks=%2d (%4d B/row) ctas=%3d : %.0f cycles per 128-row round\n", chunks, chunks * 16, c, (double)h[0] / iters);
      }
      cudaFree(idx); cudaFree(tab); delete[] hidx;
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("done: %s\n", cudaGetErrorString(e));
  return 0;
}
