// l2_lat.cu -- developer micro-benchmark: dependent-load latency from L1 / L2 / HBM and redux.sync / ballot latency on B200.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/bin/l2_lat tools/l2_lat.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstdint>
#include <algorithm>
#include <random>

__global__ void chase(const uint32_t *__restrict__ next, int hops, int mode, long long *out, uint32_t *sink) {
  uint32_t i = threadIdx.x;
  // warm
  for (int h = 0; h < 64; ++h) i = mode ? __ldcg(&next[i]) : __ldg(&next[i]);
  long long t0 = clock64();
  for (int h = 0; h < hops; ++h) i = mode ? __ldcg(&next[i]) : __ldg(&next[i]);
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = t1 - t0; }
  sink[threadIdx.x] = i;
}

__global__ void redux_lat(long long *out, uint32_t *sink) {
  uint32_t v = threadIdx.x * 2654435761u;
  long long t0 = clock64();
#pragma unroll 1
  for (int h = 0; h < 1024; ++h) v = __reduce_max_sync(0xffffffffu, v) + threadIdx.x;
  long long t1 = clock64();
#pragma unroll 1
  for (int h = 0; h < 1024; ++h) v = __ballot_sync(0xffffffffu, v & 1) + threadIdx.x;
  long long t2 = clock64();
#pragma unroll 1
  for (int h = 0; h < 1024; ++h) v = __shfl_sync(0xffffffffu, v, (v + 1) & 31) + threadIdx.x;
  long long t3 = clock64();
  if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; }
  sink[threadIdx.x] = v;
}

int main() {
  long long *out; uint32_t *sink;
  cudaMalloc(&out, 64); cudaMalloc(&sink, 4096);
  for (size_t bytes : {16384ul, 8ul << 20, 32ul << 20, 1024ul << 20}) {
    size_t n = bytes / 4;
    // random cyclic permutation over 128-byte lines
    size_t lines = n / 32;
    std::vector<uint32_t> perm(lines);
    for (size_t k = 0; k < lines; ++k) perm[k] = (uint32_t)k;
    std::mt19937 rng(1);
    std::shuffle(perm.begin(), perm.end(), rng);
    std::vector<uint32_t> h(n, 0);
    for (size_t k = 0; k < lines; ++k) h[(size_t)perm[k] * 32] = perm[(k + 1) % lines] * 32;
    uint32_t *d; cudaMalloc(&d, bytes);
    cudaMemcpy(d, h.data(), bytes, cudaMemcpyHostToDevice);
    for (int mode = 0; mode < 2; ++mode) {
      int hops = 4096;
      chase<<<1, 1>>>(d + 0, hops, mode, out, sink);
      chase<<<1, 1>>>(d + 0, hops, mode, out, sink);
      long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
      printf("footprint %8zu KB %s: %.1f cycles per dependent load\n", bytes >> 10, mode ? "ld.cg (L2)" : "ld.nc (L1)", (double)c / hops);
    }
    cudaFree(d);
  }
  redux_lat<<<1, 32>>>(out, sink);
  long long c[3]; cudaMemcpy(c, out, 24, cudaMemcpyDeviceToHost);
  printf("redux.sync.max + add: %.1f cycles; ballot + add: %.1f; shfl + add: %.1f\n", c[0] / 1024.0, c[1] / 1024.0, c[2] / 1024.0);
  return 0;
}
