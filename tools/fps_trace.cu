// fps_trace.cu -- developer tool (not part of the product library): phase-by-phase cycle
// breakdown of the cluster FPS kernel.  Includes fps.cu with BQA_FPS_TRACE so the kernel
// accumulates clock64() deltas at its phase boundaries for (cluster 0, CTA 0, thread 0).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/fps_trace tools/fps_trace.cu
#define BQA_FPS_TRACE 1
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstdarg>
#include <cmath>
#include "../bridgeqa_b200/csrc/common.cuh"
namespace bqa {
int set_error(int code, const char *fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fprintf(stderr, "\n"); return code; }
void count_launch(int) {}
int check_launch(const char *what) { cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", what, cudaGetErrorString(e)); return 2; } return 0; }
int ref_opt_n_threads(int w) { int p = (int)(std::log((double)w) / std::log(2.0)); int v = 1 << p; return v > 512 ? 512 : (v < 1 ? 1 : v); }
}
#include "../bridgeqa_b200/csrc/fps.cu"

int main(int argc, char **argv) {
  int b = argc > 1 ? atoi(argv[1]) : 16, n = argc > 2 ? atoi(argv[2]) : 40000, m = argc > 3 ? atoi(argv[3]) : 2048;
  std::vector<float> h((size_t)b * n * 3);
  srand(1);
  for (auto &v : h) v = (float)rand() / RAND_MAX * 8.f - 4.f;
  float *xyz, *nx; int *idx; unsigned long long *trace;
  cudaMalloc(&xyz, h.size() * 4); cudaMalloc(&nx, (size_t)b * m * 12); cudaMalloc(&idx, (size_t)b * m * 4);
  cudaMalloc(&trace, 64 * 8); cudaMemset(trace, 0, 64 * 8);
  cudaMemcpy(xyz, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  bqa::g_fps_trace = trace;
  for (int rep = 0; rep < 3; ++rep) {
    cudaMemset(trace, 0, 64 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    int rc = bqa::fps_dispatch(b, n, m, 1, m, xyz, idx, nx, nullptr, false, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long t[64]; cudaMemcpy(t, trace, sizeof(t), cudaMemcpyDeviceToHost);
    printf("rc=%d b=%d n=%d m=%d  %.3f ms  %.3f us/iter\n", rc, b, n, m, ms, 1e3 * ms / (m - 1));
    const char *names[] = {"compute+warp-redux", "syncthreads", "warp0 reduce+send", "mbar wait", "recv reduce", "total"};
    for (int i = 0; i < 6; ++i) printf("   %-20s %8.1f cycles/iter\n", names[i], (double)t[i] / (m - 1));
  }
  return 0;
}
