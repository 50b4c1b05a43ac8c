// interpolate.cu -- three_nn / three_interpolate (+grad) for sm_100a.
//
// Replaces /root/reference/lib/pointnet2/_ext_src/src/interpolate_gpu.cu:9-154
// (all three launched with grid = B).  Here: one thread per unknown point over
// shared-memory tiles of the known set (three_nn); one thread per output column with a
// channel slab per CTA (three_interpolate), so index/weight rows are read once per slab.
//
// Bit-exact: d = fma(dz,dz,fma(dx,dx,dy*dy)); strict `<` cascade in ascending k
// (lowest index wins ties); fewer than 3 known points leave dist2 = +inf (the
// reference's (float)1e40) and idx = 0.  Interpolation is
// fma(p3,w3,fma(p1,w1,p2*w2)), the contraction nvcc emits for interpolate_gpu.cu:99-100.
#include "common.cuh"

namespace bqa {
namespace {

constexpr int kNNThreads = 128;
constexpr int kKnownTile = 1024;

__global__ void __launch_bounds__(kNNThreads)
three_nn_kernel(int n, int m, const float *__restrict__ unknown_all,
                const float *__restrict__ known_all, float *__restrict__ dist2_all,
                int *__restrict__ idx_all) {
  __shared__ float tile[kKnownTile * 3];
  const int scene = blockIdx.y;
  const int j = blockIdx.x * kNNThreads + threadIdx.x;
  const float *known = known_all + (size_t)scene * m * 3;
  const bool live = j < n;
  float ux = 0.f, uy = 0.f, uz = 0.f;
  if (live) {
    const float *u = unknown_all + ((size_t)scene * n + j) * 3;
    ux = u[0]; uy = u[1]; uz = u[2];
  }
  // the reference keeps doubles initialised to 1e40; every d is a float, so comparing
  // in float against +inf takes the same branches, and (float)1e40 == +inf on output.
  float best1 = INFINITY, best2 = INFINITY, best3 = INFINITY;
  int i1 = 0, i2 = 0, i3 = 0;
  for (int base = 0; base < m; base += kKnownTile) {
    const int tn = min(kKnownTile, m - base);
    __syncthreads();
    for (int t = threadIdx.x; t < tn * 3; t += kNNThreads) tile[t] = known[(size_t)base * 3 + t];
    __syncthreads();
    if (live) {
#pragma unroll 4
      for (int t = 0; t < tn; ++t) {
        const float d = sqdist3(ux, uy, uz, tile[t * 3], tile[t * 3 + 1], tile[t * 3 + 2]);
        const int k = base + t;
        if (d < best1) {
          best3 = best2; i3 = i2; best2 = best1; i2 = i1; best1 = d; i1 = k;
        } else if (d < best2) {
          best3 = best2; i3 = i2; best2 = d; i2 = k;
        } else if (d < best3) {
          best3 = d; i3 = k;
        }
      }
    }
  }
  if (live) {
    float *o = dist2_all + ((size_t)scene * n + j) * 3;
    int *oi = idx_all + ((size_t)scene * n + j) * 3;
    o[0] = best1; o[1] = best2; o[2] = best3;
    oi[0] = i1; oi[1] = i2; oi[2] = i3;
  }
}

constexpr int kIThreads = 128;
constexpr int kIChan = 16;

__global__ void __launch_bounds__(kIThreads)
three_interpolate_kernel(int c, int m, int n, const float *__restrict__ points,
                         const int *__restrict__ idx, const float *__restrict__ weight,
                         float *__restrict__ out) {
  const int scene = blockIdx.z;
  const int j = blockIdx.x * kIThreads + threadIdx.x;
  if (j >= n) return;
  const size_t q = ((size_t)scene * n + j) * 3;
  const int a1 = idx[q], a2 = idx[q + 1], a3 = idx[q + 2];
  const float w1 = weight[q], w2 = weight[q + 1], w3 = weight[q + 2];
  const int l0 = blockIdx.y * kIChan, lend = min(l0 + kIChan, c);
  for (int l = l0; l < lend; ++l) {
    const float *p = points + ((size_t)scene * c + l) * m;
    out[((size_t)scene * c + l) * n + j] =
        __fmaf_rn(__ldg(p + a3), w3, __fmaf_rn(__ldg(p + a1), w1, __fmul_rn(__ldg(p + a2), w2)));
  }
}

__global__ void __launch_bounds__(kIThreads)
three_interpolate_grad_kernel(int c, int n, int m, const float *__restrict__ grad_out,
                              const int *__restrict__ idx, const float *__restrict__ weight,
                              float *__restrict__ grad_points) {
  const int scene = blockIdx.z;
  const int j = blockIdx.x * kIThreads + threadIdx.x;
  if (j >= n) return;
  const size_t q = ((size_t)scene * n + j) * 3;
  const int a1 = idx[q], a2 = idx[q + 1], a3 = idx[q + 2];
  const float w1 = weight[q], w2 = weight[q + 1], w3 = weight[q + 2];
  const int l0 = blockIdx.y * kIChan, lend = min(l0 + kIChan, c);
  for (int l = l0; l < lend; ++l) {
    const float g = grad_out[((size_t)scene * c + l) * n + j];
    float *gp = grad_points + ((size_t)scene * c + l) * m;
    atomicAdd(gp + a1, __fmul_rn(g, w1));
    atomicAdd(gp + a2, __fmul_rn(g, w2));
    atomicAdd(gp + a3, __fmul_rn(g, w3));
  }
}

}  // namespace

int three_nn_dispatch(int b, int n, int m, const float *unknown, const float *known, float *dist2,
                      int *idx, cudaStream_t stream) {
  if (b == 0 || n == 0) return BQA_OK;
  dim3 grid((unsigned)ceil_div(n, kNNThreads), (unsigned)b);
  three_nn_kernel<<<grid, kNNThreads, 0, stream>>>(n, m, unknown, known, dist2, idx);
  count_launch();
  return check_launch("three_nn_kernel");
}

int three_interpolate_dispatch(int b, int c, int m, int n, const float *points, const int *idx,
                               const float *weight, float *out, cudaStream_t stream) {
  if (b == 0 || c == 0 || n == 0) return BQA_OK;
  dim3 grid((unsigned)ceil_div(n, kIThreads), (unsigned)ceil_div(c, kIChan), (unsigned)b);
  three_interpolate_kernel<<<grid, kIThreads, 0, stream>>>(c, m, n, points, idx, weight, out);
  count_launch();
  return check_launch("three_interpolate_kernel");
}

int three_interpolate_grad_dispatch(int b, int c, int n, int m, const float *grad_out,
                                    const int *idx, const float *weight, float *grad_points,
                                    cudaStream_t stream) {
  BQA_CUDA(cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)b * c * m, stream));
  if (b == 0 || c == 0 || n == 0) return BQA_OK;
  dim3 grid((unsigned)ceil_div(n, kIThreads), (unsigned)ceil_div(c, kIChan), (unsigned)b);
  three_interpolate_grad_kernel<<<grid, kIThreads, 0, stream>>>(c, n, m, grad_out, idx, weight,
                                                                grad_points);
  count_launch();
  return check_launch("three_interpolate_grad_kernel");
}

}  // namespace bqa
