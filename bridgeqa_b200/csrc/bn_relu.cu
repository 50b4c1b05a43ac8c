// bn_relu.cu -- train-mode BatchNorm + ReLU (+ max over nsample) for the SharedMLP blocks of
// the SA / FP layers, forward and backward, for sm_100a.  HBM-bound streaming kernels.
//
// Replaces, in model.train(), what the reference runs per SharedMLP block after the 1x1 conv
// (lib/pointnet2/pytorch_utils.py:11-36, 73-80: nn.BatchNorm2d in training mode, then the shared
// nn.ReLU(inplace=True)) and, for the last block of an SA layer, the F.max_pool2d over nsample
// that follows (pointnet2_modules.py:259-262), plus their autograd backward.  On B200 cuDNN's
// bn_bw_1C11 / bn_fw_tr_1C11 kernels and torch's max-pool kernels take 25 of the 42 ms of a DET
// training step at 16 x 40000 points; these passes are pure streaming work:
//
//   forward   stats:  one read of y            -> per-channel sum, sum of squares (double)
//             apply:  one read of y, one write -> x = relu(a*y + b),  a = gamma*invstd, b = beta - mean*a
//             apply+max (last block of an SA layer): one read of y -> out[b,c,j] = max_s x, argmax;
//                       the (B,C,npoint,nsample) activation is never written
//   backward  stats:  read dx, y               -> sum dz, sum dz*xhat   (dz = dx * [x > 0])
//             apply:  read dx, y, write dy     -> dy = a * (dz - mean(dz) - xhat * mean(dz*xhat))
//             max variants: dz is dout at the argmax position (and x > 0), zero elsewhere
//
// Layout: y (B, C, L) contiguous, L = npoint*nsample (SA) or n (FP): the reference's NCHW.
// Per-channel sums are accumulated as fp32 inside a CTA (<= 16384 elements) and as double
// across CTAs (atomicAdd), so the result does not depend on the CTA order beyond 1e-16.
// BatchNorm semantics (torch.nn.functional.batch_norm, training=True): biased variance for the
// normalisation, unbiased (N/(N-1)) for running_var, running = (1-momentum)*running + momentum*batch.
#include "common.cuh"

namespace bqa {
namespace {

constexpr int kThreads = 256;
constexpr int kChunk = 16384;              // elements of one (b, c) row per CTA

struct Affine { float a, b, mean, invstd; };

// torch's ReLU propagates NaN (threshold: x <= 0 ? 0 : x); fmaxf(x, 0) would turn a NaN activation
// into 0 and hide a diverging run
__device__ __forceinline__ float relu_nan(float v) { return v > 0.f ? v : (v != v ? v : 0.f); }

__device__ __forceinline__ Affine affine_of(int c, const float *mean, const float *invstd,
                                            const float *gamma, const float *beta) {
  Affine f;
  f.mean = mean[c];
  f.invstd = invstd[c];
  f.a = gamma[c] * f.invstd;
  f.b = beta[c] - f.mean * f.a;
  return f;
}

// block-wide sum of two floats -> thread 0 adds them to two doubles
__device__ __forceinline__ void block_add2(float s1, float s2, double *d1, double *d2) {
  __shared__ float red[2][kThreads / 32];
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { red[0][wid] = s1; red[1][wid] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t1 = 0.f, t2 = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) { t1 += red[0][w]; t2 += red[1][w]; }
    atomicAdd(d1, (double)t1);
    atomicAdd(d2, (double)t2);
  }
}

// ---- forward ------------------------------------------------------------------------------

__global__ void __launch_bounds__(kThreads)
bn_stats_kernel(int c, long long l, const float *__restrict__ y, double *__restrict__ sums) {
  const int ch = blockIdx.y;
  const float *row = y + ((size_t)blockIdx.z * c + ch) * l;
  const long long i0 = (long long)blockIdx.x * kChunk;
  const long long i1 = min(l, i0 + kChunk);
  // sums are taken around a per-channel shift K (the channel's first element): E[(y-K)^2] - E[y-K]^2
  // does not cancel catastrophically when |mean| >> std, which E[y^2] - E[y]^2 in fp32 does
  const float K = __ldg(y + (size_t)ch * l);
  float s1 = 0.f, s2 = 0.f;
  if ((l & 3) == 0) {
    const float4 *r4 = reinterpret_cast<const float4 *>(row);
    // 4 independent 16-byte loads in flight per thread, 4 accumulator pairs
    float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
    long long i = i0 / 4 + threadIdx.x;
    const long long e4 = i1 / 4;
    for (; i + 3 * kThreads < e4; i += 4 * kThreads) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(r4 + i + u * kThreads);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        v[u].x -= K; v[u].y -= K; v[u].z -= K; v[u].w -= K;
        a1[u] += (v[u].x + v[u].y) + (v[u].z + v[u].w);
        a2[u] += (v[u].x * v[u].x + v[u].y * v[u].y) + (v[u].z * v[u].z + v[u].w * v[u].w);
      }
    }
    for (; i < e4; i += kThreads) {
      float4 v = __ldg(r4 + i);
      v.x -= K; v.y -= K; v.z -= K; v.w -= K;
      a1[0] += (v.x + v.y) + (v.z + v.w);
      a2[0] += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
    s1 = (a1[0] + a1[1]) + (a1[2] + a1[3]);
    s2 = (a2[0] + a2[1]) + (a2[2] + a2[3]);
  } else {
    for (long long i = i0 + threadIdx.x; i < i1; i += kThreads) {
      const float v = __ldg(row + i) - K;
      s1 += v;
      s2 += v * v;
    }
  }
  block_add2(s1, s2, &sums[ch], &sums[c + ch]);
}

__global__ void bn_finalize_kernel(int c, long long l, double count, const float *__restrict__ y,
                                   const double *__restrict__ sums, float eps,
                                   float momentum, float *__restrict__ mean, float *__restrict__ invstd,
                                   float *__restrict__ running_mean, float *__restrict__ running_var) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  const double ms = sums[ch] / count;                 // mean of (y - K), K as in bn_stats_kernel
  double var = sums[c + ch] / count - ms * ms;
  if (var < 0.0) var = 0.0;
  const double m = (double)y[(size_t)ch * l] + ms;
  mean[ch] = (float)m;
  invstd[ch] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)m;
  if (running_var) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unbiased;
  }
}

// sums = [sum (y - K), sum (y - K)^2] accumulated elsewhere (the conv kernel's epilogue, conv_tf32.cu)
// around the per-channel shift K = shift[ch] (NULL: 0); shift may alias running_mean (read first)
__global__ void bn_finalize_shifted_kernel(int c, double count, const double *__restrict__ sums,
                                           const float *shift, float eps, float momentum,
                                           float *__restrict__ mean, float *__restrict__ invstd,
                                           float *running_mean, float *running_var) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  const double k = shift ? (double)shift[ch] : 0.0;
  const double ms = sums[ch] / count;
  double var = sums[c + ch] / count - ms * ms;
  if (var < 0.0) var = 0.0;
  const double m = k + ms;
  mean[ch] = (float)m;
  invstd[ch] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)m;
  if (running_var) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unbiased;
  }
}

__global__ void __launch_bounds__(kThreads)
bn_relu_apply_kernel(int c, long long l, const float *__restrict__ y, const float *__restrict__ mean,
                     const float *__restrict__ invstd, const float *__restrict__ gamma,
                     const float *__restrict__ beta, float *__restrict__ x) {
  const int ch = blockIdx.y;
  const Affine f = affine_of(ch, mean, invstd, gamma, beta);
  const size_t off = ((size_t)blockIdx.z * c + ch) * l;
  const long long i0 = (long long)blockIdx.x * kChunk;
  const long long i1 = min(l, i0 + kChunk);
  if ((l & 3) == 0) {
    const float4 *r4 = reinterpret_cast<const float4 *>(y + off);
    float4 *o4 = reinterpret_cast<float4 *>(x + off);
    for (long long i = i0 / 4 + threadIdx.x; i < i1 / 4; i += kThreads) {
      const float4 v = __ldg(r4 + i);
      o4[i] = make_float4(relu_nan(fmaf(v.x, f.a, f.b)), relu_nan(fmaf(v.y, f.a, f.b)),
                          relu_nan(fmaf(v.z, f.a, f.b)), relu_nan(fmaf(v.w, f.a, f.b)));
    }
  } else {
    for (long long i = i0 + threadIdx.x; i < i1; i += kThreads)
      x[off + i] = relu_nan(fmaf(__ldg(y + off + i), f.a, f.b));
  }
}

// One (b, c) row per blockIdx.y; a thread handles kMaxU float4s (4 consecutive neighbours each),
// ns / 4 adjacent lanes hold one (b, c, centre) group.  row4 = npoint * ns / 4 float4s per row.
constexpr int kMaxU = 4;

__global__ void __launch_bounds__(kThreads)
bn_relu_max_kernel(int c, int row4, int lpg_shift, const float *__restrict__ y,
                   const float *__restrict__ mean, const float *__restrict__ invstd,
                   const float *__restrict__ gamma, const float *__restrict__ beta,
                   float *__restrict__ out, int *__restrict__ argmax) {
  const int ch = blockIdx.y % c;
  const Affine f = affine_of(ch, mean, invstd, gamma, beta);
  const size_t row = blockIdx.y;
  const float4 *r4 = reinterpret_cast<const float4 *>(y) + row * row4;
  const int lpg = 1 << lpg_shift;
  const int e0 = blockIdx.x * (kThreads * kMaxU) + threadIdx.x;
  float4 v[kMaxU];
#pragma unroll
  for (int u = 0; u < kMaxU; ++u) {
    const int e = e0 + u * kThreads;
    v[u] = e < row4 ? __ldg(r4 + e) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int u = 0; u < kMaxU; ++u) {
    const int e = e0 + u * kThreads;
    const int s0 = (e & (lpg - 1)) * 4;
    const float z0 = fmaf(v[u].x, f.a, f.b), z1 = fmaf(v[u].y, f.a, f.b), z2 = fmaf(v[u].z, f.a, f.b),
                z3 = fmaf(v[u].w, f.a, f.b);
    float best = z0;
    int bi = s0;
    if (z1 > best) { best = z1; bi = s0 + 1; }
    if (z2 > best) { best = z2; bi = s0 + 2; }
    if (z3 > best) { best = z3; bi = s0 + 3; }
    for (int o = 1; o < lpg; o <<= 1) {            // groups are lpg-aligned inside the warp
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (e < row4 && (e & (lpg - 1)) == 0) {
      const size_t g = row * (size_t)(row4 >> lpg_shift) + (size_t)(e >> lpg_shift);
      out[g] = relu_nan(best);
      argmax[g] = bi;
    }
  }
}

// ---- backward -----------------------------------------------------------------------------

__global__ void __launch_bounds__(kThreads)
bn_relu_bwd_stats_kernel(int c, long long l, const float *__restrict__ dx, const float *__restrict__ y,
                         const float *__restrict__ mean, const float *__restrict__ invstd,
                         const float *__restrict__ gamma, const float *__restrict__ beta,
                         double *__restrict__ sums) {
  const int ch = blockIdx.y;
  const Affine f = affine_of(ch, mean, invstd, gamma, beta);
  const size_t off = ((size_t)blockIdx.z * c + ch) * l;
  const long long i0 = (long long)blockIdx.x * kChunk;
  const long long i1 = min(l, i0 + kChunk);
  float s1 = 0.f, s2 = 0.f;
  auto acc = [&](float yv, float g) {
    const float dz = fmaf(yv, f.a, f.b) > 0.f ? g : 0.f;
    s1 += dz;
    s2 += dz * ((yv - f.mean) * f.invstd);
  };
  if ((l & 3) == 0) {
    const float4 *y4 = reinterpret_cast<const float4 *>(y + off);
    const float4 *g4 = reinterpret_cast<const float4 *>(dx + off);
    for (long long i = i0 / 4 + threadIdx.x; i < i1 / 4; i += kThreads) {
      const float4 v = __ldg(y4 + i), g = __ldg(g4 + i);
      acc(v.x, g.x); acc(v.y, g.y); acc(v.z, g.z); acc(v.w, g.w);
    }
  } else {
    for (long long i = i0 + threadIdx.x; i < i1; i += kThreads) acc(__ldg(y + off + i), __ldg(dx + off + i));
  }
  block_add2(s1, s2, &sums[ch], &sums[c + ch]);
}

__global__ void __launch_bounds__(kThreads)
bn_relu_bwd_apply_kernel(int c, long long l, double count, const float *__restrict__ dx,
                         const float *__restrict__ y, const float *__restrict__ mean,
                         const float *__restrict__ invstd, const float *__restrict__ gamma,
                         const float *__restrict__ beta, const double *__restrict__ sums,
                         float *__restrict__ dy) {
  const int ch = blockIdx.y;
  const Affine f = affine_of(ch, mean, invstd, gamma, beta);
  const float m1 = (float)(sums[ch] / count), m2 = (float)(sums[c + ch] / count);
  const size_t off = ((size_t)blockIdx.z * c + ch) * l;
  const long long i0 = (long long)blockIdx.x * kChunk;
  const long long i1 = min(l, i0 + kChunk);
  auto grad = [&](float yv, float g) {
    const float dz = fmaf(yv, f.a, f.b) > 0.f ? g : 0.f;
    return f.a * (dz - m1 - ((yv - f.mean) * f.invstd) * m2);
  };
  if ((l & 3) == 0) {
    const float4 *y4 = reinterpret_cast<const float4 *>(y + off);
    const float4 *g4 = reinterpret_cast<const float4 *>(dx + off);
    float4 *o4 = reinterpret_cast<float4 *>(dy + off);
    for (long long i = i0 / 4 + threadIdx.x; i < i1 / 4; i += kThreads) {
      const float4 v = __ldg(y4 + i), g = __ldg(g4 + i);
      o4[i] = make_float4(grad(v.x, g.x), grad(v.y, g.y), grad(v.z, g.z), grad(v.w, g.w));
    }
  } else {
    for (long long i = i0 + threadIdx.x; i < i1; i += kThreads)
      dy[off + i] = grad(__ldg(y + off + i), __ldg(dx + off + i));
  }
}

// max variants: the upstream gradient lives on (B, C, npoint); it reaches position argmax only
__global__ void __launch_bounds__(kThreads)
bn_relu_max_bwd_stats_kernel(int c, long long np, int ns, const float *__restrict__ dout,
                             const int *__restrict__ argmax, const float *__restrict__ y,
                             const float *__restrict__ mean, const float *__restrict__ invstd,
                             const float *__restrict__ gamma, const float *__restrict__ beta,
                             double *__restrict__ sums) {
  const int ch = blockIdx.y;
  const Affine f = affine_of(ch, mean, invstd, gamma, beta);
  const size_t row = (size_t)blockIdx.z * c + ch;
  const long long j = (long long)blockIdx.x * kThreads + threadIdx.x;
  float s1 = 0.f, s2 = 0.f;
  if (j < np) {
    const int p = argmax[row * np + j];
    const float yv = __ldg(y + (row * np + j) * ns + p);
    const float dz = fmaf(yv, f.a, f.b) > 0.f ? dout[row * np + j] : 0.f;
    s1 = dz;
    s2 = dz * ((yv - f.mean) * f.invstd);
  }
  block_add2(s1, s2, &sums[ch], &sums[c + ch]);
}

__global__ void __launch_bounds__(kThreads)
bn_relu_max_bwd_apply_kernel(int c, int row4, int lpg_shift, double count,
                             const float *__restrict__ dout, const int *__restrict__ argmax,
                             const float *__restrict__ y, const float *__restrict__ mean,
                             const float *__restrict__ invstd, const float *__restrict__ gamma,
                             const float *__restrict__ beta, const double *__restrict__ sums,
                             float *__restrict__ dy) {
  const int ch = blockIdx.y % c;
  const Affine f = affine_of(ch, mean, invstd, gamma, beta);
  const float m1 = (float)(sums[ch] / count), m2 = (float)(sums[c + ch] / count);
  const size_t row = blockIdx.y;
  const float4 *r4 = reinterpret_cast<const float4 *>(y) + row * row4;
  float4 *o4 = reinterpret_cast<float4 *>(dy) + row * row4;
  const size_t g0 = row * (size_t)(row4 >> lpg_shift);
  const int lpg = 1 << lpg_shift;
  const int e0 = blockIdx.x * (kThreads * kMaxU) + threadIdx.x;
  float4 v[kMaxU];
  int p[kMaxU];
  float g[kMaxU];
#pragma unroll
  for (int u = 0; u < kMaxU; ++u) {
    const int e = e0 + u * kThreads;
    const bool live = e < row4;
    v[u] = live ? __ldg(r4 + e) : make_float4(0.f, 0.f, 0.f, 0.f);
    p[u] = live ? __ldg(argmax + g0 + (e >> lpg_shift)) : -1;
    g[u] = live ? __ldg(dout + g0 + (e >> lpg_shift)) : 0.f;
  }
#pragma unroll
  for (int u = 0; u < kMaxU; ++u) {
    const int e = e0 + u * kThreads;
    if (e >= row4) continue;
    const int s0 = (e & (lpg - 1)) * 4;
    auto grad = [&](float yv, int s) {
      const float dz = (s == p[u] && fmaf(yv, f.a, f.b) > 0.f) ? g[u] : 0.f;
      return f.a * (dz - m1 - ((yv - f.mean) * f.invstd) * m2);
    };
    o4[e] = make_float4(grad(v[u].x, s0), grad(v[u].y, s0 + 1), grad(v[u].z, s0 + 2), grad(v[u].w, s0 + 3));
  }
}

__global__ void bn_param_grad_kernel(int c, const double *__restrict__ sums, float *__restrict__ dgamma,
                                     float *__restrict__ dbeta) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  dbeta[ch] = (float)sums[ch];
  dgamma[ch] = (float)sums[c + ch];
}

dim3 row_grid(int b, int c, long long l) {
  return dim3((unsigned)((l + kChunk - 1) / kChunk), (unsigned)c, (unsigned)b);
}

}  // namespace

bool bn_relu_max_supported(int ns) { return ns >= 4 && ns <= 128 && (ns & (ns - 1)) == 0; }

int bn_stats_dispatch(int b, int c, long long l, const float *y, double *sums, float eps, float momentum,
                      float *mean, float *invstd, float *running_mean, float *running_var,
                      cudaStream_t stream) {
  if (b > 65535 || c > 65535) return set_error(BQA_ERR_UNSUPPORTED, "bn: b or c > 65535");
  BQA_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)c, stream));
  bn_stats_kernel<<<row_grid(b, c, l), kThreads, 0, stream>>>(c, l, y, sums);
  count_launch();
  if (int rc = check_launch("bn_stats_kernel")) return rc;
  bn_finalize_kernel<<<ceil_div(c, 128), 128, 0, stream>>>(c, l, (double)b * (double)l, y, sums, eps, momentum,
                                                          mean, invstd, running_mean, running_var);
  count_launch();
  return check_launch("bn_finalize_kernel");
}

int bn_finalize_shifted_dispatch(int c, double count, const double *sums, const float *shift, float eps,
                                 float momentum, float *mean, float *invstd, float *running_mean,
                                 float *running_var, cudaStream_t stream) {
  bn_finalize_shifted_kernel<<<ceil_div(c, 128), 128, 0, stream>>>(c, count, sums, shift, eps, momentum, mean,
                                                                  invstd, running_mean, running_var);
  count_launch();
  return check_launch("bn_finalize_shifted_kernel");
}

int bn_relu_apply_dispatch(int b, int c, long long l, const float *y, const float *mean, const float *invstd,
                           const float *gamma, const float *beta, float *x, cudaStream_t stream) {
  if (b > 65535 || c > 65535) return set_error(BQA_ERR_UNSUPPORTED, "bn: b or c > 65535");
  bn_relu_apply_kernel<<<row_grid(b, c, l), kThreads, 0, stream>>>(c, l, y, mean, invstd, gamma, beta, x);
  count_launch();
  return check_launch("bn_relu_apply_kernel");
}

int bn_relu_max_dispatch(int b, int c, long long np, int ns, const float *y, const float *mean,
                         const float *invstd, const float *gamma, const float *beta, float *out, int *argmax,
                         cudaStream_t stream) {
  const long long row4 = np * (ns / 4);
  if ((long long)b * c > 65535 || row4 > 0x7fffffffll)
    return set_error(BQA_ERR_UNSUPPORTED, "bn_relu_max: b*c=%lld rows / %lld per row not supported",
                     (long long)b * c, row4);
  int shift = 0;
  while ((4 << shift) < ns) ++shift;
  dim3 grid((unsigned)((row4 + kThreads * kMaxU - 1) / (kThreads * kMaxU)), (unsigned)(b * c));
  bn_relu_max_kernel<<<grid, kThreads, 0, stream>>>(c, (int)row4, shift, y, mean, invstd, gamma, beta, out,
                                                   argmax);
  count_launch();
  return check_launch("bn_relu_max_kernel");
}

int bn_relu_backward_dispatch(int b, int c, long long l, const float *dx, const float *y, const float *mean,
                              const float *invstd, const float *gamma, const float *beta, double *sums,
                              float *dy, float *dgamma, float *dbeta, cudaStream_t stream) {
  if (b > 65535 || c > 65535) return set_error(BQA_ERR_UNSUPPORTED, "bn: b or c > 65535");
  BQA_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)c, stream));
  bn_relu_bwd_stats_kernel<<<row_grid(b, c, l), kThreads, 0, stream>>>(c, l, dx, y, mean, invstd, gamma, beta,
                                                                      sums);
  count_launch();
  if (int rc = check_launch("bn_relu_bwd_stats_kernel")) return rc;
  bn_relu_bwd_apply_kernel<<<row_grid(b, c, l), kThreads, 0, stream>>>(c, l, (double)b * (double)l, dx, y, mean,
                                                                      invstd, gamma, beta, sums, dy);
  count_launch();
  if (int rc = check_launch("bn_relu_bwd_apply_kernel")) return rc;
  bn_param_grad_kernel<<<ceil_div(c, 128), 128, 0, stream>>>(c, sums, dgamma, dbeta);
  count_launch();
  return check_launch("bn_param_grad_kernel");
}

int bn_relu_max_backward_dispatch(int b, int c, long long np, int ns, const float *dout, const int *argmax,
                                  const float *y, const float *mean, const float *invstd, const float *gamma,
                                  const float *beta, double *sums, float *dy, float *dgamma, float *dbeta,
                                  cudaStream_t stream) {
  if (b > 65535 || c > 65535) return set_error(BQA_ERR_UNSUPPORTED, "bn: b or c > 65535");
  BQA_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)c, stream));
  dim3 g((unsigned)((np + kThreads - 1) / kThreads), (unsigned)c, (unsigned)b);
  bn_relu_max_bwd_stats_kernel<<<g, kThreads, 0, stream>>>(c, np, ns, dout, argmax, y, mean, invstd, gamma, beta,
                                                          sums);
  count_launch();
  if (int rc = check_launch("bn_relu_max_bwd_stats_kernel")) return rc;
  const long long row4 = np * (ns / 4);
  if ((long long)b * c > 65535 || row4 > 0x7fffffffll)
    return set_error(BQA_ERR_UNSUPPORTED, "bn_relu_max: b*c=%lld rows / %lld per row not supported",
                     (long long)b * c, row4);
  int shift = 0;
  while ((4 << shift) < ns) ++shift;
  dim3 ag((unsigned)((row4 + kThreads * kMaxU - 1) / (kThreads * kMaxU)), (unsigned)(b * c));
  bn_relu_max_bwd_apply_kernel<<<ag, kThreads, 0, stream>>>(
      c, (int)row4, shift, (double)b * (double)np * (double)ns, dout, argmax, y, mean, invstd, gamma, beta, sums, dy);
  count_launch();
  if (int rc = check_launch("bn_relu_max_bwd_apply_kernel")) return rc;
  bn_param_grad_kernel<<<ceil_div(c, 128), 128, 0, stream>>>(c, sums, dgamma, dbeta);
  count_launch();
  return check_launch("bn_param_grad_kernel");
}

}  // namespace bqa
