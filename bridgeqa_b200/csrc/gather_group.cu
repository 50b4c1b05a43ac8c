// gather_group.cu -- index gathers and their scatter-add gradients for sm_100a.
//
// Replaces /root/reference/lib/pointnet2/_ext_src/src/sampling_gpu.cu:8-57
// (gather_points[_grad]) and src/group_points_gpu.cu:8-75 (group_points[_grad]).
// The reference launches grid = B (group) or (B, C) (gather) -- 16 CTAs on a 148-SM
// part; here the (row, channel, scene) space is tiled so the grid covers the chip,
// the index row is read once per CTA and reused for kChan channels, and all index /
// output traffic is coalesced (only the gathered reads are scattered, and they hit L2:
// one channel row of a 40k-point scene is 160 KB).
//
// These are the un-fused parity ops (bit-exact copies).  The inference path uses the
// fused SA kernel instead and never materialises the grouped tensor.
#include "common.cuh"

namespace bqa {
namespace {

constexpr int kThreads = 256;
constexpr int kChan = 8;  // channels handled per CTA for one slab of indices

// out[(s*c + l)*e_total + e] = points[(s*c + l)*n + idx[s*e_total + e]]
// covers gather (e_total = m) and group (e_total = npoints*nsample).
__global__ void __launch_bounds__(kThreads)
gather_rows_kernel(int c, int n, int e_total, const float *__restrict__ points,
                   const int *__restrict__ idx, float *__restrict__ out) {
  const int scene = blockIdx.z;
  const int l0 = blockIdx.y * kChan;
  const int e = blockIdx.x * kThreads + threadIdx.x;
  if (e >= e_total) return;
  const int a = idx[(size_t)scene * e_total + e];
  const int lend = min(l0 + kChan, c);
  for (int l = l0; l < lend; ++l) {
    const size_t row = (size_t)scene * c + l;
    out[row * e_total + e] = __ldg(points + row * n + a);
  }
}

// grad_points[(s*c + l)*n + idx[s*e_total + e]] += grad_out[(s*c + l)*e_total + e]
__global__ void __launch_bounds__(kThreads)
scatter_add_rows_kernel(int c, int n, int e_total, const float *__restrict__ grad_out,
                        const int *__restrict__ idx, float *__restrict__ grad_points) {
  const int scene = blockIdx.z;
  const int l0 = blockIdx.y * kChan;
  const int e = blockIdx.x * kThreads + threadIdx.x;
  if (e >= e_total) return;
  const int a = idx[(size_t)scene * e_total + e];
  const int lend = min(l0 + kChan, c);
  for (int l = l0; l < lend; ++l) {
    const size_t row = (size_t)scene * c + l;
    atomicAdd(grad_points + row * n + a, grad_out[row * e_total + e]);
  }
}

// Same scatter-add through a POINT-MAJOR accumulator acc (b, n, c4) (c4 = c rounded up to 4):
// four channels of one point are contiguous there, so one vector reduction
// (red.global.add.v4.f32, sm_90+) replaces four scalar atomics -- the scatter is bound by the
// number of L2 reduction operations (each source point receives ~nsample/2 adds per channel).
// A transpose then writes grad_points (b, c, n); it needs no zero-fill.
__global__ void __launch_bounds__(kThreads)
scatter_add_pm_kernel(int c, int c4, int n, int e_total, const float *__restrict__ grad_out,
                      const int *__restrict__ idx, float *__restrict__ acc) {
  const int scene = blockIdx.z;
  const int l0 = blockIdx.y * 4;
  const int e = blockIdx.x * kThreads + threadIdx.x;
  if (e >= e_total) return;
  const int a = idx[(size_t)scene * e_total + e];
  float v[4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
    v[u] = (l0 + u < c) ? grad_out[((size_t)scene * c + l0 + u) * e_total + e] : 0.f;
  float *dst = acc + ((size_t)scene * n + a) * c4 + l0;
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3])
               : "memory");
}

// acc (b, n, c4) -> out (b, c, n), dropping the padding channels
__global__ void __launch_bounds__(256)
transpose_nc_kernel(int c, int c4, int n, const float *__restrict__ acc, float *__restrict__ out) {
  __shared__ float tile[32][33];
  const int scene = blockIdx.z;
  const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int nn = n0 + r, cc = c0 + tx;
    tile[r][tx] = (nn < n && cc < c4) ? acc[((size_t)scene * n + nn) * c4 + cc] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int cc = c0 + r, nn = n0 + tx;
    if (cc < c && nn < n) out[((size_t)scene * c + cc) * n + nn] = tile[tx][r];
  }
}

// (b,c,n) -> (b,n,c) through a 32x32 shared tile (both sides coalesced)
__global__ void __launch_bounds__(256)
transpose_cn_kernel(int c, int n, const float *__restrict__ in, float *__restrict__ out) {
  __shared__ float tile[32][33];
  const int scene = blockIdx.z;
  const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const float *src = in + (size_t)scene * c * n;
  float *dst = out + (size_t)scene * c * n;
  for (int r = ty; r < 32; r += 8) {
    const int cc = c0 + r, nn = n0 + tx;
    tile[r][tx] = (cc < c && nn < n) ? src[(size_t)cc * n + nn] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int nn = n0 + r, cc = c0 + tx;
    if (nn < n && cc < c) dst[(size_t)nn * c + cc] = tile[tx][r];
  }
}

int check_grid(long long e_total, int c, int b) {
  if (ceil_div_ll(e_total, kThreads) > 2147483647ll || ceil_div(c, kChan) > 65535 || b > 65535)
    return set_error(BQA_ERR_UNSUPPORTED, "gather/group: grid too large (e=%lld c=%d b=%d)", e_total, c, b);
  return BQA_OK;
}

// QueryAndGroup.forward in one pass (pointnet2_utils.py:347-359: two group_points launches,
// subtract, divide, cat) for a POINT-MAJOR feature source -- the (B, N, 3+C) input cloud itself at
// SA1.  out (B, 3+C, E) channel-major, E = npoint * nsample:
//   out[b, 0:3, e] = (xyz[b, idx[e]] - new_xyz[b, e / nsample]) [/ radius]
//   out[b, 3+k, e] = feat[b, idx[e], k]
// A CTA owns 32 consecutive positions e: each warp reads whole feature rows (coalesced, C
// contiguous floats) into a padded shared tile, then the tile is written back channel by channel
// as 128-byte rows.  The channel-major gather it replaces reads 4 useful bytes per 32-byte sector.
__global__ void __launch_bounds__(256)
group_concat_pm_kernel(int n, int c, int feat_stride, long long e_total, int nsample, float radius,
                       int normalize, const float *__restrict__ xyz, const float *__restrict__ new_xyz,
                       const float *__restrict__ feat, const int *__restrict__ idx, float *__restrict__ out) {
  extern __shared__ float tile[];                  // [c + 3][33]
  __shared__ int s_idx[32];
  const int scene = blockIdx.y;
  const long long e0 = (long long)blockIdx.x * 32;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int npos = (int)min(32ll, e_total - e0);
  if (threadIdx.x < 32) s_idx[threadIdx.x] = threadIdx.x < npos ? idx[(size_t)scene * e_total + e0 + threadIdx.x] : 0;
  __syncthreads();
  const long long npoint = e_total / nsample;
  for (int pos = wid; pos < npos; pos += 8) {
    const int k = s_idx[pos];
    const float *row = feat + ((size_t)scene * n + k) * feat_stride;
    for (int ch = lane; ch < c; ch += 32) tile[(3 + ch) * 33 + pos] = __ldg(row + ch);
    if (lane < 3) {
      const long long j = (e0 + pos) / nsample;
      float v = xyz[((size_t)scene * n + k) * 3 + lane] - new_xyz[((size_t)scene * npoint + j) * 3 + lane];
      // pointnet2_utils.py:351-352 `grouped_xyz /= radius`: on CUDA tensors torch evaluates a
      // division by a Python scalar as a multiplication by fl(1.0f / fl(radius))
      // (ATen BinaryDivTrueKernel.cu), which is what the reference therefore computes on the GPU
      if (normalize) v = __fmul_rn(v, __fdiv_rn(1.0f, radius));
      tile[lane * 33 + pos] = v;
    }
  }
  __syncthreads();
  for (int ch = wid; ch < c + 3; ch += 8)
    if (lane < npos) out[((size_t)scene * (c + 3) + ch) * e_total + e0 + lane] = tile[ch * 33 + lane];
}

}  // namespace

int gather_rows_dispatch(int b, int c, int n, long long e_total, const float *points,
                         const int *idx, float *out, cudaStream_t stream) {
  if (b == 0 || c == 0 || e_total == 0) return BQA_OK;
  if (int rc = check_grid(e_total, c, b)) return rc;
  dim3 grid((unsigned)ceil_div_ll(e_total, kThreads), (unsigned)ceil_div(c, kChan), (unsigned)b);
  gather_rows_kernel<<<grid, kThreads, 0, stream>>>(c, n, (int)e_total, points, idx, out);
  count_launch();
  return check_launch("gather_rows_kernel");
}

int scatter_add_rows_dispatch(int b, int c, int n, long long e_total, const float *grad_out,
                              const int *idx, float *grad_points, cudaStream_t stream) {
  // torch::zeros in the reference wrappers (sampling.cpp:51-53, group_points.cpp:49-51)
  BQA_CUDA(cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)b * c * n, stream));
  if (b == 0 || c == 0 || e_total == 0) return BQA_OK;
  if (int rc = check_grid(e_total, c, b)) return rc;
  dim3 grid((unsigned)ceil_div_ll(e_total, kThreads), (unsigned)ceil_div(c, kChan), (unsigned)b);
  scatter_add_rows_kernel<<<grid, kThreads, 0, stream>>>(c, n, (int)e_total, grad_out, idx, grad_points);
  count_launch();
  return check_launch("scatter_add_rows_kernel");
}

long long scatter_add_workspace_bytes(int b, int c, int n) {
  return 4ll * b * n * ((c + 3) / 4 * 4);
}

// scatter_add_rows_dispatch with a caller-provided accumulator of scatter_add_workspace_bytes()
int scatter_add_rows_ws_dispatch(int b, int c, int n, long long e_total, const float *grad_out, const int *idx,
                                 float *grad_points, float *acc, cudaStream_t stream) {
  if (b == 0 || c == 0 || n == 0) return BQA_OK;
  const int c4 = (c + 3) / 4 * 4;
  BQA_CUDA(cudaMemsetAsync(acc, 0, sizeof(float) * (size_t)b * n * c4, stream));
  if (e_total > 0) {
    if (ceil_div_ll(e_total, kThreads) > 2147483647ll || c4 / 4 > 65535 || b > 65535)
      return set_error(BQA_ERR_UNSUPPORTED, "group grad: grid too large (e=%lld c=%d b=%d)", e_total, c, b);
    dim3 grid((unsigned)ceil_div_ll(e_total, kThreads), (unsigned)(c4 / 4), (unsigned)b);
    scatter_add_pm_kernel<<<grid, kThreads, 0, stream>>>(c, c4, n, (int)e_total, grad_out, idx, acc);
    count_launch();
    if (int rc = check_launch("scatter_add_pm_kernel")) return rc;
  }
  dim3 tg((unsigned)ceil_div(n, 32), (unsigned)ceil_div(c, 32), (unsigned)b);
  transpose_nc_kernel<<<tg, 256, 0, stream>>>(c, c4, n, acc, grad_points);
  count_launch();
  return check_launch("transpose_nc_kernel");
}

int transpose_cn_dispatch(int b, int c, int n, const float *in, float *out, cudaStream_t stream) {
  if (b == 0 || c == 0 || n == 0) return BQA_OK;
  dim3 grid((unsigned)ceil_div(n, 32), (unsigned)ceil_div(c, 32), (unsigned)b);
  transpose_cn_kernel<<<grid, 256, 0, stream>>>(c, n, in, out);
  count_launch();
  return check_launch("transpose_cn_kernel");
}

}  // namespace bqa

namespace bqa {
int group_concat_pm_dispatch(int b, int n, int c, int feat_stride, long long e_total, int nsample, float radius,
                             int normalize, const float *xyz, const float *new_xyz, const float *feat,
                             const int *idx, float *out, cudaStream_t stream) {
  const size_t smem = sizeof(float) * 33 * (size_t)(c + 3);
  if (smem > 200 * 1024 || b > 65535) return set_error(BQA_ERR_UNSUPPORTED, "group_concat: c=%d too wide", c);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    BQA_CUDA(cudaFuncSetAttribute(group_concat_pm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  dim3 grid((unsigned)((e_total + 31) / 32), (unsigned)b);
  group_concat_pm_kernel<<<grid, 256, smem, stream>>>(n, c, feat_stride, e_total, nsample, radius, normalize, xyz,
                                                      new_xyz, feat, idx, out);
  count_launch();
  return check_launch("group_concat_pm_kernel");
}
}  // namespace bqa
