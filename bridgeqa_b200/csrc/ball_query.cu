// ball_query.cu -- radius search with the reference's "first nsample in index order"
// semantics, for sm_100a.
//
// Replaces /root/reference/lib/pointnet2/_ext_src/src/ball_query_gpu.cu:9-54
// (grid = B, each thread scans all n points from global memory for m/512 queries).
//
// Layout: one thread owns one query (its centre stays in registers); a CTA of
// kQueries queries walks the scene's xyz in tiles staged in shared memory by the
// bulk-copy engine (cp.async.bulk -> mbarrier, double buffered), so every point is read
// from L2/HBM once per CTA and then broadcast to all lanes from shared memory.
// Hits are appended in ascending k, so the row is "the first nsample hits"; the
// reference's "first hit back-fills the whole row" is applied once at the end.
//
// Bit-exact: d2 = fma(dz,dz,fma(dx,dx,dy*dy)) (nvcc's contraction of
// ball_query_gpu.cu:30-31), compared `<` against radius*radius computed in fp32.
#include "common.cuh"

namespace bqa {
namespace {

constexpr int kQueries = 128;      // threads per CTA == queries per CTA
constexpr int kTile = 1920;        // points per shared-memory tile (22.5 KB as raw xyz)
constexpr int kStages = 2;

__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes,
                                          uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

__global__ void __launch_bounds__(kQueries)
ball_query_kernel(int n, int m, float radius2, int nsample, const float *__restrict__ new_xyz_all,
                  const float *__restrict__ xyz_all, int *__restrict__ idx_all) {
  __shared__ __align__(128) float tile[kStages][kTile * 3];
  __shared__ __align__(8) uint64_t bars[kStages];

  const int tid = threadIdx.x;
  const int scene = blockIdx.y;
  const int j = blockIdx.x * kQueries + tid;
  const float *xyz = xyz_all + (size_t)scene * n * 3;
  const bool live = j < m;

  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (live) {
    const float *q = new_xyz_all + ((size_t)scene * m + j) * 3;
    qx = q[0]; qy = q[1]; qz = q[2];
  }
  int *row = idx_all + ((size_t)scene * m + (live ? j : 0)) * nsample;

  // The bulk copy needs 16-byte aligned source and size; a scene starts at
  // scene*n*12 bytes, which is 16-byte aligned only when scene*n is a multiple of 4.
  // `head` points (0..3) are read with plain loads so the bulk part starts aligned.
  const uintptr_t base_addr = reinterpret_cast<uintptr_t>(xyz);
  int head = 0;
  while (head < n && ((base_addr + (size_t)head * 12) & 15)) ++head;
  const int nbulk = ((n - head) / 4) * 4;         // whole 48-byte groups
  const int ntiles = (nbulk + kTile - 1) / kTile;

  const uint32_t bar0 = smem_u32(&bars[0]);
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(bar0 + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue = [&](int t) {
    const int s = t % kStages;
    const int cnt = min(kTile, nbulk - t * kTile);
    const uint32_t bytes = (uint32_t)cnt * 12u;
    mbar_arrive_expect_tx(bar0 + 8 * s, bytes);
    bulk_load(smem_u32(&tile[s][0]), xyz + ((size_t)head + (size_t)t * kTile) * 3, bytes,
              bar0 + 8 * s);
  };
  if (tid == 0) {
    for (int t = 0; t < kStages && t < ntiles; ++t) issue(t);
  }

  int cnt = live ? 0 : nsample;  // dead lanes are "full" from the start
  int first = 0;

  auto visit = [&](float x, float y, float z, int k) {
    const float d2 = sqdist3(qx, qy, qz, x, y, z);
    if (d2 < radius2 && cnt < nsample) {
      if (cnt == 0) first = k;
      row[cnt] = k;
      ++cnt;
    }
  };

  for (int k = 0; k < head; ++k) visit(xyz[k * 3], xyz[k * 3 + 1], xyz[k * 3 + 2], k);

  for (int t = 0; t < ntiles; ++t) {
    const int s = t % kStages;
    mbar_wait(bar0 + 8 * s, (t / kStages) & 1);
    const int tn = min(kTile, nbulk - t * kTile);
    const int kbase = head + t * kTile;
    // whole CTA done -> stop streaming (the reference's `cnt < nsample` loop exit)
    const bool all_full = __syncthreads_and(cnt >= nsample);
    if (all_full) {
      // drain the copies still in flight before this CTA's shared memory goes away
      if (tid == 0)
        for (int u = t + 1; u < ntiles && u < t + kStages; ++u)
          mbar_wait(bar0 + 8 * (u % kStages), (u / kStages) & 1);
      break;
    }
    if (cnt < nsample) {
      const float *tp = &tile[s][0];
#pragma unroll 8
      for (int i = 0; i < tn; ++i) visit(tp[i * 3], tp[i * 3 + 1], tp[i * 3 + 2], kbase + i);
    }
    __syncthreads();  // everyone is done with stage s
    if (tid == 0 && t + kStages < ntiles) issue(t + kStages);
  }
  for (int k = head + nbulk; k < n; ++k) visit(xyz[k * 3], xyz[k * 3 + 1], xyz[k * 3 + 2], k);

  if (live) {
    // ball_query_gpu.cu:33-37: the first hit pre-fills the row; ball_query.cpp:19-21:
    // an empty ball stays zero.
    const int fill = cnt == 0 ? 0 : first;
    for (int l = cnt; l < nsample; ++l) row[l] = fill;
  }
}

}  // namespace

int ball_query_dispatch(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                        const float *xyz, int *idx, cudaStream_t stream) {
  const float radius2 = radius * radius;  // ball_query_gpu.cu:21, fp32 product
  dim3 grid((unsigned)ceil_div(m, kQueries), (unsigned)b);
  ball_query_kernel<<<grid, kQueries, 0, stream>>>(n, m, radius2, nsample, new_xyz, xyz, idx);
  count_launch();
  return check_launch("ball_query_kernel");
}

}  // namespace bqa
