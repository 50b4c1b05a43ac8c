// ball_query.cu -- radius search with the reference's "first nsample in index order"
// semantics, for sm_100a.
//
// Replaces /root/reference/lib/pointnet2/_ext_src/src/ball_query_gpu.cu:9-54
// (grid = B, each thread scans all n points from global memory for m/512 queries).
//
// Layout: one thread owns one query (its centre stays in registers); a CTA of 128
// queries walks xyz in tiles staged in shared memory by the bulk-copy engine
// (cp.async.bulk -> mbarrier, double buffered), so every point is read from L2 once per
// CTA and then broadcast to all lanes: 4 points = three 16-byte shared loads, 24 FMA-pipe
// ops and ONE compare (hits are rare, ~25 in 40000, so the per-point bookkeeping sits
// behind that compare).
//
// A thread-per-query scan of 40000 points is latency-bound with the 1024 warps the
// BASELINE batch offers (measured 42 % issue utilisation), so large scenes are cut into S
// index-ordered segments handled by different CTAs (S x more warps).  Each segment appends
// its hits to its own list in a workspace; the LAST CTA to finish a block of 128 queries
// (atomic ticket) concatenates the lists in segment order, which is exactly "the first
// nsample hits in ascending index", applies the reference's back-fill with the first hit
// and writes the rows coalesced.
//
// Scenes of >= kGridMinPoints points (with a workspace) do not come here at all: they are
// binned into a cell grid and searched by ball_query_grid.cu.
//
// Bit-exact: d2 = fma(dz,dz,fma(dx,dx,dy*dy)) (nvcc's contraction of
// ball_query_gpu.cu:30-31), compared `<` against radius*radius computed in fp32.
#include <cmath>

#include "common.cuh"

namespace bqa {
namespace {

constexpr int kQueries = 128;      // threads per CTA == queries per CTA
constexpr int kTile = 960;         // points per shared-memory tile (11.25 KB as raw xyz)
constexpr int kStages = 2;

__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes,
                                          uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

struct BqWorkspace {       // carved out of the caller's workspace when S > 1
  unsigned int *ticket;    // [b * qblocks]   zeroed by the dispatcher (S > 1)
  int *count;              // [b * m * S]
  int *hits;               // [b * m * S * nsample]
};

__global__ void __launch_bounds__(kQueries)
ball_query_kernel(int n, int m, int q_stride, int q_offset, float radius2, int nsample,
                  int nseg, int seg_len,
                  const float *__restrict__ new_xyz_all, const float *__restrict__ xyz_all,
                  int *__restrict__ idx_all, BqWorkspace ws) {
  __shared__ __align__(128) float tile[kStages][kTile * 3];
  __shared__ __align__(8) uint64_t bars[kStages];
  __shared__ int s_last;

  const int tid = threadIdx.x;
  const int scene = blockIdx.z;
  const int seg = blockIdx.y;
  const int j = blockIdx.x * kQueries + tid;
  const float *xyz = xyz_all + (size_t)scene * n * 3;
  const bool live = j < m;
  const int k_begin = seg * seg_len;
  const int k_end = min(n, k_begin + seg_len);

  float qx = 0.f, qy = 0.f, qz = 0.f;
  // the m queries of a scene are centres [q_offset, q_offset + m) of its q_stride centres (a
  // slice of the sampling); the workspace is indexed by the local query number
  const size_t qglob = (size_t)scene * q_stride + q_offset;
  if (live) {
    const float *q = new_xyz_all + (qglob + j) * 3;
    qx = q[0]; qy = q[1]; qz = q[2];
  }
  // where this thread appends its hits: the final row (one segment) or its segment's list
  const size_t qlin = (size_t)scene * m + (live ? j : 0);
  int *row = nseg == 1 ? idx_all + (qglob + (live ? j : 0)) * nsample
                       : ws.hits + (qlin * nseg + seg) * nsample;

  // The bulk copy needs 16-byte aligned source and size.  `head` points (0..3) are read
  // with plain loads so the bulk part starts aligned; the tail (< 4 points) likewise.
  const uintptr_t base_addr = reinterpret_cast<uintptr_t>(xyz + (size_t)k_begin * 3);
  int head = 0;
  while (k_begin + head < k_end && ((base_addr + (size_t)head * 12) & 15)) ++head;
  const int nbulk = max(0, (k_end - k_begin - head) / 4) * 4;   // whole 48-byte groups
  const int ntiles = (nbulk + kTile - 1) / kTile;

  const uint32_t bar0 = smem_u32(&bars[0]);
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(bar0 + 8 * s, 1);
    fence_mbar_init_cluster();
  }
  __syncthreads();

  auto issue = [&](int t) {
    const int s = t % kStages;
    const int cnt = min(kTile, nbulk - t * kTile);
    const uint32_t bytes = (uint32_t)cnt * 12u;
    mbar_arrive_expect_tx(bar0 + 8 * s, bytes);
    bulk_load(smem_u32(&tile[s][0]), xyz + ((size_t)k_begin + head + (size_t)t * kTile) * 3, bytes,
              bar0 + 8 * s);
  };
  if (tid == 0) {
    for (int t = 0; t < kStages && t < ntiles; ++t) issue(t);
  }

  int cnt = live ? 0 : nsample;  // dead lanes are "full" from the start
  int first = 0;

  auto hit = [&](float d2, int k) {
    if (d2 < radius2 && cnt < nsample) {
      if (cnt == 0) first = k;
      row[cnt] = k;
      ++cnt;
    }
  };
  auto visit = [&](int k) {
    hit(sqdist3(qx, qy, qz, xyz[(size_t)k * 3], xyz[(size_t)k * 3 + 1], xyz[(size_t)k * 3 + 2]), k);
  };

  for (int k = k_begin; k < min(k_end, k_begin + head); ++k) visit(k);

  for (int t = 0; t < ntiles; ++t) {
    const int s = t % kStages;
    mbar_wait(bar0 + 8 * s, (t / kStages) & 1);
    const int tn = min(kTile, nbulk - t * kTile);
    const int kbase = k_begin + head + t * kTile;
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
      const float4 *t4 = reinterpret_cast<const float4 *>(&tile[s][0]);
#pragma unroll 2
      for (int g = 0; g < tn / 4; ++g) {
        const float4 a = t4[3 * g], bq = t4[3 * g + 1], c = t4[3 * g + 2];
        const float d0 = sqdist3(qx, qy, qz, a.x, a.y, a.z);
        const float d1 = sqdist3(qx, qy, qz, a.w, bq.x, bq.y);
        const float d2 = sqdist3(qx, qy, qz, bq.z, bq.w, c.x);
        const float d3 = sqdist3(qx, qy, qz, c.y, c.z, c.w);
        if (fminf(fminf(d0, d1), fminf(d2, d3)) < radius2) {
          const int k = kbase + 4 * g;
          hit(d0, k); hit(d1, k + 1); hit(d2, k + 2); hit(d3, k + 3);
        }
      }
    }
    __syncthreads();  // everyone is done with stage s
    if (tid == 0 && t + kStages < ntiles) issue(t + kStages);
  }
  for (int k = k_begin + head + nbulk; k < k_end; ++k) visit(k);

  if (nseg == 1) {
    if (live) {
      // ball_query_gpu.cu:33-37: the first hit pre-fills the row; ball_query.cpp:19-21:
      // an empty ball stays zero.
      const int fill = cnt == 0 ? 0 : first;
      for (int l = cnt; l < nsample; ++l) row[l] = fill;
    }
    return;
  }

  // ---- S > 1: publish this segment's list, last CTA of the query block merges ------------
  if (live) ws.count[qlin * nseg + seg] = cnt;
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int t = atomicAdd(&ws.ticket[scene * gridDim.x + blockIdx.x], 1u);
    s_last = (t == (unsigned)nseg - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int q0 = blockIdx.x * kQueries;
  const int nq = min(kQueries, m - q0);
  for (int e = tid; e < nq * nsample; e += kQueries) {      // consecutive threads = consecutive slots
    const int q = e / nsample, slot = e - q * nsample;
    const int jq = q0 + q;
    const size_t ql = ((size_t)scene * m + jq) * nseg;
    int remaining = slot, value = 0, first_hit = 0;
    bool found = false, seen = false;
    for (int sgm = 0; sgm < nseg; ++sgm) {
      const int c = __ldcg(&ws.count[ql + sgm]);
      if (c > 0 && !seen) { first_hit = __ldcg(&ws.hits[(ql + sgm) * nsample]); seen = true; }
      if (!found && remaining < c) {
        value = __ldcg(&ws.hits[(ql + sgm) * nsample + remaining]);
        found = true;
      }
      remaining -= c;
    }
    idx_all[(qglob + jq) * nsample + slot] = found ? value : first_hit;
  }
}

int plan_segments(int b, int n, int m) {
  // enough CTAs for ~6 per SM (24 warps), but never segments shorter than 2048 points
  const int qblocks = ceil_div(m, kQueries) * b;
  int s = 1;
  while (s < 8 && qblocks * s < 148 * 6 && n / (s * 2) >= 2048) s *= 2;
  return s;
}

long long ticket_bytes(int b, int m) {
  return ((long long)b * ceil_div(m, kQueries) * 4 + 255) / 256 * 256;
}

}  // namespace

// ball_query_grid.cu
long long ball_query_grid_bytes(int b, int n);
bool ball_query_grid_wanted(int n, int m);
int ball_query_grid_build(int b, int n, float radius, const float *xyz, void *grid, cudaStream_t stream);
int ball_query_grid_search(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                           const float *xyz, int *idx, const void *grid, cudaStream_t stream,
                           int q_stride, int q_offset);

long long ball_query_workspace_bytes(int b, int n, int m, int nsample) {
  if (ball_query_grid_wanted(n, m)) return ball_query_grid_bytes(b, n);
  const int s = plan_segments(b, n, m);
  if (s == 1) return 0;
  return ticket_bytes(b, m) + 4ll * b * m * s + 4ll * b * m * s * nsample;
}

int ball_query_dispatch(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                        const float *xyz, int *idx, void *workspace, cudaStream_t stream,
                        int q_stride, int q_offset) {
  if (workspace && ball_query_grid_wanted(n, m)) {
    // large scene: bin once, then a warp per query over the cells its ball touches
    if (int rc = ball_query_grid_build(b, n, radius, xyz, workspace, stream)) return rc;
    return ball_query_grid_search(b, n, m, radius, nsample, new_xyz, xyz, idx, workspace, stream,
                                  q_stride, q_offset);
  }
  const float radius2 = radius * radius;  // ball_query_gpu.cu:21, fp32 product
  const int nseg = workspace ? plan_segments(b, n, m) : 1;
  BqWorkspace ws = {nullptr, nullptr, nullptr};
  int seg_len = n;
  char *wsp = reinterpret_cast<char *>(workspace);
  if (nseg > 1) {
    const long long tickets = ticket_bytes(b, m);
    ws.ticket = reinterpret_cast<unsigned int *>(wsp);
    ws.count = reinterpret_cast<int *>(wsp + tickets);
    ws.hits = ws.count + (size_t)b * m * nseg;
    BQA_CUDA(cudaMemsetAsync(ws.ticket, 0, (size_t)tickets, stream));
    seg_len = ceil_div(ceil_div(n, nseg), 4) * 4;     // multiple of 4 points keeps 16-byte alignment
  }
  if (b > 65535) return set_error(BQA_ERR_UNSUPPORTED, "ball_query: batch too large");
  dim3 grid((unsigned)ceil_div(m, kQueries), (unsigned)nseg, (unsigned)b);
  ball_query_kernel<<<grid, kQueries, 0, stream>>>(n, m, q_stride, q_offset, radius2, nsample, nseg,
                                                   seg_len, new_xyz, xyz, idx, ws);
  count_launch();
  return check_launch("ball_query_kernel");
}

}  // namespace bqa
