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
// Queries are handed to threads in ascending x (a per-scene bitonic sort of the m centres),
// so the 32 queries of a warp lie in a thin x-slab and one |dx| test per point -- 4 FADD,
// two |.|-FMNMX and one compare per 4 points -- rejects ~93 % of the points for the whole
// warp before the exact 3-D distance is evaluated.  The filter is conservative
// (|dx| >= r(1+1e-6) implies fl(dx*dx) >= r*r, and the exact d2 >= fl(dx*dx) because the
// other two terms are non-negative and rounding is monotone), so the hit set is unchanged.
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

struct BqWorkspace {       // carved out of the caller's workspace
  int *perm;               // [b * m]  queries in ascending x (NULL: identity)
  unsigned int *ticket;    // [b * qblocks]   zeroed by the dispatcher (S > 1)
  int *count;              // [b * m * S]
  int *hits;               // [b * m * S * nsample]
};

constexpr int kSortMax = 4096;     // centres per scene the shared-memory sort handles

// one CTA per scene: perm = argsort(new_xyz[:, 0]) (ties by index), bitonic in shared memory
__global__ void __launch_bounds__(1024)
sort_queries_kernel(int m, int q_stride, int q_offset, const float *__restrict__ new_xyz_all,
                    int *__restrict__ perm_all) {
  __shared__ unsigned long long keys[kSortMax];
  const int scene = blockIdx.x;
  int npow = 1;
  while (npow < m) npow <<= 1;
  for (int i = threadIdx.x; i < npow; i += blockDim.x) {
    unsigned long long k = ~0ull;
    if (i < m) {
      const uint32_t u = __float_as_uint(new_xyz_all[((size_t)scene * q_stride + q_offset + i) * 3]);
      const uint32_t ord = (u & 0x80000000u) ? ~u : (u | 0x80000000u);     // float order -> uint order
      k = ((unsigned long long)ord << 32) | (uint32_t)i;
    }
    keys[i] = k;
  }
  __syncthreads();
  for (int size = 2; size <= npow; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < npow / 2; i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool up = (lo & size) == 0;
        const unsigned long long a = keys[lo], c = keys[hi];
        if ((a > c) == up) { keys[lo] = c; keys[hi] = a; }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < m; i += blockDim.x) perm_all[(size_t)scene * m + i] = (int)(uint32_t)keys[i];
}

__global__ void __launch_bounds__(kQueries)
ball_query_kernel(int n, int m, int q_stride, int q_offset, float radius2, float rfilt, int nsample,
                  int nseg, int seg_len,
                  const float *__restrict__ new_xyz_all, const float *__restrict__ xyz_all,
                  int *__restrict__ idx_all, BqWorkspace ws) {
  __shared__ __align__(128) float tile[kStages][kTile * 3];
  __shared__ __align__(8) uint64_t bars[kStages];
  __shared__ int s_last;

  const int tid = threadIdx.x;
  const int scene = blockIdx.z;
  const int seg = blockIdx.y;
  const int slot = blockIdx.x * kQueries + tid;
  const float *xyz = xyz_all + (size_t)scene * n * 3;
  const bool live = slot < m;
  const int j = (live && ws.perm) ? ws.perm[(size_t)scene * m + slot] : slot;
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
        // x-slab filter (conservative, see the header): most groups end here for the whole warp;
        // behind it each point is re-tested on its own so the exact distance is only evaluated
        // for the ~1 in 4 points of a passing group that can actually be inside the ball
        const float e0 = fabsf(qx - a.x), e1 = fabsf(qx - a.w), e2 = fabsf(qx - bq.z), e3 = fabsf(qx - c.y);
        if (!(fminf(fminf(e0, e1), fminf(e2, e3)) < rfilt)) continue;
        const int k = kbase + 4 * g;
        if (e0 < rfilt) hit(sqdist3(qx, qy, qz, a.x, a.y, a.z), k);
        if (e1 < rfilt) hit(sqdist3(qx, qy, qz, a.w, bq.x, bq.y), k + 1);
        if (e2 < rfilt) hit(sqdist3(qx, qy, qz, bq.z, bq.w, c.x), k + 2);
        if (e3 < rfilt) hit(sqdist3(qx, qy, qz, c.y, c.z, c.w), k + 3);
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
    const int jq = ws.perm ? __ldg(&ws.perm[(size_t)scene * m + q0 + q]) : q0 + q;
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
long long perm_bytes(int b, int m) { return ((long long)b * m * 4 + 255) / 256 * 256; }
// sorting pays off once the scan is long enough to amortise the extra launch
bool want_sort(int n, int m) { return m >= 64 && m <= kSortMax && n >= 4096; }

}  // namespace

long long ball_query_workspace_bytes(int b, int n, int m, int nsample) {
  const int s = plan_segments(b, n, m);
  long long bytes = want_sort(n, m) ? perm_bytes(b, m) : 0;
  if (s > 1) bytes += ticket_bytes(b, m) + 4ll * b * m * s + 4ll * b * m * s * nsample;
  return bytes;
}

int ball_query_dispatch(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                        const float *xyz, int *idx, void *workspace, cudaStream_t stream,
                        int q_stride, int q_offset) {
  const float radius2 = radius * radius;  // ball_query_gpu.cu:21, fp32 product
  const int nseg = workspace ? plan_segments(b, n, m) : 1;
  BqWorkspace ws = {nullptr, nullptr, nullptr, nullptr};
  int seg_len = n;
  char *wsp = reinterpret_cast<char *>(workspace);
  if (workspace && want_sort(n, m)) {
    ws.perm = reinterpret_cast<int *>(wsp);
    wsp += perm_bytes(b, m);
    sort_queries_kernel<<<b, 1024, 0, stream>>>(m, q_stride, q_offset, new_xyz, ws.perm);
    count_launch();
    if (int rc = check_launch("sort_queries_kernel")) return rc;
  }
  // |dx| >= rfilt  =>  fl(dx*dx) >= radius2  (1e-6 relative margin, rounded up)
  const float rfilt = nextafterf((float)(sqrt((double)radius2) * (1.0 + 1e-6)), INFINITY);
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
  ball_query_kernel<<<grid, kQueries, 0, stream>>>(n, m, q_stride, q_offset, radius2, rfilt, nsample, nseg,
                                                   seg_len, new_xyz, xyz, idx, ws);
  count_launch();
  return check_launch("ball_query_kernel");
}

}  // namespace bqa
