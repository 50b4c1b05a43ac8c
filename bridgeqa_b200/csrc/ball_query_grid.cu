// ball_query_grid.cu -- radius search over a uniform cell grid, for large scenes (sm_100a).
//
// Same contract as ball_query.cu (replaces
// /root/reference/lib/pointnet2/_ext_src/src/ball_query_gpu.cu:9-54): the first `nsample`
// indices k, ascending, with d2 < radius*radius; the first hit back-fills the row; an empty
// ball stays zero.  The reference (and ball_query.cu) evaluate all n points per query; at
// SA1 sizes (n = 40000, ~20-60 hits per ball) 99.9 % of those tests fail.  Here the scene is
// binned once into cells of about one radius (a counting sort: bbox -> count -> scan ->
// scatter, 4 short launches that only depend on xyz, so a caller can run them on a side
// stream while the sampling is still going), and a WARP per query then
//   1. walks the <= 3x3 rows of cells its ball can touch (each row is one contiguous run of
//      the cell-sorted array, read coalesced as float4 {x, y, z, original index}),
//   2. applies the exact reference test d2 = fma(dz,dz,fma(dx,dx,dy*dy)) < r*r to each
//      candidate and appends the hit's ORIGINAL index to a per-warp list in shared memory,
//   3. rank-sorts that list (indices are distinct, so rank = #smaller) and writes the first
//      nsample in ascending index order plus the back-fill -- exactly the row the index-order
//      scan would have produced.
// The hit set is unchanged because the candidate set is a superset of the ball:
//   d2 < r*r  =>  |qx - px| < rfilt (ball_query.cu header)  =>  px in [fl(qx - rfilt), fl(qx + rfilt)]
// (px is a float strictly inside the real interval, so it is bounded by the rounded ends), and
// the cell coordinate is a monotone function of the coordinate evaluated by ONE device
// function for points and interval ends alike; same for y and z.  Cell size never matters for
// correctness (only for speed), so the grid a caller built for one radius serves any radius.
// Lists longer than kHitCap (dense clutter) fall back to the index-order scan of that query.
#include <cmath>
#include <cstdlib>

#include "common.cuh"

namespace bqa {
namespace {

constexpr int kCells = 32768;        // cells per scene (cell size grows until the bbox fits)
constexpr int kHitCap = 512;         // hits a warp can rank-sort in shared memory
constexpr int kWarps = 8;            // queries per CTA
constexpr int kGridMinPoints = 4096; // bqa_ball_query takes this path from here (given a workspace)

struct GridHeader {                  // one per scene, written by grid_bbox_kernel
  float mn[3];
  float inv_h;
  int g[3];
  int cells;
};

struct GridView {                    // carved out of the caller's buffer
  GridHeader *header;                // [b]
  int *start;                        // [b * kCells]  first slot of each cell
  int *cursor;                       // [b * kCells]  after the scatter: one past its last slot
  int *cell_of;                      // [b * n]
  float4 *sorted;                    // [b * n]       {x, y, z, bits(original index)}
};

__device__ __forceinline__ int cell_coord(float x, float mn, float inv_h, int g) {
  // monotone non-decreasing in x: rounded subtraction, multiplication by a positive constant,
  // round-down conversion (NaN -> 0, saturating) and the clamp all preserve order
  const int c = __float2int_rd(__fmul_rn(__fsub_rn(x, mn), inv_h));
  return min(max(c, 0), g - 1);
}

__device__ __forceinline__ float finite_or(float v, float alt) {
  return (fabsf(v) <= 3.0e38f) ? v : alt;      // false for NaN and inf
}

// one CTA per scene: bounding box of the finite coordinates, grid dimensions, zeroed counters
__global__ void __launch_bounds__(1024)
grid_bbox_kernel(int n, float cell, const float *__restrict__ xyz_all, GridView gv) {
  __shared__ float red[6][32];
  __shared__ GridHeader hdr;
  const int scene = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const float *xyz = xyz_all + (size_t)scene * n * 3;
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int k = tid; k < n; k += 1024) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float v = xyz[(size_t)k * 3 + a];
      if (fabsf(v) <= 3.0e38f) { lo[a] = fminf(lo[a], v); hi[a] = fmaxf(hi[a], v); }
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    for (int o = 16; o > 0; o >>= 1) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
    if (lane == 0) { red[a][wid] = lo[a]; red[3 + a][wid] = hi[a]; }
  }
  __syncthreads();
  if (tid == 0) {
    float ext[3];
    for (int a = 0; a < 3; ++a) {
      float l = INFINITY, h = -INFINITY;
      for (int w = 0; w < 32; ++w) { l = fminf(l, red[a][w]); h = fmaxf(h, red[3 + a][w]); }
      l = finite_or(l, 0.f);
      h = finite_or(h, l);
      hdr.mn[a] = l;
      ext[a] = fmaxf(h - l, 0.f);
    }
    float h = finite_or(cell, 1.f);
    if (!(h > 1e-12f)) h = 1e-12f;
    int g[3] = {1, 1, 1};
    bool ok = false;
    for (int it = 0; it < 200 && !ok; ++it) {
      long long cells = 1;
      for (int a = 0; a < 3; ++a) {
        g[a] = (int)fminf(ext[a] / h, 65535.f) + 1;
        cells *= g[a];
      }
      ok = cells <= kCells;
      if (!ok) h *= 1.25f;
    }
    if (!ok) g[0] = g[1] = g[2] = 1;
    float inv_h = 1.f / h;
    if (!(inv_h > 0.f) || !(fabsf(inv_h) <= 3.0e38f)) { inv_h = 1.f; g[0] = g[1] = g[2] = 1; }
    hdr.inv_h = inv_h;
    hdr.g[0] = g[0]; hdr.g[1] = g[1]; hdr.g[2] = g[2];
    hdr.cells = g[0] * g[1] * g[2];
    gv.header[scene] = hdr;
  }
  __syncthreads();
  int *start = gv.start + (size_t)scene * kCells;
  for (int c = tid; c < hdr.cells; c += 1024) start[c] = 0;
}

__global__ void __launch_bounds__(256)
grid_count_kernel(int n, const float *__restrict__ xyz_all, GridView gv) {
  const int scene = blockIdx.y;
  const int k = blockIdx.x * 256 + threadIdx.x;
  if (k >= n) return;
  const GridHeader h = gv.header[scene];
  const float *p = xyz_all + ((size_t)scene * n + k) * 3;
  const int cx = cell_coord(p[0], h.mn[0], h.inv_h, h.g[0]);
  const int cy = cell_coord(p[1], h.mn[1], h.inv_h, h.g[1]);
  const int cz = cell_coord(p[2], h.mn[2], h.inv_h, h.g[2]);
  const int c = (cz * h.g[1] + cy) * h.g[0] + cx;
  gv.cell_of[(size_t)scene * n + k] = c;
  atomicAdd(&gv.start[(size_t)scene * kCells + c], 1);
}

// one CTA per scene: counts -> exclusive prefix sums (start) and a working copy (cursor).
// Each thread owns a run of cells (a multiple of 4, <= 64) that it keeps in registers as int4.
__global__ void __launch_bounds__(1024)
grid_scan_kernel(GridView gv) {
  __shared__ int warp_sum[32];
  const int scene = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int cells = gv.header[scene].cells;
  int4 *start4 = reinterpret_cast<int4 *>(gv.start + (size_t)scene * kCells);
  int4 *cursor4 = reinterpret_cast<int4 *>(gv.cursor + (size_t)scene * kCells);
  const int per4 = ((cells + 1023) / 1024 + 3) / 4;      // int4 groups per thread, 1..16
  const int g0 = tid * per4;                             // first group of this thread
  constexpr int kMaxGroups = kCells / 1024 / 4;
  int4 v[kMaxGroups];
  int sum = 0;
#pragma unroll
  for (int i = 0; i < kMaxGroups; ++i) {
    v[i] = make_int4(0, 0, 0, 0);
    if (i < per4) {
      const int c = (g0 + i) * 4;
      if (c < cells) {
        v[i] = start4[g0 + i];
        if (c + 1 >= cells) v[i].y = 0;                   // beyond the grid: never zeroed, ignore
        if (c + 2 >= cells) v[i].z = 0;
        if (c + 3 >= cells) v[i].w = 0;
      }
      sum += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
  int incl = sum;
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_sum[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int w = warp_sum[lane];
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    warp_sum[lane] = w;
  }
  __syncthreads();
  int run = incl - sum + (wid ? warp_sum[wid - 1] : 0);
#pragma unroll
  for (int i = 0; i < kMaxGroups; ++i) {
    if (i < per4 && (g0 + i) * 4 < cells) {
      int4 o;
      o.x = run; run += v[i].x;
      o.y = run; run += v[i].y;
      o.z = run; run += v[i].z;
      o.w = run; run += v[i].w;
      start4[g0 + i] = o;
      cursor4[g0 + i] = o;
    }
  }
}

__global__ void __launch_bounds__(256)
grid_scatter_kernel(int n, const float *__restrict__ xyz_all, GridView gv) {
  const int scene = blockIdx.y;
  const int k = blockIdx.x * 256 + threadIdx.x;
  if (k >= n) return;
  const float *p = xyz_all + ((size_t)scene * n + k) * 3;
  const int c = gv.cell_of[(size_t)scene * n + k];
  const int pos = atomicAdd(&gv.cursor[(size_t)scene * kCells + c], 1);
  gv.sorted[(size_t)scene * n + pos] = make_float4(p[0], p[1], p[2], __int_as_float(k));
}

__global__ void __launch_bounds__(kWarps * 32)
ball_query_grid_kernel(int n, int m, int q_stride, int q_offset, float radius2, float rfilt, int nsample,
                       const float *__restrict__ new_xyz_all, const float *__restrict__ xyz_all,
                       int *__restrict__ idx_all, GridView gv) {
  __shared__ __align__(16) int s_hits[kWarps][kHitCap];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int scene = blockIdx.y;
  const int j = blockIdx.x * kWarps + wid;
  if (j >= m) return;                               // warp-uniform; no block barrier below
  const unsigned lt = (1u << lane) - 1u;
  const size_t qglob = (size_t)scene * q_stride + q_offset + j;
  const float qx = new_xyz_all[qglob * 3], qy = new_xyz_all[qglob * 3 + 1], qz = new_xyz_all[qglob * 3 + 2];
  int *row = idx_all + qglob * nsample;
  int *hits = s_hits[wid];

  const GridHeader h = gv.header[scene];
  const int *start = gv.start + (size_t)scene * kCells;
  const int *cursor = gv.cursor + (size_t)scene * kCells;
  const float4 *sorted = gv.sorted + (size_t)scene * n;
  const int cx0 = cell_coord(__fsub_rn(qx, rfilt), h.mn[0], h.inv_h, h.g[0]);
  const int cx1 = cell_coord(__fadd_rn(qx, rfilt), h.mn[0], h.inv_h, h.g[0]);
  const int cy0 = cell_coord(__fsub_rn(qy, rfilt), h.mn[1], h.inv_h, h.g[1]);
  const int cy1 = cell_coord(__fadd_rn(qy, rfilt), h.mn[1], h.inv_h, h.g[1]);
  const int cz0 = cell_coord(__fsub_rn(qz, rfilt), h.mn[2], h.inv_h, h.g[2]);
  const int cz1 = cell_coord(__fadd_rn(qz, rfilt), h.mn[2], h.inv_h, h.g[2]);

  int count = 0;
  bool overflow = false;
  // a NaN centre has no hits; its cell range above is still well defined (cell 0)
  const int nrows = (cy1 - cy0 + 1) * (cz1 - cz0 + 1);
  const int ny = cy1 - cy0 + 1;
  for (int r0 = 0; r0 < nrows && !overflow; r0 += 32) {
    // lanes fetch the extents of up to 32 cell rows at once, then the warp walks them
    int seg_b = 0, seg_e = 0;
    if (r0 + lane < nrows) {
      const int r = r0 + lane;
      const int cz = cz0 + r / ny, cy = cy0 + r % ny;
      const int base = (cz * h.g[1] + cy) * h.g[0];
      seg_b = __ldg(&start[base + cx0]);
      seg_e = __ldg(&cursor[base + cx1]);
    }
    const int rows_here = min(32, nrows - r0);
    for (int r = 0; r < rows_here && !overflow; ++r) {
      const int sb = __shfl_sync(0xffffffffu, seg_b, r), se = __shfl_sync(0xffffffffu, seg_e, r);
      for (int base = sb; base < se; base += 32) {
        const int p = base + lane;
        bool hit = false;
        int k = 0;
        if (p < se) {
          const float4 v = __ldg(&sorted[p]);
          hit = sqdist3(qx, qy, qz, v.x, v.y, v.z) < radius2;
          k = __float_as_int(v.w);
        }
        const unsigned mask = __ballot_sync(0xffffffffu, hit);
        if (mask) {
          const int add = __popc(mask);
          if (count + add > kHitCap) { overflow = true; break; }
          if (hit) hits[count + __popc(mask & lt)] = k;
          count += add;
        }
      }
    }
  }
  __syncwarp();

  if (overflow) {
    // dense clutter: the reference's own index-order scan for this query (stops at nsample)
    const float *xyz = xyz_all + (size_t)scene * n * 3;
    int cnt = 0, first = 0;
    for (int base = 0; base < n && cnt < nsample; base += 32) {
      const int k = base + lane;
      const bool hit = k < n && sqdist3(qx, qy, qz, xyz[(size_t)k * 3], xyz[(size_t)k * 3 + 1],
                                        xyz[(size_t)k * 3 + 2]) < radius2;
      const unsigned mask = __ballot_sync(0xffffffffu, hit);
      if (mask) {
        if (cnt == 0) first = base + __ffs(mask) - 1;
        const int pos = cnt + __popc(mask & lt);
        if (hit && pos < nsample) row[pos] = k;
        cnt += __popc(mask);
      }
    }
    cnt = min(cnt, nsample);
    for (int l = cnt + lane; l < nsample; l += 32) row[l] = first;      // first == 0 for an empty ball
    return;
  }

  // rank sort: indices are distinct, so the number of smaller ones is the output slot
  int smallest = 0x7fffffff;
  for (int e0 = 0; e0 < count; e0 += 32) {
    const int e = e0 + lane;
    const int mine = e < count ? hits[e] : 0x7fffffff;
    int rank = 0;
    int t = 0;
    for (; t + 4 <= count; t += 4) {
      const int4 o = *reinterpret_cast<const int4 *>(&hits[t]);
      rank += (o.x < mine) + (o.y < mine) + (o.z < mine) + (o.w < mine);
    }
    for (; t < count; ++t) rank += hits[t] < mine;
    if (e < count && rank < nsample) row[rank] = mine;
    smallest = min(smallest, mine);
  }
  smallest = __reduce_min_sync(0xffffffffu, smallest);
  const int fill = count ? smallest : 0;
  for (int l = count + lane; l < nsample; l += 32) row[l] = fill;
}

long long align256(long long v) { return (v + 255) / 256 * 256; }

GridView carve(void *grid, int b, int n) {
  char *p = reinterpret_cast<char *>(grid);
  GridView gv;
  gv.header = reinterpret_cast<GridHeader *>(p);      p += align256((long long)b * sizeof(GridHeader));
  gv.start = reinterpret_cast<int *>(p);              p += align256(4ll * b * kCells);
  gv.cursor = reinterpret_cast<int *>(p);             p += align256(4ll * b * kCells);
  gv.cell_of = reinterpret_cast<int *>(p);            p += align256(4ll * b * n);
  gv.sorted = reinterpret_cast<float4 *>(p);
  return gv;
}

}  // namespace

const float4 *ball_query_grid_sorted(const void *grid, int b, int n) {
  return carve(const_cast<void *>(grid), b, n).sorted;
}

long long ball_query_grid_bytes(int b, int n) {
  return align256((long long)b * sizeof(GridHeader)) + 2 * align256(4ll * b * kCells) +
         align256(4ll * b * n) + align256(16ll * b * n);
}

// worth it once the index-order scan is long; below that the scan kernel wins on launch count
// (5 launches here against 1).  BQA_BQ_GRID_MIN overrides the threshold (measurement only).
bool ball_query_grid_wanted(int n, int m) {
  static const int min_points = [] {
    const char *e = getenv("BQA_BQ_GRID_MIN");
    return e ? atoi(e) : kGridMinPoints;
  }();
  return n >= min_points && m >= 1;
}

int ball_query_grid_build(int b, int n, float radius, const float *xyz, void *grid, cudaStream_t stream) {
  if (b > 65535) return set_error(BQA_ERR_UNSUPPORTED, "ball_query grid: batch too large");
  GridView gv = carve(grid, b, n);
  const float radius2 = radius * radius;
  const float rfilt = nextafterf((float)(sqrt((double)radius2) * (1.0 + 1e-6)), INFINITY);
  grid_bbox_kernel<<<b, 1024, 0, stream>>>(n, rfilt, xyz, gv);
  count_launch();
  if (int rc = check_launch("grid_bbox_kernel")) return rc;
  dim3 pg((unsigned)ceil_div(n, 256), (unsigned)b);
  grid_count_kernel<<<pg, 256, 0, stream>>>(n, xyz, gv);
  count_launch();
  if (int rc = check_launch("grid_count_kernel")) return rc;
  grid_scan_kernel<<<b, 1024, 0, stream>>>(gv);
  count_launch();
  if (int rc = check_launch("grid_scan_kernel")) return rc;
  grid_scatter_kernel<<<pg, 256, 0, stream>>>(n, xyz, gv);
  count_launch();
  return check_launch("grid_scatter_kernel");
}

int ball_query_grid_search(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                           const float *xyz, int *idx, const void *grid, cudaStream_t stream,
                           int q_stride, int q_offset) {
  if (b > 65535) return set_error(BQA_ERR_UNSUPPORTED, "ball_query grid: batch too large");
  GridView gv = carve(const_cast<void *>(grid), b, n);
  const float radius2 = radius * radius;  // ball_query_gpu.cu:21, fp32 product
  const float rfilt = nextafterf((float)(sqrt((double)radius2) * (1.0 + 1e-6)), INFINITY);
  dim3 qg((unsigned)ceil_div(m, kWarps), (unsigned)b);
  ball_query_grid_kernel<<<qg, kWarps * 32, 0, stream>>>(n, m, q_stride, q_offset, radius2, rfilt, nsample,
                                                         new_xyz, xyz, idx, gv);
  count_launch();
  return check_launch("ball_query_grid_kernel");
}

}  // namespace bqa
