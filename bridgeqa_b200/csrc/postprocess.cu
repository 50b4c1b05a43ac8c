// postprocess.cu -- detection post-processing on the device, for sm_100a.
//
// Replaces the host loops of /root/reference/lib/ap_helper.py:86-178 (`parse_predictions`): the
// per-proposal point-in-box counts over the 40k-point cloud (`remove_empty_box`, :89-100) and the
// greedy 3-D NMS of /root/reference/utils/nms.py:74-152 (`nms_3d_faster`, `nms_3d_faster_samecls`).
//
// NMS semantics (one CTA per scene, double precision like the reference's float64 NumPy arrays):
// boxes are visited in descending score; a visited box that is still alive is picked and
// suppresses every later box j whose overlap o = inter / (area_i + area_j - inter)  (or
// inter / area_j with old_type) exceeds the threshold -- and, for the same-class variant, only if
// cls_i == cls_j.  l, w, h = max(0, min(upper) - max(lower)).  The reference walks np.argsort(score)
// from its END; NumPy's default argsort leaves the order of equal scores unspecified (its AVX-512
// sort is not stable even for 8 elements), so ties are resolved as a STABLE ascending argsort
// walked from the end would: among equal scores the HIGHER index is visited first
// (tests/golden/make_golden_post.py records both the default and the kind="stable" reference runs).
#include "common.cuh"

namespace bqa {
namespace {

constexpr int kNmsMax = 1024;
constexpr int kNmsThreads = 256;

__global__ void __launch_bounds__(kNmsThreads)
nms3d_kernel(int k, const float *__restrict__ boxes_all, const int *__restrict__ valid_all, double thr,
             int old_type, int same_cls, int *__restrict__ pick_all, int *__restrict__ order_all) {
  // boxes (b, k, 8): x1 y1 z1 x2 y2 z2 score cls
  __shared__ unsigned long long keys[kNmsMax];
  __shared__ unsigned char alive[kNmsMax];
  __shared__ int s_cur, s_npick;
  const int scene = blockIdx.x, tid = threadIdx.x;
  const float *boxes = boxes_all + (size_t)scene * k * 8;
  const int *valid = valid_all ? valid_all + (size_t)scene * k : nullptr;
  int *pick = pick_all + (size_t)scene * k;
  int *order = order_all ? order_all + (size_t)scene * k : nullptr;
  int npow = 1;
  while (npow < k) npow <<= 1;
  for (int i = tid; i < npow; i += kNmsThreads) {
    unsigned long long key = ~0ull;                       // invalid / padding sorts last
    if (i < k && (!valid || valid[i])) {
      const uint32_t u = __float_as_uint(boxes[i * 8 + 6]);
      const uint32_t ord = (u & 0x80000000u) ? ~u : (u | 0x80000000u);      // float order -> uint order
      key = ((unsigned long long)(~ord) << 32) | (uint32_t)(~(uint32_t)i);  // descending score, then descending index
    }
    keys[i] = key;
    if (i < k) { pick[i] = 0; if (order) order[i] = -1; }
  }
  __syncthreads();
  for (int size = 2; size <= npow; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < npow / 2; i += kNmsThreads) {
        const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
        const bool up = (lo & size) == 0;
        const unsigned long long a = keys[lo], c = keys[hi];
        if ((a > c) == up) { keys[lo] = c; keys[hi] = a; }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < npow; i += kNmsThreads) alive[i] = keys[i] != ~0ull;
  if (tid == 0) { s_cur = 0; s_npick = 0; }
  __syncthreads();
  while (true) {
    if (tid == 0) {
      int c = s_cur;
      while (c < k && !alive[c]) ++c;
      s_cur = c;
      if (c < k) {
        const int bi = (int)(~(uint32_t)keys[c]);
        pick[bi] = 1;
        if (order) order[s_npick] = bi;
        ++s_npick;
      }
    }
    __syncthreads();
    const int c = s_cur;
    if (c >= k) break;
    const int bi = (int)(~(uint32_t)keys[c]);
    const double ix1 = boxes[bi * 8 + 0], iy1 = boxes[bi * 8 + 1], iz1 = boxes[bi * 8 + 2];
    const double ix2 = boxes[bi * 8 + 3], iy2 = boxes[bi * 8 + 4], iz2 = boxes[bi * 8 + 5];
    const float icls = boxes[bi * 8 + 7];
    const double iarea = __dmul_rn(__dmul_rn(ix2 - ix1, iy2 - iy1), iz2 - iz1);
    for (int p = c + 1 + tid; p < k; p += kNmsThreads) {
      if (!alive[p]) continue;
      const int bj = (int)(~(uint32_t)keys[p]);
      const double jx1 = boxes[bj * 8 + 0], jy1 = boxes[bj * 8 + 1], jz1 = boxes[bj * 8 + 2];
      const double jx2 = boxes[bj * 8 + 3], jy2 = boxes[bj * 8 + 4], jz2 = boxes[bj * 8 + 5];
      const double l = fmax(0.0, fmin(ix2, jx2) - fmax(ix1, jx1));
      const double w = fmax(0.0, fmin(iy2, jy2) - fmax(iy1, jy1));
      const double h = fmax(0.0, fmin(iz2, jz2) - fmax(iz1, jz1));
      const double inter = __dmul_rn(__dmul_rn(l, w), h);
      const double jarea = __dmul_rn(__dmul_rn(jx2 - jx1, jy2 - jy1), jz2 - jz1);
      const double o = old_type ? inter / jarea : inter / (__dadd_rn(iarea, jarea) - inter);
      if (o > thr && (!same_cls || boxes[bj * 8 + 7] == icls)) alive[p] = 0;
    }
    if (tid == 0) { alive[c] = 0; }
    __syncthreads();
  }
}

// counts[b, j] = number of points p of scene b with lo_j <= p <= hi_j (component-wise, inclusive)
constexpr int kBoxesPerCta = 8;
__global__ void __launch_bounds__(256)
count_in_boxes_kernel(int n, int k, const float *__restrict__ xyz_all, const float *__restrict__ lohi_all,
                      int *__restrict__ counts_all) {
  __shared__ int s_cnt[kBoxesPerCta];
  const int scene = blockIdx.y;
  const int j0 = blockIdx.x * kBoxesPerCta;
  const float *xyz = xyz_all + (size_t)scene * n * 3;
  float lo[kBoxesPerCta][3], hi[kBoxesPerCta][3];
  int cnt[kBoxesPerCta];
#pragma unroll
  for (int u = 0; u < kBoxesPerCta; ++u) {
    const int j = min(j0 + u, k - 1);
    const float *bx = lohi_all + ((size_t)scene * k + j) * 6;
#pragma unroll
    for (int a = 0; a < 3; ++a) { lo[u][a] = bx[a]; hi[u][a] = bx[3 + a]; }
    cnt[u] = 0;
  }
  if (threadIdx.x < kBoxesPerCta) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  for (int p = threadIdx.x; p < n; p += 256) {
    const float x = xyz[(size_t)p * 3], y = xyz[(size_t)p * 3 + 1], z = xyz[(size_t)p * 3 + 2];
#pragma unroll
    for (int u = 0; u < kBoxesPerCta; ++u)
      cnt[u] += (x >= lo[u][0] && x <= hi[u][0] && y >= lo[u][1] && y <= hi[u][1] && z >= lo[u][2] && z <= hi[u][2]);
  }
#pragma unroll
  for (int u = 0; u < kBoxesPerCta; ++u) {
    int c = cnt[u];
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cnt[u], c);
  }
  __syncthreads();
  if (threadIdx.x < kBoxesPerCta && j0 + threadIdx.x < k)
    counts_all[(size_t)scene * k + j0 + threadIdx.x] = s_cnt[threadIdx.x];
}

}  // namespace

int nms3d_dispatch(int b, int k, const float *boxes, const int *valid, double thr, int old_type, int same_cls,
                   int *pick, int *order, cudaStream_t stream) {
  if (k > kNmsMax) return set_error(BQA_ERR_UNSUPPORTED, "nms3d: at most %d boxes per scene (got %d)", kNmsMax, k);
  nms3d_kernel<<<b, kNmsThreads, 0, stream>>>(k, boxes, valid, thr, old_type, same_cls, pick, order);
  count_launch();
  return check_launch("nms3d_kernel");
}

int count_in_boxes_dispatch(int b, int n, int k, const float *xyz, const float *lohi, int *counts,
                            cudaStream_t stream) {
  if (b > 65535) return set_error(BQA_ERR_UNSUPPORTED, "count_in_boxes: batch too large");
  dim3 grid((unsigned)ceil_div(k, kBoxesPerCta), (unsigned)b);
  count_in_boxes_kernel<<<grid, 256, 0, stream>>>(n, k, xyz, lohi, counts);
  count_launch();
  return check_launch("count_in_boxes_kernel");
}

}  // namespace bqa
