// nn_distance.cu -- nearest-neighbour distances between two small point sets, both directions,
// without materialising the (B, N, M, 3) difference tensor (sm_100a).
//
// Replaces /root/reference/utils/nn_distance.py:25-52 (`nn_distance`, used by the VoteNet losses,
// lib/loss_helper.py:66,91,144):  pc_diff = pc1[:, :, None] - pc2[:, None]; pc_dist = sum over the
// last axis of pc_diff**2 | |pc_diff| | huber(pc_diff, delta); dist1, idx1 = min over M; dist2,
// idx2 = min over N.  Same arithmetic as the torch expression evaluates on the GPU, so the result
// is bit-identical: fp32 subtraction, per-component term with separate roundings (no FMA), terms
// added as (t0 + t2) + t1 (the order torch's CUDA reduction uses for a length-3 axis; checked on
// B200 against all three orders), strict `<` scan in ascending index (torch.min returns the first
// minimum).
#include "common.cuh"

namespace bqa {
namespace {

constexpr int kTile = 256;

__device__ __forceinline__ float term(float d, int mode, float delta) {
  if (mode == 0) return __fmul_rn(d, d);                       // pc_diff**2
  const float a = fabsf(d);
  if (mode == 1) return a;                                      // |pc_diff|
  const float q = fminf(a, delta);                              // huber_loss, nn_distance.py:6-23
  const float lin = __fsub_rn(a, q);
  return __fadd_rn(__fmul_rn(0.5f, __fmul_rn(q, q)), __fmul_rn(delta, lin));
}

// queries 0..n-1 are the points of pc1 (searching pc2, writing dist1/idx1), queries n..n+m-1 the
// points of pc2 (searching pc1, writing dist2/idx2); the sign of the difference is pc1 - pc2 in both.
__global__ void __launch_bounds__(kTile)
nn_distance_kernel(int n, int m, int mode, float delta, const float *__restrict__ pc1,
                   const float *__restrict__ pc2, float *__restrict__ dist1, long long *__restrict__ idx1,
                   float *__restrict__ dist2, long long *__restrict__ idx2) {
  __shared__ float tile[kTile * 3];
  const int scene = blockIdx.y;
  const int q = blockIdx.x * kTile + threadIdx.x;
  const bool first = (long long)blockIdx.x * kTile < n;          // a CTA never straddles the two roles:
  const int role_q = first ? q : q - ((n + kTile - 1) / kTile) * kTile;   // grid.x = ceil(n/T) + ceil(m/T)
  const int nq = first ? n : m, no = first ? m : n;
  const float *mine = (first ? pc1 : pc2) + (size_t)scene * nq * 3;
  const float *other = (first ? pc2 : pc1) + (size_t)scene * no * 3;
  const bool live = role_q < nq;
  float x = 0.f, y = 0.f, z = 0.f;
  if (live) { x = mine[role_q * 3]; y = mine[role_q * 3 + 1]; z = mine[role_q * 3 + 2]; }
  float best = INFINITY;
  int bi = 0;
  for (int t0 = 0; t0 < no; t0 += kTile) {
    const int cnt = min(kTile, no - t0);
    __syncthreads();
    for (int e = threadIdx.x; e < cnt * 3; e += kTile) tile[e] = other[(size_t)t0 * 3 + e];
    __syncthreads();
    if (live) {
      for (int k = 0; k < cnt; ++k) {
        const float ox = tile[k * 3], oy = tile[k * 3 + 1], oz = tile[k * 3 + 2];
        const float dx = first ? __fsub_rn(x, ox) : __fsub_rn(ox, x);
        const float dy = first ? __fsub_rn(y, oy) : __fsub_rn(oy, y);
        const float dz = first ? __fsub_rn(z, oz) : __fsub_rn(oz, z);
        // torch.sum over a length-3 innermost axis on CUDA splits the axis over 2 threads
        // (elements 0 and 2 | element 1) and combines: (t0 + t2) + t1 -- measured, not (t0 + t1) + t2
        const float d = __fadd_rn(__fadd_rn(term(dx, mode, delta), term(dz, mode, delta)), term(dy, mode, delta));
        if (d < best || (t0 + k == 0)) { best = d; bi = t0 + k; }
      }
    }
  }
  if (live) {
    if (first) { dist1[(size_t)scene * n + role_q] = best; idx1[(size_t)scene * n + role_q] = bi; }
    else { dist2[(size_t)scene * m + role_q] = best; idx2[(size_t)scene * m + role_q] = bi; }
  }
}

}  // namespace

int nn_distance_dispatch(int b, int n, int m, int mode, float delta, const float *pc1, const float *pc2,
                         float *dist1, long long *idx1, float *dist2, long long *idx2, cudaStream_t stream) {
  if (b > 65535) return set_error(BQA_ERR_UNSUPPORTED, "nn_distance: batch %d > 65535 (split the call)", b);
  dim3 grid((unsigned)(ceil_div(n, kTile) + ceil_div(m, kTile)), (unsigned)b);
  nn_distance_kernel<<<grid, kTile, 0, stream>>>(n, m, mode, delta, pc1, pc2, dist1, idx1, dist2, idx2);
  count_launch();
  return check_launch("nn_distance_kernel");
}

}  // namespace bqa
