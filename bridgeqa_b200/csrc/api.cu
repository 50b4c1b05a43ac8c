// api.cu -- the extern "C" surface declared in include/bqa_pointnet2.h.
// Argument checks mirror the reference wrappers' intent (lib/pointnet2/_ext_src/src/
// *.cpp) but report through a status code + bqa_last_error() instead of AT_ASSERT or
// exit(-1) (cuda_utils.h:30-39).
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace bqa {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(BQA_ERR_CUDA, "%s launch failed: %s", what, cudaGetErrorString(e));
  return BQA_OK;
}

int ref_opt_n_threads(int work_size) {
  // cuda_utils.h:15-19, same expression so the same host libm decides the power of two
  const int pow_2 = (int)(std::log(static_cast<double>(work_size)) / std::log(2.0));
  int v = 1 << pow_2;
  if (v > 512) v = 512;
  if (v < 1) v = 1;
  return v;
}

// dispatchers implemented in the kernel translation units
int fps_dispatch(int b, int n, int m, int j_begin, int j_end, const float *xyz, int *idxs,
                 float *new_xyz, float *scratch, bool exclusive, const int *run_flags,
                 cudaStream_t stream);
int nn_distance_dispatch(int b, int n, int m, int mode, float delta, const float *pc1, const float *pc2,
                         float *dist1, long long *idx1, float *dist2, long long *idx2, cudaStream_t stream);
int nms3d_dispatch(int b, int k, const float *boxes, const int *valid, double thr, int old_type, int same_cls,
                   int *pick, int *order, cudaStream_t stream);
int count_in_boxes_dispatch(int b, int n, int k, const float *xyz, const float *lohi, int *counts,
                            cudaStream_t stream);
bool bn_relu_max_supported(int ns);
int bn_finalize_shifted_dispatch(int c, double count, const double *sums, const float *shift, float eps,
                                 float momentum, float *mean, float *invstd, float *running_mean,
                                 float *running_var, cudaStream_t stream);
bool conv1x1_tf32_supported(int b, int cin, int cout, long long p, int ldw);
int conv1x1_tf32_forward(int b, int cin, int cout, int p, const float *x, const float *w, int ldw, float *y,
                         const float *shift, double *sums, cudaStream_t stream);
int conv1x1_tf32_wgrad(int b, int cin, int cout, int p, const float *x, const float *dy, float *dw,
                       cudaStream_t stream);
int bn_stats_dispatch(int b, int c, long long l, const float *y, double *sums, float eps, float momentum,
                      float *mean, float *invstd, float *running_mean, float *running_var,
                      cudaStream_t stream);
int bn_relu_apply_dispatch(int b, int c, long long l, const float *y, const float *mean, const float *invstd,
                           const float *gamma, const float *beta, float *x, cudaStream_t stream);
int bn_relu_max_dispatch(int b, int c, long long np, int ns, const float *y, const float *mean,
                         const float *invstd, const float *gamma, const float *beta, float *out, int *argmax,
                         cudaStream_t stream);
int bn_relu_backward_dispatch(int b, int c, long long l, const float *dx, const float *y, const float *mean,
                              const float *invstd, const float *gamma, const float *beta, double *sums,
                              float *dy, float *dgamma, float *dbeta, cudaStream_t stream);
int bn_relu_max_backward_dispatch(int b, int c, long long np, int ns, const float *dout, const int *argmax,
                                  const float *y, const float *mean, const float *invstd, const float *gamma,
                                  const float *beta, double *sums, float *dy, float *dgamma, float *dbeta,
                                  cudaStream_t stream);
bool fps_sorted_supported(int n, int m);
int fps_sorted_dispatch(int b, int n, int m, const float *xyz, const void *grid, int *idxs,
                        float *new_xyz, int lean, cudaStream_t stream);
bool fps_prefix_check_supported(int n, int m);
int fps_prefix_check_dispatch(int b, int n, int m, const float *xyz, float *v_scratch, int *run_flags,
                              cudaStream_t stream);
int fps_identity_fill_dispatch(int b, int n, int m, const float *xyz, const int *run_flags, int *idxs,
                               float *new_xyz, cudaStream_t stream);
long long fps_scratch_bytes(int b, int n);
int ball_query_dispatch(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                        const float *xyz, int *idx, void *workspace, cudaStream_t stream,
                        int q_stride, int q_offset);
long long ball_query_workspace_bytes(int b, int n, int m, int nsample);
long long ball_query_grid_bytes(int b, int n);
int ball_query_grid_build(int b, int n, float radius, const float *xyz, void *grid, cudaStream_t stream);
int ball_query_grid_search(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                           const float *xyz, int *idx, const void *grid, cudaStream_t stream,
                           int q_stride, int q_offset);
int gather_rows_dispatch(int b, int c, int n, long long e_total, const float *points, const int *idx,
                         float *out, cudaStream_t stream);
int scatter_add_rows_dispatch(int b, int c, int n, long long e_total, const float *grad_out,
                              const int *idx, float *grad_points, cudaStream_t stream);
int transpose_cn_dispatch(int b, int c, int n, const float *in, float *out, cudaStream_t stream);
long long scatter_add_workspace_bytes(int b, int c, int n);
int scatter_add_rows_ws_dispatch(int b, int c, int n, long long e_total, const float *grad_out, const int *idx,
                                 float *grad_points, float *acc, cudaStream_t stream);
int group_concat_pm_dispatch(int b, int n, int c, int feat_stride, long long e_total, int nsample, float radius,
                             int normalize, const float *xyz, const float *new_xyz, const float *feat,
                             const int *idx, float *out, cudaStream_t stream);
int three_nn_dispatch(int b, int n, int m, const float *unknown, const float *known, float *dist2,
                      int *idx, cudaStream_t stream);
int three_interpolate_dispatch(int b, int c, int m, int n, const float *points, const int *idx,
                               const float *weight, float *out, cudaStream_t stream);
int three_interpolate_grad_dispatch(int b, int c, int n, int m, const float *grad_out, const int *idx,
                                    const float *weight, float *grad_points, cudaStream_t stream);

int sa_supported(int nsample, int npoint, int c, int c1, int c2, int c3);
int pack_weight_dispatch(int c_out, int c_in, int kpad, int xyz_first, int fp16, const float *w,
                         void *packed, cudaStream_t stream);
int sa_forward_dispatch(int b, int n, int npoint, int nsample, int c, const float *xyz,
                        const float *new_xyz, const float *feat_pm, int feat_stride, const int *idx, float radius,
                        int normalize_xyz, int c1, int c2, int c3, const void *w1p, const float *b1,
                        const void *w2p, const float *b2, const void *w3p, const float *b3,
                        float *out_cm, float *out_pm, int fp16, cudaStream_t stream, int npoint_total,
                        int j_offset);

int sa_v2_supported(int nsample, int npoint, int c, int c1, int c2, int c3);
int pack_weight_v2_dispatch(int c_out, int c_in, int kpad, int mode, int fp16, const float *w, const float *bias,
                            void *packed, cudaStream_t stream);
int to_point_major_16_dispatch(int b, int c, int n, int stride, int fp16, const float *in, void *out,
                               cudaStream_t stream);
int rows_to_16_dispatch(long long rows, int c, int row_stride, int first, int stride, int fp16, const float *in,
                        void *out, cudaStream_t stream);
int sa_v2_forward_dispatch(int b, int n, int npoint, int nsample, int c, const float *xyz, const float *new_xyz,
                           const void *feat16, int stride16, const int *idx, float radius, int normalize_xyz,
                           int c1, int c2, int c3, const void *w1p, const void *w2p, const void *w3p,
                           const float *b3, float *out_cm, float *out_pm, void *out_pm16, int fp16,
                           cudaStream_t stream);

int fp_supported(int n, int m, int c_known, int c_skip, int c1, int c2);
int fp_forward_dispatch(int b, int n, int m, int c_known, int c_skip, const float *unknown,
                        const float *known, const float *known_feat, int known_stride,
                        const float *skip_feat, int skip_stride, int c1, int c2, const void *w,
                        const float *b1, const float *b2, float *out_cm, float *out_pm, int fp16,
                        cudaStream_t stream, const void *skip16, int skip16_stride);

}  // namespace bqa

using namespace bqa;

#define NONNEG(x) BQA_REQUIRE((x) >= 0, "%s: %s must be >= 0 (got %d)", __func__, #x, (int)(x))
#define PTR(p) BQA_REQUIRE((p) != nullptr, "%s: %s is NULL", __func__, #p)

extern "C" {

int bqa_abi_version(void) { return BQA_ABI_VERSION; }
const char *bqa_last_error(void) { return g_err; }
long long bqa_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

long long bqa_fps_scratch_bytes(int b, int n) { return fps_scratch_bytes(b, n); }

int bqa_furthest_point_sampling(int b, int n, int m, const float *xyz, int *idxs, float *new_xyz,
                                float *scratch, void *stream) {
  NONNEG(b); NONNEG(n); NONNEG(m);
  if (b == 0 || m == 0) return BQA_OK;
  BQA_REQUIRE(n > 0, "%s: n must be > 0 when m > 0", __func__);
  PTR(xyz); PTR(idxs);
  return fps_dispatch(b, n, m, 1, m, xyz, idxs, new_xyz, scratch, false, nullptr, (cudaStream_t)stream);
}

int bqa_fps_grid_supported(int n, int m) { return fps_sorted_supported(n, m) ? 1 : 0; }

int bqa_furthest_point_sampling_grid(int b, int n, int m, const float *xyz, const void *grid, int *idxs,
                                     float *new_xyz, void *stream) {
  NONNEG(b); NONNEG(n); NONNEG(m);
  if (b == 0 || m == 0) return BQA_OK;
  BQA_REQUIRE(fps_sorted_supported(n, m), "%s: n=%d is outside what the sorted kernel takes "
              "(see bqa_fps_grid_supported)", __func__, n);
  PTR(xyz); PTR(grid); PTR(idxs);
  return fps_sorted_dispatch(b, n, m, xyz, grid, idxs, new_xyz, 0, (cudaStream_t)stream);
}

int bqa_furthest_point_sampling_grid_lean(int b, int n, int m, const float *xyz, const void *grid, int *idxs,
                                          float *new_xyz, void *stream) {
  NONNEG(b); NONNEG(n); NONNEG(m);
  if (b == 0 || m == 0) return BQA_OK;
  BQA_REQUIRE(fps_sorted_supported(n, m), "%s: n=%d is outside what the sorted kernel takes "
              "(see bqa_fps_grid_supported)", __func__, n);
  PTR(xyz); PTR(grid); PTR(idxs);
  return fps_sorted_dispatch(b, n, m, xyz, grid, idxs, new_xyz, 1, (cudaStream_t)stream);
}

int bqa_fps_prefix_check(int b, int n, int m, const float *xyz, float *v_scratch, int *run_flags,
                         void *stream) {
  NONNEG(b); NONNEG(n); NONNEG(m);
  if (b == 0) return BQA_OK;
  PTR(run_flags);
  BQA_REQUIRE(fps_prefix_check_supported(n, m), "%s: need 1 <= m <= min(n, 8192), got n=%d m=%d",
              __func__, n, m);
  PTR(xyz); PTR(v_scratch);
  return fps_prefix_check_dispatch(b, n, m, xyz, v_scratch, run_flags, (cudaStream_t)stream);
}

int bqa_furthest_point_sampling_cond(int b, int n, int m, const float *xyz, const int *run_flags,
                                     int *idxs, float *new_xyz, float *scratch, void *stream) {
  NONNEG(b); NONNEG(n); NONNEG(m);
  if (b == 0 || m == 0) return BQA_OK;
  BQA_REQUIRE(n > 0 && m <= n, "%s: need 0 < m <= n", __func__);
  PTR(xyz); PTR(idxs); PTR(run_flags);
  if (int rc = fps_identity_fill_dispatch(b, n, m, xyz, run_flags, idxs, new_xyz, (cudaStream_t)stream))
    return rc;
  return fps_dispatch(b, n, m, 1, m, xyz, idxs, new_xyz, scratch, false, run_flags, (cudaStream_t)stream);
}

int bqa_furthest_point_sampling_slice(int b, int n, int m, int j_begin, int j_end, const float *xyz,
                                      int *idxs, float *new_xyz, float *state, int exclusive,
                                      void *stream) {
  NONNEG(b); NONNEG(n); NONNEG(m);
  BQA_REQUIRE(j_begin >= 1 && j_begin <= j_end && j_end <= m, "%s: need 1 <= j_begin <= j_end <= m", __func__);
  if (b == 0 || m == 0) return BQA_OK;
  BQA_REQUIRE(n > 0, "%s: n must be > 0 when m > 0", __func__);
  PTR(xyz); PTR(idxs);
  return fps_dispatch(b, n, m, j_begin, j_end, xyz, idxs, new_xyz, state, exclusive != 0, nullptr,
                      (cudaStream_t)stream);
}

int bqa_gather_points(int b, int c, int n, int m, const float *points, const int *idx, float *out,
                      void *stream) {
  NONNEG(b); NONNEG(c); NONNEG(n); NONNEG(m);
  if ((long long)b * c * m == 0) return BQA_OK;
  PTR(points); PTR(idx); PTR(out);
  return gather_rows_dispatch(b, c, n, m, points, idx, out, (cudaStream_t)stream);
}

int bqa_gather_points_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                           float *grad_points, void *stream) {
  NONNEG(b); NONNEG(c); NONNEG(n); NONNEG(m);
  if ((long long)b * c * n == 0) return BQA_OK;
  PTR(grad_points);
  if (m > 0) { PTR(grad_out); PTR(idx); }
  return scatter_add_rows_dispatch(b, c, n, m, grad_out, idx, grad_points, (cudaStream_t)stream);
}

long long bqa_ball_query_workspace_bytes(int b, int n, int m, int nsample) {
  if (b <= 0 || n <= 0 || m <= 0 || nsample <= 0) return 0;
  return ball_query_workspace_bytes(b, n, m, nsample);
}

int bqa_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                   const float *xyz, int *idx, void *workspace, void *stream) {
  NONNEG(b); NONNEG(n); NONNEG(m); NONNEG(nsample);
  if ((long long)b * m * nsample == 0) return BQA_OK;
  PTR(new_xyz); PTR(idx);
  if (n > 0) PTR(xyz);
  return ball_query_dispatch(b, n, m, radius, nsample, new_xyz, xyz, idx, workspace,
                             (cudaStream_t)stream, m, 0);
}

int bqa_ball_query_slice(int b, int n, int m_total, int j_begin, int j_count, float radius, int nsample,
                         const float *new_xyz, const float *xyz, int *idx, void *workspace,
                         void *stream) {
  NONNEG(b); NONNEG(n); NONNEG(m_total); NONNEG(nsample);
  BQA_REQUIRE(j_begin >= 0 && j_count >= 0 && j_begin + j_count <= m_total,
              "%s: slice [%d, %d) outside [0, %d)", __func__, j_begin, j_begin + j_count, m_total);
  if ((long long)b * j_count * nsample == 0) return BQA_OK;
  PTR(new_xyz); PTR(idx);
  if (n > 0) PTR(xyz);
  return ball_query_dispatch(b, n, j_count, radius, nsample, new_xyz, xyz, idx, workspace,
                             (cudaStream_t)stream, m_total, j_begin);
}

long long bqa_ball_query_grid_bytes(int b, int n) {
  if (b <= 0 || n <= 0) return 0;
  return ball_query_grid_bytes(b, n);
}

int bqa_ball_query_grid_build(int b, int n, float radius, const float *xyz, void *grid, void *stream) {
  NONNEG(b); NONNEG(n);
  if ((long long)b * n == 0) return BQA_OK;
  PTR(xyz); PTR(grid);
  return ball_query_grid_build(b, n, radius, xyz, grid, (cudaStream_t)stream);
}

int bqa_ball_query_grid_search(int b, int n, int m_total, int j_begin, int j_count, float radius,
                               int nsample, const float *new_xyz, const float *xyz, int *idx,
                               const void *grid, void *stream) {
  NONNEG(b); NONNEG(n); NONNEG(m_total); NONNEG(nsample);
  BQA_REQUIRE(j_begin >= 0 && j_count >= 0 && j_begin + j_count <= m_total,
              "%s: slice [%d, %d) outside [0, %d)", __func__, j_begin, j_begin + j_count, m_total);
  if ((long long)b * j_count * nsample == 0) return BQA_OK;
  PTR(new_xyz); PTR(idx);
  BQA_REQUIRE(n > 0, "%s: a grid needs at least one point", __func__);
  PTR(xyz); PTR(grid);
  return ball_query_grid_search(b, n, j_count, radius, nsample, new_xyz, xyz, idx, grid,
                                (cudaStream_t)stream, m_total, j_begin);
}

int bqa_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                     const int *idx, float *out, void *stream) {
  NONNEG(b); NONNEG(c); NONNEG(n); NONNEG(npoints); NONNEG(nsample);
  const long long e = (long long)npoints * nsample;
  if ((long long)b * c * e == 0) return BQA_OK;
  PTR(points); PTR(idx); PTR(out);
  return gather_rows_dispatch(b, c, n, e, points, idx, out, (cudaStream_t)stream);
}

long long bqa_group_points_grad_workspace_bytes(int b, int c, int n) {
  if (b <= 0 || c <= 0 || n <= 0) return 0;
  return scatter_add_workspace_bytes(b, c, n);
}

int bqa_group_points_grad_ws(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                             const int *idx, float *grad_points, void *workspace, void *stream) {
  NONNEG(b); NONNEG(c); NONNEG(n); NONNEG(npoints); NONNEG(nsample);
  if ((long long)b * c * n == 0) return BQA_OK;
  PTR(grad_points); PTR(workspace);
  const long long e = (long long)npoints * nsample;
  if (e > 0) { PTR(grad_out); PTR(idx); }
  BQA_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "%s: workspace must be 16-byte aligned", __func__);
  return scatter_add_rows_ws_dispatch(b, c, n, e, grad_out, idx, grad_points, (float *)workspace,
                                      (cudaStream_t)stream);
}

int bqa_group_concat_point_major(int b, int n, int c, int feat_stride, int npoint, int nsample,
                                 const float *xyz, const float *new_xyz, const float *feat_pm,
                                 const int *idx, float radius, int normalize_xyz, float *out, void *stream) {
  NONNEG(b); NONNEG(n); NONNEG(c); NONNEG(npoint); NONNEG(nsample);
  const long long e = (long long)npoint * nsample;
  if ((long long)b * e == 0) return BQA_OK;
  PTR(xyz); PTR(new_xyz); PTR(idx); PTR(out);
  if (c > 0) PTR(feat_pm);
  BQA_REQUIRE(feat_stride >= c, "%s: feat_stride=%d < c=%d", __func__, feat_stride, c);
  BQA_REQUIRE(!normalize_xyz || radius > 0.f, "%s: radius must be > 0", __func__);
  return group_concat_pm_dispatch(b, n, c, feat_stride, e, nsample, radius, normalize_xyz, xyz, new_xyz,
                                  feat_pm, idx, out, (cudaStream_t)stream);
}

int bqa_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                          const int *idx, float *grad_points, void *stream) {
  NONNEG(b); NONNEG(c); NONNEG(n); NONNEG(npoints); NONNEG(nsample);
  if ((long long)b * c * n == 0) return BQA_OK;
  PTR(grad_points);
  const long long e = (long long)npoints * nsample;
  if (e > 0) { PTR(grad_out); PTR(idx); }
  return scatter_add_rows_dispatch(b, c, n, e, grad_out, idx, grad_points, (cudaStream_t)stream);
}

int bqa_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2,
                 int *idx, void *stream) {
  NONNEG(b); NONNEG(n); NONNEG(m);
  if ((long long)b * n == 0) return BQA_OK;
  PTR(unknown); PTR(dist2); PTR(idx);
  if (m > 0) PTR(known);
  return three_nn_dispatch(b, n, m, unknown, known, dist2, idx, (cudaStream_t)stream);
}

int bqa_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx,
                          const float *weight, float *out, void *stream) {
  NONNEG(b); NONNEG(c); NONNEG(m); NONNEG(n);
  if ((long long)b * c * n == 0) return BQA_OK;
  PTR(points); PTR(idx); PTR(weight); PTR(out);
  return three_interpolate_dispatch(b, c, m, n, points, idx, weight, out, (cudaStream_t)stream);
}

int bqa_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                               const float *weight, float *grad_points, void *stream) {
  NONNEG(b); NONNEG(c); NONNEG(m); NONNEG(n);
  if ((long long)b * c * m == 0) return BQA_OK;
  PTR(grad_points);
  if (n > 0) { PTR(grad_out); PTR(idx); PTR(weight); }
  return three_interpolate_grad_dispatch(b, c, n, m, grad_out, idx, weight, grad_points,
                                         (cudaStream_t)stream);
}

int bqa_transpose_to_point_major(int b, int c, int n, const float *in, float *out, void *stream) {
  NONNEG(b); NONNEG(c); NONNEG(n);
  if ((long long)b * c * n == 0) return BQA_OK;
  PTR(in); PTR(out);
  return transpose_cn_dispatch(b, c, n, in, out, (cudaStream_t)stream);
}

int bqa_pack_weight_16(int c_out, int c_in, int k_pad, int xyz_first, int precision, const float *w,
                       void *packed, void *stream) {
  BQA_REQUIRE(precision == 0 || precision == 1, "%s: precision must be 0 (bf16) or 1 (fp16)", __func__);
  BQA_REQUIRE(c_out > 0 && c_in > 0, "%s: empty weight", __func__);
  BQA_REQUIRE(k_pad >= c_in && k_pad % 16 == 0, "%s: k_pad=%d must be a multiple of 16 >= c_in=%d",
              __func__, k_pad, c_in);
  BQA_REQUIRE(!xyz_first || c_in >= 3, "%s: xyz_first needs c_in >= 3", __func__);
  PTR(w); PTR(packed);
  return pack_weight_dispatch(c_out, c_in, k_pad, xyz_first, precision, w, packed, (cudaStream_t)stream);
}

int bqa_sa_mlp_max_supported(int nsample, int npoint, int c, int c1, int c2, int c3) {
  return sa_supported(nsample, npoint, c, c1, c2, c3);
}

int bqa_sa_mlp_max_forward(int b, int n, int npoint, int nsample, int c, const float *xyz,
                           const float *new_xyz, const float *feat_pm, int feat_stride, const int *idx, float radius,
                           int normalize_xyz, int c1, int c2, int c3, const void *w1p, const float *b1,
                           const void *w2p, const float *b2, const void *w3p, const float *b3,
                           float *out_cm, float *out_pm, int precision, void *stream) {
  NONNEG(b); NONNEG(n); NONNEG(npoint); NONNEG(nsample); NONNEG(c);
  BQA_REQUIRE(precision == 0 || precision == 1, "%s: precision must be 0 (bf16) or 1 (fp16)", __func__);
  if ((long long)b * npoint == 0) return BQA_OK;
  PTR(xyz); PTR(new_xyz); PTR(idx); PTR(w1p); PTR(b1); PTR(w2p); PTR(b2); PTR(w3p); PTR(b3); PTR(out_cm);
  BQA_REQUIRE((c == 0) == (feat_pm == nullptr), "%s: feat_pm must be NULL iff c == 0", __func__);
  BQA_REQUIRE(c == 0 || feat_stride >= c, "%s: feat_stride=%d < c=%d", __func__, feat_stride, c);
  BQA_REQUIRE(!normalize_xyz || radius > 0.f, "%s: radius must be > 0", __func__);
  return sa_forward_dispatch(b, n, npoint, nsample, c, xyz, new_xyz, feat_pm, feat_stride, idx, radius,
                             normalize_xyz, c1, c2, c3, w1p, b1, w2p, b2, w3p, b3, out_cm, out_pm,
                             precision, (cudaStream_t)stream, npoint, 0);
}

int bqa_pack_weight_16_v2(int c_out, int c_in, int k_pad, int mode, int precision, const float *w,
                          const float *bias, void *packed, void *stream) {
  BQA_REQUIRE(precision == 0 || precision == 1, "%s: precision must be 0 (bf16) or 1 (fp16)", __func__);
  BQA_REQUIRE(c_out > 0 && c_in > 0, "%s: empty weight", __func__);
  BQA_REQUIRE(mode >= 0 && mode <= 2, "%s: mode must be 0, 1 or 2", __func__);
  BQA_REQUIRE(k_pad % 16 == 0, "%s: k_pad=%d must be a multiple of 16", __func__, k_pad);
  if (mode == 0) BQA_REQUIRE(k_pad >= c_in, "%s: k_pad=%d < c_in=%d", __func__, k_pad, c_in);
  if (mode == 1) BQA_REQUIRE(c_in >= 3 && k_pad == (c_in - 3 + 5 + 15) / 16 * 16,
                             "%s: mode 1 needs c_in >= 3 and k_pad = roundup16(c_in + 2)", __func__);
  if (mode == 2) BQA_REQUIRE(k_pad == c_in + 16, "%s: mode 2 needs k_pad = c_in + 16", __func__);
  PTR(w); PTR(packed);
  return pack_weight_v2_dispatch(c_out, c_in, k_pad, mode, precision, w, bias, packed, (cudaStream_t)stream);
}

int bqa_to_point_major_16(int b, int c, int n, int stride, int precision, const float *in, void *out,
                          void *stream) {
  NONNEG(b); NONNEG(c); NONNEG(n);
  BQA_REQUIRE(precision == 0 || precision == 1, "%s: precision must be 0 (bf16) or 1 (fp16)", __func__);
  BQA_REQUIRE(stride >= c && stride % 8 == 0, "%s: stride=%d must be a multiple of 8 >= c=%d", __func__, stride, c);
  if ((long long)b * c * n == 0) return BQA_OK;
  BQA_REQUIRE(b <= 65535, "%s: batch too large", __func__);
  PTR(in); PTR(out);
  return to_point_major_16_dispatch(b, c, n, stride, precision, in, out, (cudaStream_t)stream);
}

int bqa_rows_to_16(long long rows, int c, int row_stride, int first, int stride, int precision,
                   const float *in, void *out, void *stream) {
  BQA_REQUIRE(rows >= 0 && c >= 0 && first >= 0, "%s: negative size", __func__);
  BQA_REQUIRE(precision == 0 || precision == 1, "%s: precision must be 0 (bf16) or 1 (fp16)", __func__);
  BQA_REQUIRE(row_stride >= first + c, "%s: row_stride=%d < first + c = %d", __func__, row_stride, first + c);
  BQA_REQUIRE(stride >= c && stride % 8 == 0 && stride > 0, "%s: stride=%d must be a positive multiple of 8 >= c=%d",
              __func__, stride, c);
  if (rows == 0) return BQA_OK;
  PTR(in); PTR(out);
  return rows_to_16_dispatch(rows, c, row_stride, first, stride, precision, in, out, (cudaStream_t)stream);
}

int bqa_sa_mlp_max_v2_supported(int nsample, int npoint, int c, int c1, int c2, int c3) {
  return sa_v2_supported(nsample, npoint, c, c1, c2, c3);
}

int bqa_sa_mlp_max_forward_v2(int b, int n, int npoint, int nsample, int c, const float *xyz,
                              const float *new_xyz, const void *feat16, int stride16, const int *idx,
                              float radius, int normalize_xyz, int c1, int c2, int c3, const void *w1p,
                              const void *w2p, const void *w3p, const float *b3, float *out_cm,
                              float *out_pm, void *out_pm16, int precision, void *stream) {
  NONNEG(b); NONNEG(n); NONNEG(npoint); NONNEG(nsample); NONNEG(c);
  BQA_REQUIRE(precision == 0 || precision == 1, "%s: precision must be 0 (bf16) or 1 (fp16)", __func__);
  if ((long long)b * npoint == 0) return BQA_OK;
  PTR(xyz); PTR(new_xyz); PTR(idx); PTR(w1p); PTR(w2p); PTR(w3p); PTR(b3); PTR(out_cm);
  BQA_REQUIRE((c == 0) == (feat16 == nullptr), "%s: feat16 must be NULL iff c == 0", __func__);
  BQA_REQUIRE(c == 0 || (stride16 % 8 == 0 && stride16 >= (c + 7) / 8 * 8),
              "%s: stride16=%d must be a multiple of 8 >= roundup8(c=%d)", __func__, stride16, c);
  BQA_REQUIRE(c == 0 || (reinterpret_cast<uintptr_t>(feat16) & 15) == 0, "%s: feat16 must be 16-byte aligned", __func__);
  BQA_REQUIRE(!normalize_xyz || radius > 0.f, "%s: radius must be > 0", __func__);
  return sa_v2_forward_dispatch(b, n, npoint, nsample, c, xyz, new_xyz, feat16, stride16, idx, radius,
                                normalize_xyz, c1, c2, c3, w1p, w2p, w3p, b3, out_cm, out_pm, out_pm16,
                                precision, (cudaStream_t)stream);
}

int bqa_fp_mlp_supported(int n, int m, int c_known, int c_skip, int c1, int c2) {
  return fp_supported(n, m, c_known, c_skip, c1, c2);
}

int bqa_fp_mlp_forward(int b, int n, int m, int c_known, int c_skip, const float *unknown,
                       const float *known, const float *known_feat, int known_stride,
                       const float *skip_feat, int skip_stride, int c1, int c2, const void *w,
                       const float *b1, const float *b2, float *out_cm, float *out_pm, int precision,
                       void *stream, const void *skip16, int skip16_stride) {
  NONNEG(b); NONNEG(n); NONNEG(m);
  if ((long long)b * n == 0) return BQA_OK;
  BQA_REQUIRE(precision == 0 || precision == 1, "%s: precision must be 0 (bf16) or 1 (fp16)", __func__);
  PTR(unknown); PTR(known); PTR(known_feat); PTR(w); PTR(b1); PTR(b2); PTR(out_cm);
  BQA_REQUIRE(skip_feat != nullptr || skip16 != nullptr, "%s: skip_feat and skip16 are both NULL", __func__);
  BQA_REQUIRE(known_stride >= c_known && (skip_feat == nullptr || skip_stride >= c_skip),
              "%s: row strides too small", __func__);
  return fp_forward_dispatch(b, n, m, c_known, c_skip, unknown, known, known_feat, known_stride,
                             skip_feat, skip_stride, c1, c2, w, b1, b2, out_cm, out_pm, precision,
                             (cudaStream_t)stream, skip16, skip16_stride);
}

int bqa_nn_distance(int b, int n, int m, const float *pc1, const float *pc2, int mode, float delta,
                    float *dist1, long long *idx1, float *dist2, long long *idx2, void *stream) {
  NONNEG(b); NONNEG(n); NONNEG(m);
  BQA_REQUIRE(mode >= 0 && mode <= 2, "%s: mode must be 0 (squared L2), 1 (L1) or 2 (Huber)", __func__);
  if (b == 0 || (n == 0 && m == 0)) return BQA_OK;
  BQA_REQUIRE(n > 0 && m > 0, "%s: both point sets must be non-empty (torch.min over an empty axis raises)", __func__);
  PTR(pc1); PTR(pc2); PTR(dist1); PTR(idx1); PTR(dist2); PTR(idx2);
  return nn_distance_dispatch(b, n, m, mode, delta, pc1, pc2, dist1, idx1, dist2, idx2, (cudaStream_t)stream);
}

int bqa_nms3d(int b, int k, const float *boxes, const int *valid, double iou_threshold, int old_type,
              int same_class, int *pick_mask, int *pick_order, void *stream) {
  NONNEG(b); NONNEG(k);
  if ((long long)b * k == 0) return BQA_OK;
  PTR(boxes); PTR(pick_mask);
  return nms3d_dispatch(b, k, boxes, valid, iou_threshold, old_type != 0, same_class != 0, pick_mask, pick_order,
                        (cudaStream_t)stream);
}

int bqa_count_points_in_boxes(int b, int n, int k, const float *xyz, const float *box_lo_hi, int *counts,
                              void *stream) {
  NONNEG(b); NONNEG(n); NONNEG(k);
  if ((long long)b * k == 0) return BQA_OK;
  PTR(box_lo_hi); PTR(counts);
  if (n > 0) PTR(xyz);
  return count_in_boxes_dispatch(b, n, k, xyz, box_lo_hi, counts, (cudaStream_t)stream);
}

#define ALIGNED16(p) BQA_REQUIRE((reinterpret_cast<uintptr_t>(p) & 15) == 0, "%s: %s must be 16-byte aligned", __func__, #p)

int bqa_bn_relu_max_supported(int nsample) { return bn_relu_max_supported(nsample) ? 1 : 0; }

int bqa_bn_train_stats(int b, int c, long long l, const float *y, double *sums_scratch, float eps,
                       float momentum, float *mean, float *invstd, float *running_mean,
                       float *running_var, void *stream) {
  NONNEG(b); NONNEG(c);
  BQA_REQUIRE(l >= 0, "%s: l must be >= 0", __func__);
  BQA_REQUIRE((long long)b * l > 0 || c == 0, "%s: batch statistics of an empty tensor", __func__);
  if (c == 0) return BQA_OK;
  PTR(y); PTR(sums_scratch); PTR(mean); PTR(invstd);
  ALIGNED16(y);
  return bn_stats_dispatch(b, c, l, y, sums_scratch, eps, momentum, mean, invstd, running_mean, running_var,
                           (cudaStream_t)stream);
}

int bqa_bn_finalize_shifted(int c, double count, const double *sums, const float *shift, float eps,
                            float momentum, float *mean, float *invstd, float *running_mean,
                            float *running_var, void *stream) {
  NONNEG(c);
  if (c == 0) return BQA_OK;
  BQA_REQUIRE(count > 0.0, "%s: batch statistics of an empty tensor", __func__);
  PTR(sums); PTR(mean); PTR(invstd);
  return bn_finalize_shifted_dispatch(c, count, sums, shift, eps, momentum, mean, invstd, running_mean,
                                      running_var, (cudaStream_t)stream);
}

int bqa_conv1x1_tf32_supported(int b, int cin, int cout, long long p, int ldw) {
  return conv1x1_tf32_supported(b, cin, cout, p, ldw) ? 1 : 0;
}

int bqa_conv1x1_tf32_forward(int b, int cin, int cout, long long p, const float *x, const float *w, int ldw,
                             float *y, const float *shift, double *sums, void *stream) {
  NONNEG(b); NONNEG(cin); NONNEG(cout);
  BQA_REQUIRE(p >= 0, "%s: p must be >= 0", __func__);
  if ((long long)b * cout * p == 0) return BQA_OK;
  PTR(x); PTR(w); PTR(y);
  BQA_REQUIRE(conv1x1_tf32_supported(b, cin, cout, p, ldw),
              "%s: needs p %% 4 == 0, ldw %% 4 == 0, ldw >= cin (b=%d cin=%d cout=%d p=%lld ldw=%d)", __func__, b,
              cin, cout, p, ldw);
  return conv1x1_tf32_forward(b, cin, cout, (int)p, x, w, ldw, y, shift, sums, (cudaStream_t)stream);
}

int bqa_conv1x1_tf32_wgrad(int b, int cin, int cout, long long p, const float *x, const float *dy, float *dw,
                           void *stream) {
  NONNEG(b); NONNEG(cin); NONNEG(cout);
  BQA_REQUIRE(p >= 0, "%s: p must be >= 0", __func__);
  if ((long long)cin * cout == 0) return BQA_OK;
  PTR(dw);
  if ((long long)b * p == 0) return BQA_OK;
  PTR(x); PTR(dy);
  BQA_REQUIRE((p % 4) == 0 && cout <= 256, "%s: needs p %% 4 == 0 and cout <= 256 (cout=%d p=%lld)", __func__, cout, p);
  return conv1x1_tf32_wgrad(b, cin, cout, (int)p, x, dy, dw, (cudaStream_t)stream);
}

int bqa_bn_relu_forward(int b, int c, long long l, const float *y, const float *mean, const float *invstd,
                        const float *gamma, const float *beta, float *x, void *stream) {
  NONNEG(b); NONNEG(c);
  if ((long long)b * c * l <= 0) return BQA_OK;
  PTR(y); PTR(mean); PTR(invstd); PTR(gamma); PTR(beta); PTR(x);
  ALIGNED16(y); ALIGNED16(x);
  return bn_relu_apply_dispatch(b, c, l, y, mean, invstd, gamma, beta, x, (cudaStream_t)stream);
}

int bqa_bn_relu_max_forward(int b, int c, int npoint, int nsample, const float *y, const float *mean,
                            const float *invstd, const float *gamma, const float *beta, float *out,
                            int *argmax, void *stream) {
  NONNEG(b); NONNEG(c); NONNEG(npoint);
  BQA_REQUIRE(bn_relu_max_supported(nsample), "%s: nsample=%d must be a power of two in [4, 128]", __func__, nsample);
  if ((long long)b * c * npoint == 0) return BQA_OK;
  PTR(y); PTR(mean); PTR(invstd); PTR(gamma); PTR(beta); PTR(out); PTR(argmax);
  ALIGNED16(y);
  return bn_relu_max_dispatch(b, c, npoint, nsample, y, mean, invstd, gamma, beta, out, argmax,
                              (cudaStream_t)stream);
}

int bqa_bn_relu_backward(int b, int c, long long l, const float *dx, const float *y, const float *mean,
                         const float *invstd, const float *gamma, const float *beta, double *sums_scratch,
                         float *dy, float *dgamma, float *dbeta, void *stream) {
  NONNEG(b); NONNEG(c);
  if (c == 0) return BQA_OK;
  BQA_REQUIRE((long long)b * l > 0, "%s: empty tensor", __func__);
  PTR(dx); PTR(y); PTR(mean); PTR(invstd); PTR(gamma); PTR(beta); PTR(sums_scratch); PTR(dy); PTR(dgamma); PTR(dbeta);
  ALIGNED16(y); ALIGNED16(dx); ALIGNED16(dy);
  return bn_relu_backward_dispatch(b, c, l, dx, y, mean, invstd, gamma, beta, sums_scratch, dy, dgamma, dbeta,
                                   (cudaStream_t)stream);
}

int bqa_bn_relu_max_backward(int b, int c, int npoint, int nsample, const float *dout, const int *argmax,
                             const float *y, const float *mean, const float *invstd, const float *gamma,
                             const float *beta, double *sums_scratch, float *dy, float *dgamma, float *dbeta,
                             void *stream) {
  NONNEG(b); NONNEG(c); NONNEG(npoint);
  BQA_REQUIRE(bn_relu_max_supported(nsample), "%s: nsample=%d must be a power of two in [4, 128]", __func__, nsample);
  if (c == 0) return BQA_OK;
  BQA_REQUIRE((long long)b * npoint > 0, "%s: empty tensor", __func__);
  PTR(dout); PTR(argmax); PTR(y); PTR(mean); PTR(invstd); PTR(gamma); PTR(beta); PTR(sums_scratch); PTR(dy);
  PTR(dgamma); PTR(dbeta);
  ALIGNED16(y); ALIGNED16(dy);
  return bn_relu_max_backward_dispatch(b, c, npoint, nsample, dout, argmax, y, mean, invstd, gamma, beta,
                                       sums_scratch, dy, dgamma, dbeta, (cudaStream_t)stream);
}

}  // extern "C"
