/*
 * bqa_pointnet2.h -- C ABI of libbqa_pointnet2.so, the B200 (sm_100a) replacement for
 * BridgeQA's `pointnet2._ext` pybind module.
 *
 * Every entry point takes raw DEVICE pointers, int sizes and a `cudaStream_t` passed
 * as `void*`; it enqueues work on that stream and returns immediately (no host sync,
 * no allocation).  Return value: 0 = ok, nonzero = BQA_ERR_* (message via
 * bqa_last_error(), thread-local).  The library never calls exit(): the reference's
 * CUDA_CHECK_ERRORS() (lib/pointnet2/_ext_src/include/cuda_utils.h:30-39) does.
 *
 * "replaces" cites the reference interface the entry stands in for, relative to
 * /root/reference/lib/pointnet2/_ext_src/.  Tensor layouts and index dtype (int32)
 * are the reference's.  Output buffers need NOT be pre-zeroed (the reference relies
 * on torch::zeros for ball_query rows and every *_grad; here the library does it).
 */
#ifndef BQA_POINTNET2_H_
#define BQA_POINTNET2_H_

#ifdef __cplusplus
extern "C" {
#endif

#define BQA_OK 0
#define BQA_ERR_INVALID_ARG 1
#define BQA_ERR_CUDA 2
#define BQA_ERR_UNSUPPORTED 3

#define BQA_ABI_VERSION 1

#if defined(__GNUC__)
#define BQA_API __attribute__((visibility("default")))
#else
#define BQA_API
#endif

/* ---- library ---------------------------------------------------------------- */
BQA_API int bqa_abi_version(void);
/* last error message of the calling thread ("" if none) */
BQA_API const char *bqa_last_error(void);
/* number of kernels this library has launched since load (all threads) */
BQA_API long long bqa_launch_count(void);

/* ---- furthest point sampling -------------------------------------------------
 * replaces: furthest_point_sampling(points (B,N,3), nsamples) -> (B,nsamples) int32
 *           src/sampling.cpp:66-87, kernel src/sampling_gpu.cu:69-173.
 * xyz (b,n,3) f32, idxs (b,m) i32.  Bit-exact index contract incl. the 1e-3 norm
 * skip and the 512-slot tree tie-break.  The reference's (b,n) `temp` scratch is not
 * needed: running min-distances live in registers of a thread-block cluster.
 * new_xyz (b,m,3) f32 may be NULL; when given, the sampled coordinates are written
 * too (fused gather_points of src/sampling_gpu.cu:8-20 on xyz).
 * scratch: NULL unless bqa_fps_scratch_bytes(b, n) > 0 (scenes beyond 131072 points,
 * which fall back to a global-memory variant needing the reference's (b,n) temp).  */
BQA_API long long bqa_fps_scratch_bytes(int b, int n);
BQA_API int bqa_furthest_point_sampling(int b, int n, int m, const float *xyz, int *idxs,
                                float *new_xyz, float *scratch, void *stream);

/* Sampling over the ball query's cell grid (bqa_ball_query_grid_build, below): the same result
 * as bqa_furthest_point_sampling, bit for bit, but each warp holds a spatially compact run of the
 * cell-sorted points and skips every iteration whose new sample provably cannot lower any of
 * its min-distances (conservative bounding-box test), so only a few percent of the warps do the
 * update in a typical iteration.  grid: the buffer built for this xyz (any radius).
 * bqa_fps_grid_supported(n, m) != 0: 512 <= n <= 147456.  */
BQA_API int bqa_fps_grid_supported(int n, int m);
BQA_API int bqa_furthest_point_sampling_grid(int b, int n, int m, const float *xyz, const void *grid,
                                             int *idxs, float *new_xyz, void *stream);
/* Same result, bit for bit, from the THROUGHPUT variant of that kernel: only the running
 * min-distances stay in registers, coordinates and tie keys are read from shared memory by the
 * warps that do update, so a scene occupies ceil(n / 13824) SMs (3 for 40k points) instead of 6
 * at ~35 % more latency per iteration (0.81 vs 0.61 us at 16 x 40k on B200).  Meant for several batches in flight, where SM-time and not
 * the latency of one chain bounds the throughput.  n > 110592 runs the kernel above.  */
BQA_API int bqa_furthest_point_sampling_grid_lean(int b, int n, int m, const float *xyz, const void *grid,
                                                  int *idxs, float *new_xyz, void *stream);

/* Sampling a cloud that is already in sampling order (SA2-4 sample the previous level's
 * centres, models/backbone_module.py:52-86, where the reference itself notes the result "is just
 * 0,1,...,1023", :111).  bqa_fps_prefix_check proves that in parallel, per scene:
 *   run_flags[s] = 0  when furthest_point_sampling(xyz[s, :n'], m') is PROVABLY (0,1,...,m'-1) for
 *                     every n' <= n and m' <= min(m, n') -- every step has a strict unique maximum,
 *                     so no reduction order or tie-break can matter;
 *   run_flags[s] = 1  otherwise (ties, duplicates, points in the skip set, non-finite values).
 * v_scratch: b*m floats.  bqa_furthest_point_sampling_cond then writes the identity prefix
 * (idxs[s] = 0..m-1, new_xyz[s] = xyz[s, :m]) for scenes flagged 0 and runs the real serial chain
 * for the others, so its output always equals bqa_furthest_point_sampling's.  */
BQA_API int bqa_fps_prefix_check(int b, int n, int m, const float *xyz, float *v_scratch,
                                 int *run_flags, void *stream);
BQA_API int bqa_furthest_point_sampling_cond(int b, int n, int m, const float *xyz,
                                             const int *run_flags, int *idxs, float *new_xyz,
                                             float *scratch, void *stream);

/* Sliced sampling: produce samples j_begin .. j_end-1 only (1 <= j_begin <= j_end <= m; sample 0
 * is written by the slice with j_begin == 1).  Slices must be issued in order on one stream; the
 * running min-distances travel between them in `state` (b,n) f32.  Results are identical to one
 * full call.  Lets consumers of the first centres (ball query, SA MLP) run on another stream
 * underneath the later slices.  exclusive != 0 keeps other kernels off the SMs a slice runs on. */
BQA_API int bqa_furthest_point_sampling_slice(int b, int n, int m, int j_begin, int j_end,
                                              const float *xyz, int *idxs, float *new_xyz,
                                              float *state, int exclusive, void *stream);

/* ---- gather -------------------------------------------------------------------
 * replaces: gather_points / gather_points_grad, src/sampling.cpp:15-65,
 *           kernels src/sampling_gpu.cu:8-57.
 * points (b,c,n), idx (b,m) -> out (b,c,m);  grad_out (b,c,m) -> grad_points (b,c,n) */
BQA_API int bqa_gather_points(int b, int c, int n, int m, const float *points, const int *idx,
                      float *out, void *stream);
BQA_API int bqa_gather_points_grad(int b, int c, int n, int m, const float *grad_out,
                           const int *idx, float *grad_points, void *stream);

/* ---- ball query ---------------------------------------------------------------
 * replaces: ball_query(new_xyz (B,M,3), xyz (B,N,3), radius, nsample) -> (B,M,nsample)
 *           src/ball_query.cpp:8-32, kernel src/ball_query_gpu.cu:9-54.
 * First `nsample` indices k (ascending) with d2 < radius*radius, first hit back-fills
 * the row, empty ball -> zeros.
 * workspace: optional device scratch of bqa_ball_query_workspace_bytes(b,n,m,nsample) bytes
 * (0 for small scenes).  With it, scenes of >= 4096 points are binned into a cell grid and
 * each ball only visits the cells it touches, and mid-size scenes are scanned in index-ordered
 * segments by several CTAs per query block; with NULL a single CTA per query block scans the
 * whole scene.  The result is identical either way.  */
BQA_API long long bqa_ball_query_workspace_bytes(int b, int n, int m, int nsample);
BQA_API int bqa_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                   const float *xyz, int *idx, void *workspace, void *stream);

/* Slice form: only the centres [j_begin, j_begin + j_count) of each scene's m_total centres are
 * queried; new_xyz (b,m_total,3) and idx (b,m_total,nsample) keep their full-size layout.  The
 * workspace is sized for (b, n, j_count, nsample). */
BQA_API int bqa_ball_query_slice(int b, int n, int m_total, int j_begin, int j_count, float radius,
                                 int nsample, const float *new_xyz, const float *xyz, int *idx,
                                 void *workspace, void *stream);

/* Cell-grid form for large scenes (what bqa_ball_query runs internally for n >= 4096 when it is
 * given a workspace), split in two so that a caller can bin the scene -- which only needs xyz
 * -- on another stream while the sampling that produces new_xyz is still running:
 *   grid  : device buffer of bqa_ball_query_grid_bytes(b, n) bytes
 *   build : counting sort of each scene into cells of about `radius` (4 short launches)
 *   search: one warp per centre of the slice [j_begin, j_begin + j_count); identical rows to
 *           bqa_ball_query for ANY radius (the cell size only affects speed).  */
BQA_API long long bqa_ball_query_grid_bytes(int b, int n);
BQA_API int bqa_ball_query_grid_build(int b, int n, float radius, const float *xyz, void *grid,
                                      void *stream);
BQA_API int bqa_ball_query_grid_search(int b, int n, int m_total, int j_begin, int j_count,
                                       float radius, int nsample, const float *new_xyz,
                                       const float *xyz, int *idx, const void *grid, void *stream);

/* ---- grouping -----------------------------------------------------------------
 * replaces: group_points / group_points_grad, src/group_points.cpp:12-62,
 *           kernels src/group_points_gpu.cu:8-75.
 * points (b,c,n), idx (b,npoints,nsample) -> out (b,c,npoints,nsample) */
BQA_API int bqa_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                     const int *idx, float *out, void *stream);
/* group_points_grad through a point-major accumulator (workspace of
 * bqa_group_points_grad_workspace_bytes(b,c,n) bytes, 16-byte aligned): one vector reduction per
 * 4 channels instead of 4 scalar atomics, then a transpose into grad_points (b,c,n).  Same result
 * as bqa_group_points_grad up to the order of the floating-point additions (both use atomics). */
BQA_API long long bqa_group_points_grad_workspace_bytes(int b, int c, int n);
BQA_API int bqa_group_points_grad_ws(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                                     const int *idx, float *grad_points, void *workspace, void *stream);

/* QueryAndGroup.forward after the ball query (pointnet2_utils.py:347-359: group xyz, subtract the
 * centre, optionally divide by the radius, group the features, cat) in ONE pass, for a point-major
 * feature source feat_pm (b, n, feat_stride) -- e.g. the (B,N,3+C) input cloud itself:
 * out (b, 3+c, npoint, nsample) = [ (xyz[idx] - new_xyz) (* fl(1/radius), torch's CUDA evaluation
 * of `/= radius`) ; feat_pm[idx, :c]^T ] -- bit-identical to the reference sequence on the GPU.
 * Forward only (used when neither xyz nor the features need a gradient: SA1 of the backbone). */
BQA_API int bqa_group_concat_point_major(int b, int n, int c, int feat_stride, int npoint, int nsample,
                                         const float *xyz, const float *new_xyz, const float *feat_pm,
                                         const int *idx, float radius, int normalize_xyz, float *out,
                                         void *stream);
BQA_API int bqa_group_points_grad(int b, int c, int n, int npoints, int nsample,
                          const float *grad_out, const int *idx, float *grad_points,
                          void *stream);

/* ---- three_nn / three_interpolate --------------------------------------------
 * replaces: three_nn, three_interpolate, three_interpolate_grad,
 *           src/interpolate.cpp:14-99, kernels src/interpolate_gpu.cu:9-154.
 * unknown (b,n,3), known (b,m,3) -> dist2 (b,n,3) f32 (SQUARED), idx (b,n,3) i32
 * points (b,c,m), idx (b,n,3), weight (b,n,3) -> out (b,c,n) */
BQA_API int bqa_three_nn(int b, int n, int m, const float *unknown, const float *known,
                 float *dist2, int *idx, void *stream);
BQA_API int bqa_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx,
                          const float *weight, float *out, void *stream);
BQA_API int bqa_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out,
                               const int *idx, const float *weight, float *grad_points,
                               void *stream);

/* ---- fused set-abstraction layer (inference; eval-mode BN folded into W/bias) ---------
 * replaces the chain  QueryAndGroup.forward (pointnet2_utils.py:317-376: group xyz,
 * subtract centre, divide by radius, group features, concat) -> SharedMLP of three
 * [1x1 conv -> BN -> ReLU] blocks (pytorch_utils.py:11-36) -> max_pool2d over nsample
 * (pointnet2_modules.py:259-262), without materialising the grouped tensor or any
 * activation in HBM.  16-bit operands, fp32 accumulation on tcgen05 tensor cores
 * (kind::f16).  `precision`: 0 = bf16 operands (8-bit mantissa, fp32 range); 1 = fp16
 * operands (11-bit mantissa -- the same as TF32, which is what the reference's cuDNN convs
 * use by default -- values saturate at +-65504).  Same speed either way.
 *
 * bqa_pack_weight_16: w (c_out, c_in) f32 row-major (BN already folded) -> the 16-bit
 *   shared-memory image the kernel loads, [k_pad/8][c_out][8], k_pad a multiple of 16,
 *   zero padded.  xyz_first=1: source columns are [xyz(3), feat(c_in-3)] (torch.cat order
 *   at pointnet2_utils.py:357) and are re-ordered to [feat, xyz] to match the kernel's
 *   gather.  `packed` needs 2*c_out*k_pad bytes.
 * bqa_sa_mlp_max_supported: 1 if the shape is handled (nsample in {16,32,64,128},
 *   npoint*nsample a multiple of 128, mlp widths (64,64,128) | (128,128,256) | (128,128,128)).
 * bqa_sa_mlp_max_forward: xyz (b,n,3) f32; new_xyz (b,npoint,3) f32; feat_pm (b,n,c)
 *   POINT-MAJOR f32 (NULL iff c == 0) whose consecutive points are feat_stride floats apart
 *   (feat_stride >= c; scenes are n*feat_stride apart -- lets SA1 read the features straight
 *   out of the (b,n,3+c) input cloud); idx (b,npoint,nsample) i32 from bqa_ball_query;
 *   w{1,2,3}p packed (same `precision`) with k_pad = roundup16(c+3), c1, c2; b{1,2,3} f32.
 *   grouped xyz is (xyz[idx] - new_xyz) and, when normalize_xyz, divided by `radius`.
 *   out_cm (b,c3,npoint) f32 = the reference's new_features; out_pm (b,npoint,c3) f32
 *   optional point-major copy for the next layer (NULL to skip). */
BQA_API int bqa_pack_weight_16(int c_out, int c_in, int k_pad, int xyz_first, int precision,
                               const float *w, void *packed, void *stream);
BQA_API int bqa_sa_mlp_max_supported(int nsample, int npoint, int c, int c1, int c2, int c3);
BQA_API int bqa_sa_mlp_max_forward(int b, int n, int npoint, int nsample, int c, const float *xyz,
                                   const float *new_xyz, const float *feat_pm, int feat_stride,
                                   const int *idx,
                                   float radius, int normalize_xyz, int c1, int c2, int c3,
                                   const void *w1p, const float *b1, const void *w2p,
                                   const float *b2, const void *w3p, const float *b3,
                                   float *out_cm, float *out_pm, int precision, void *stream);

/* ---- fused set-abstraction layer, warp-specialised form (csrc/sa_fused_v2.cu) ------------------
 * Same contract and arithmetic class as bqa_sa_mlp_max_forward (same reference lines:
 * pointnet2_utils.py:317-376 + pytorch_utils.py:11-36 + pointnet2_modules.py:259-262), different
 * execution: producer / MMA-issuer / epilogue warps of one persistent CTA work on up to four tiles
 * in flight, features are gathered from a 16-BIT point-major copy with cp.async, biases are folded
 * into the tensor-core contraction.
 *
 * bqa_pack_weight_16_v2: w (c_out, c_in) f32 (BN folded), bias (c_out) f32 or NULL -> 16-bit image
 *   [k_pad/8][c_out][8].  mode 0: plain, k_pad >= c_in (layer 3; bias is passed to the kernel as
 *   f32).  mode 1: layer 1 of an SA block, source columns [xyz(3), feat(c)] (c = c_in - 3), k_pad =
 *   roundup16(c + 5): feature f at K = f, xyz at K = kx..kx+2 with kx = max(c, k_pad - 16), the
 *   bias as two 16-bit halves (hi + lo) at kx+3, kx+4.  mode 2: layer 2, k_pad = c_in + 16, bias
 *   halves at K = c_in, c_in + 1.
 * bqa_to_point_major_16: (b, c, n) f32 channel-major -> (b, n, stride) 16-bit point-major, zero
 *   padded from c to stride (stride a multiple of 8, >= c).
 * bqa_rows_to_16: rows x [first, first + c) of an f32 matrix with row_stride floats per row (the
 *   (b*n, 3 + C) input cloud: first = 3) -> (rows, stride) 16-bit, zero padded.
 * bqa_sa_mlp_max_v2_supported: nsample in {16,32,64} (128 for the (64,64,128) widths), npoint a
 *   multiple of 128/nsample, widths (64,64,128) | (128,128,256) | (128,128,128), c <= 1024.
 * bqa_sa_mlp_max_forward_v2: feat16 (b, n, stride16) 16-bit point-major in `precision`'s format
 *   (NULL iff c == 0; stride16 a multiple of 8 and >= roundup8(c), rows zero padded, 16-byte
 *   aligned); w1p / w2p / w3p from bqa_pack_weight_16_v2 modes 1 / 2 / 0; b3 f32.  Outputs:
 *   out_cm (b,c3,npoint) f32; optional out_pm (b,npoint,c3) f32 and out_pm16 (b,npoint,c3) 16-bit
 *   point-major copies for the next layer. */
BQA_API int bqa_pack_weight_16_v2(int c_out, int c_in, int k_pad, int mode, int precision,
                                  const float *w, const float *bias, void *packed, void *stream);
BQA_API int bqa_to_point_major_16(int b, int c, int n, int stride, int precision, const float *in,
                                  void *out, void *stream);
BQA_API int bqa_rows_to_16(long long rows, int c, int row_stride, int first, int stride, int precision,
                           const float *in, void *out, void *stream);
BQA_API int bqa_sa_mlp_max_v2_supported(int nsample, int npoint, int c, int c1, int c2, int c3);
BQA_API int bqa_sa_mlp_max_forward_v2(int b, int n, int npoint, int nsample, int c, const float *xyz,
                                      const float *new_xyz, const void *feat16, int stride16,
                                      const int *idx, float radius, int normalize_xyz, int c1, int c2,
                                      int c3, const void *w1p, const void *w2p, const void *w3p,
                                      const float *b3, float *out_cm, float *out_pm, void *out_pm16,
                                      int precision, void *stream);

/* ---- fused feature-propagation layer (inference) -----------------------------------
 * replaces PointnetFPModule.forward (pointnet2_modules.py:376-421): three_nn ->
 * 1/(dist+1e-8) weights normalised over the 3 neighbours -> three_interpolate -> cat with
 * the skip features -> two [1x1 conv -> BN -> ReLU] blocks, in one kernel.
 * unknown (b,n,3), known (b,m,3) f32; known_feat (b,m,c_known) and skip_feat (b,n,c_skip)
 * POINT-MAJOR f32 with the given row strides (floats, multiples of 4, 16-byte aligned);
 * layer-1 input channel order is [interpolated(c_known), skip(c_skip)] like torch.cat at
 * pointnet2_modules.py:413.  w: layer-1 image (k_pad = c_known+c_skip) immediately followed
 * by the layer-2 image (k_pad = c1), both from bqa_pack_weight_16(xyz_first=0) with the same
 * `precision`; b1, b2 f32.  out_cm (b,c2,n) f32; out_pm (b,n,c2) f32 optional.
 * skip16 (optional, else NULL): a 16-bit point-major copy of the skip features in `precision`'s
 * format (what bqa_sa_mlp_max_forward_v2 writes as out_pm16), rows skip16_stride elements apart
 * (multiple of 8): copied straight into the operand with cp.async; skip_feat may then be NULL.
 * bqa_fp_mlp_supported: c1 == c2 == 256 and c_known, c_skip multiples of 64. */
BQA_API int bqa_fp_mlp_supported(int n, int m, int c_known, int c_skip, int c1, int c2);
BQA_API int bqa_fp_mlp_forward(int b, int n, int m, int c_known, int c_skip, const float *unknown,
                               const float *known, const float *known_feat, int known_stride,
                               const float *skip_feat, int skip_stride, int c1, int c2,
                               const void *w, const float *b1, const float *b2, float *out_cm,
                               float *out_pm, int precision, void *stream, const void *skip16,
                               int skip16_stride);

/* ---- nn_distance (first consumer of the hot path's outputs) -------------
 * replaces: utils/nn_distance.py:25-52 `nn_distance(pc1 (B,N,3), pc2 (B,M,3), l1smooth, delta, l1)`
 *           -> dist1 (B,N) f32, idx1 (B,N) i64, dist2 (B,M) f32, idx2 (B,M) i64, as used by the
 *           VoteNet losses (lib/loss_helper.py:66,91,144).  mode 0: sum of squares, 1: L1, 2: Huber.
 * One pass, no (B,N,M,3) tensor; bit-identical to the torch expression on the GPU.  */
BQA_API int bqa_nn_distance(int b, int n, int m, const float *pc1, const float *pc2, int mode, float delta,
                            float *dist1, long long *idx1, float *dist2, long long *idx2, void *stream);

/* ---- detection post-processing ------------------------------------------
 * replaces the host loops of lib/ap_helper.py:86-178 (parse_predictions):
 *   bqa_count_points_in_boxes: counts (b,k) = points of xyz (b,n,3) inside each axis-aligned box
 *       box_lo_hi (b,k,6) = {x1,y1,z1,x2,y2,z2}, bounds inclusive -- the `remove_empty_box` test
 *       (:89-100; ScanNet boxes have one heading bin, i.e. they are axis-aligned);
 *   bqa_nms3d: greedy 3-D NMS of utils/nms.py:74-152 (nms_3d_faster / _samecls) per scene, in
 *       double precision like the reference's float64 arrays.  boxes (b,k,8) = {x1,y1,z1,x2,y2,z2,
 *       score,class}; valid (b,k) or NULL masks boxes out before NMS; pick_mask (b,k) 0/1;
 *       pick_order (b,k) or NULL lists the picked indices in pick order, -1 padded.  k <= 1024. */
BQA_API int bqa_nms3d(int b, int k, const float *boxes, const int *valid, double iou_threshold,
                      int old_type, int same_class, int *pick_mask, int *pick_order, void *stream);
BQA_API int bqa_count_points_in_boxes(int b, int n, int k, const float *xyz, const float *box_lo_hi,
                                      int *counts, void *stream);

/* ---- train-mode BatchNorm + ReLU (+ max over nsample) --------------------------------------
 * replaces, in model.train(), the nn.BatchNorm2d (training=True) + shared nn.ReLU that follow
 * every 1x1 conv of a SharedMLP block (lib/pointnet2/pytorch_utils.py:11-36, 73-80) and, for the
 * last block of an SA layer, the F.max_pool2d over nsample (pointnet2_modules.py:259-262), with
 * their backward.  y, x, dx, dy: (b, c, l) contiguous fp32, 16-byte aligned (l = npoint*nsample).
 *   bqa_bn_train_stats     : mean / invstd of the batch (biased variance, eps inside the sqrt) and
 *                            the running-stat update (unbiased variance, momentum); either
 *                            running pointer may be NULL.  sums_scratch: 2*c doubles.
 *   bqa_bn_relu_forward    : x = relu(gamma * (y - mean) * invstd + beta)
 *   bqa_bn_relu_max_forward: out (b,c,npoint) = max over nsample of the same, argmax (first
 *                            occurrence, like max_pool2d) -- the activation is never written
 *   bqa_bn_relu_backward   : dy, dgamma, dbeta from dx (gradient w.r.t. x) -- BatchNorm's
 *                            training-mode backward with the ReLU mask recomputed from y
 *   bqa_bn_relu_max_backward: same from dout (b,c,npoint) + argmax                              */
BQA_API int bqa_bn_relu_max_supported(int nsample);
/* bqa_bn_finalize_shifted: mean / invstd (+ running-stat update) from sums = [sum (y-K), sum (y-K)^2]
 * (2*c doubles) accumulated by bqa_conv1x1_tf32_forward around K = shift[ch] (NULL: 0); count = b*l.
 * shift may be running_mean itself (read before it is updated). */
BQA_API int bqa_bn_finalize_shifted(int c, double count, const double *sums, const float *shift, float eps,
                                    float momentum, float *mean, float *invstd, float *running_mean,
                                    float *running_var, void *stream);

/* ---- 1x1 convolution of a SharedMLP block in training mode (tcgen05, TF32 operands, fp32 accumulate)
 * replaces the cuDNN calls behind nn.Conv2d / nn.Conv1d (kernel 1, no bias) of
 * lib/pointnet2/pytorch_utils.py:104-157 and their autograd backward.
 *   forward: y (b,cout,p) = w (cout, cin; rows ldw floats apart, ldw % 4 == 0) . x (b,cin,p); p % 4 == 0;
 *            sums (optional, 2*cout doubles, ACCUMULATED into: zero them first) receive the BatchNorm
 *            batch statistics sum (y - shift[c]), sum (y - shift[c])^2 from the epilogue (shift NULL: 0).
 *            The data gradient is the same call with w^T (cin, cout) and dy: dx = w^T . dy.
 *   wgrad:   dw (cout, cin) += sum over b, p of dy[b,:,p] x[b,:,p]^T  (zero dw first; cout <= 256).
 * All tensors 16-byte aligned. */
BQA_API int bqa_conv1x1_tf32_supported(int b, int cin, int cout, long long p, int ldw);
BQA_API int bqa_conv1x1_tf32_forward(int b, int cin, int cout, long long p, const float *x, const float *w,
                                     int ldw, float *y, const float *shift, double *sums, void *stream);
BQA_API int bqa_conv1x1_tf32_wgrad(int b, int cin, int cout, long long p, const float *x, const float *dy,
                                   float *dw, void *stream);
BQA_API int bqa_bn_train_stats(int b, int c, long long l, const float *y, double *sums_scratch, float eps,
                               float momentum, float *mean, float *invstd, float *running_mean,
                               float *running_var, void *stream);
BQA_API int bqa_bn_relu_forward(int b, int c, long long l, const float *y, const float *mean,
                                const float *invstd, const float *gamma, const float *beta, float *x,
                                void *stream);
BQA_API int bqa_bn_relu_max_forward(int b, int c, int npoint, int nsample, const float *y,
                                    const float *mean, const float *invstd, const float *gamma,
                                    const float *beta, float *out, int *argmax, void *stream);
BQA_API int bqa_bn_relu_backward(int b, int c, long long l, const float *dx, const float *y,
                                 const float *mean, const float *invstd, const float *gamma,
                                 const float *beta, double *sums_scratch, float *dy, float *dgamma,
                                 float *dbeta, void *stream);
BQA_API int bqa_bn_relu_max_backward(int b, int c, int npoint, int nsample, const float *dout,
                                     const int *argmax, const float *y, const float *mean,
                                     const float *invstd, const float *gamma, const float *beta,
                                     double *sums_scratch, float *dy, float *dgamma, float *dbeta,
                                     void *stream);

/* ---- layout helper -------------------------------------------------------------
 * (b,c,n) channel-major -> (b,n,c) point-major, the layout the fused SA kernel gathers
 * from (one contiguous row per neighbour).  Replaces nothing in the reference; it is
 * the inverse of the transpose in Pointnet2Backbone._break_up_pc
 * (models/backbone_module.py:74-78). */
BQA_API int bqa_transpose_to_point_major(int b, int c, int n, const float *in, float *out,
                                 void *stream);

#ifdef __cplusplus
}
#endif
#endif /* BQA_POINTNET2_H_ */
