"""Tensor-level mirror of the reference's `pointnet2._ext` pybind module.

Same nine function names, argument order, shapes, dtypes and error behaviour as
/root/reference/lib/pointnet2/_ext_src/src/bindings.cpp:6-19 (wrappers in
src/{sampling,ball_query,group_points,interpolate}.cpp), so the reference's own
pointnet2_utils.py runs on top of this module unchanged.  Each
function allocates its outputs with torch (the caching allocator owns the memory, as in
the reference wrappers), then calls the C ABI on the current stream of the inputs'
device.  Nothing here computes on the host.
"""
import ctypes

import torch

from . import _native as N

_f32 = torch.float32
_i32 = torch.int32


def _guard(t):
    return torch.cuda.device(t.device)


def furthest_point_sampling(points, nsamples, return_xyz=False):
    """sampling.cpp:66-87.  points (B,N,3) f32 -> (B,nsamples) int32.

    return_xyz=True additionally returns the sampled coordinates (B,nsamples,3), produced
    by the same kernel (the fused form of gather_points on xyz)."""
    N.check_tensor(points, "points", _f32)
    if points.dim() != 3 or points.size(2) != 3:
        raise RuntimeError("points must be (B, N, 3)")
    b, n, _ = points.shape
    nsamples = int(nsamples)
    out = torch.zeros((b, nsamples), dtype=_i32, device=points.device)
    new_xyz = torch.empty((b, nsamples, 3), dtype=_f32, device=points.device) if return_xyz else None
    with _guard(points):
        nbytes = N.lib().bqa_fps_scratch_bytes(b, n)
        scratch = torch.empty((nbytes // 4,), dtype=_f32, device=points.device) if nbytes else None
        N.call("bqa_furthest_point_sampling", b, n, nsamples, N.ptr(points), N.ptr(out),
               N.ptr(new_xyz), N.ptr(scratch), N.stream_ptr(points.device))
    return (out, new_xyz) if return_xyz else out


def gather_points(points, idx):
    """sampling.cpp:15-38.  points (B,C,N) f32, idx (B,M) int32 -> (B,C,M)."""
    N.check_tensor(points, "points", _f32)
    N.check_tensor(idx, "idx", _i32)
    b, c, n = points.shape
    m = idx.size(1)
    out = torch.empty((b, c, m), dtype=_f32, device=points.device)
    with _guard(points):
        N.call("bqa_gather_points", b, c, n, m, N.ptr(points), N.ptr(idx), N.ptr(out),
               N.stream_ptr(points.device))
    return out


def gather_points_grad(grad_out, idx, n):
    """sampling.cpp:40-65.  grad_out (B,C,M), idx (B,M) -> (B,C,n)."""
    N.check_tensor(grad_out, "grad_out", _f32)
    N.check_tensor(idx, "idx", _i32)
    b, c, m = grad_out.shape
    out = torch.empty((b, c, int(n)), dtype=_f32, device=grad_out.device)
    with _guard(grad_out):
        N.call("bqa_gather_points_grad", b, c, int(n), m, N.ptr(grad_out), N.ptr(idx), N.ptr(out),
               N.stream_ptr(grad_out.device))
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    """ball_query.cpp:8-32.  new_xyz (B,M,3), xyz (B,N,3) -> (B,M,nsample) int32."""
    N.check_tensor(new_xyz, "new_xyz", _f32)
    N.check_tensor(xyz, "xyz", _f32)
    b, m, _ = new_xyz.shape
    n = xyz.size(1)
    nsample = int(nsample)
    idx = torch.empty((b, m, nsample), dtype=_i32, device=new_xyz.device)
    with _guard(new_xyz):
        nbytes = N.lib().bqa_ball_query_workspace_bytes(b, n, m, nsample)
        work = torch.empty((nbytes,), dtype=torch.uint8, device=new_xyz.device) if nbytes else None
        N.call("bqa_ball_query", b, n, m, ctypes.c_float(radius), nsample, N.ptr(new_xyz),
               N.ptr(xyz), N.ptr(idx), N.ptr(work), N.stream_ptr(new_xyz.device))
    return idx


def group_points(points, idx):
    """group_points.cpp:12-36.  points (B,C,N), idx (B,M,S) int32 -> (B,C,M,S)."""
    N.check_tensor(points, "points", _f32)
    N.check_tensor(idx, "idx", _i32)
    b, c, n = points.shape
    _, npoints, nsample = idx.shape
    out = torch.empty((b, c, npoints, nsample), dtype=_f32, device=points.device)
    with _guard(points):
        N.call("bqa_group_points", b, c, n, npoints, nsample, N.ptr(points), N.ptr(idx),
               N.ptr(out), N.stream_ptr(points.device))
    return out


def group_concat_point_major(xyz, new_xyz, feat_pm, idx, radius, normalize_xyz):
    """QueryAndGroup's grouping + centring + cat (pointnet2_utils.py:347-359) in one pass from a
    point-major feature view feat_pm (B,N,C) with unit channel stride -> (B, 3+C, M, S)."""
    N.check_tensor(xyz, "xyz", _f32)
    N.check_tensor(new_xyz, "new_xyz", _f32)
    N.check_tensor(idx, "idx", _i32)
    if feat_pm.dtype != _f32 or not feat_pm.is_cuda or feat_pm.stride(2) != 1 or \
            feat_pm.stride(0) != feat_pm.size(1) * feat_pm.stride(1):
        raise RuntimeError("feat_pm must be a CUDA float (B,N,C) view with unit channel stride")
    b, n, c = feat_pm.shape
    _, npoints, nsample = idx.shape
    out = torch.empty((b, 3 + c, npoints, nsample), dtype=_f32, device=xyz.device)
    with _guard(xyz):
        N.call("bqa_group_concat_point_major", b, n, c, feat_pm.stride(1), npoints, nsample, N.ptr(xyz),
               N.ptr(new_xyz), N.ptr(feat_pm), N.ptr(idx), ctypes.c_float(radius),
               1 if normalize_xyz else 0, N.ptr(out), N.stream_ptr(xyz.device))
    return out


def group_points_grad(grad_out, idx, n):
    """group_points.cpp:38-62.  grad_out (B,C,M,S), idx (B,M,S) -> (B,C,n).  Wide tensors go through
    the point-major accumulator (vector reductions, bqa_group_points_grad_ws)."""
    N.check_tensor(grad_out, "grad_out", _f32)
    N.check_tensor(idx, "idx", _i32)
    b, c, npoints, nsample = grad_out.shape
    out = torch.empty((b, c, int(n)), dtype=_f32, device=grad_out.device)
    with _guard(grad_out):
        if c >= 8:
            nbytes = N.lib().bqa_group_points_grad_workspace_bytes(b, c, int(n))
            work = torch.empty((max(nbytes, 16) // 4,), dtype=_f32, device=grad_out.device)
            N.call("bqa_group_points_grad_ws", b, c, int(n), npoints, nsample, N.ptr(grad_out),
                   N.ptr(idx), N.ptr(out), N.ptr(work), N.stream_ptr(grad_out.device))
        else:
            N.call("bqa_group_points_grad", b, c, int(n), npoints, nsample, N.ptr(grad_out),
                   N.ptr(idx), N.ptr(out), N.stream_ptr(grad_out.device))
    return out


def three_nn(unknowns, knows):
    """interpolate.cpp:14-40.  -> [dist2 (B,n,3) f32 (squared), idx (B,n,3) int32]."""
    N.check_tensor(unknowns, "unknowns", _f32)
    N.check_tensor(knows, "knows", _f32)
    b, n, _ = unknowns.shape
    m = knows.size(1)
    dist2 = torch.empty((b, n, 3), dtype=_f32, device=unknowns.device)
    idx = torch.empty((b, n, 3), dtype=_i32, device=unknowns.device)
    with _guard(unknowns):
        N.call("bqa_three_nn", b, n, m, N.ptr(unknowns), N.ptr(knows), N.ptr(dist2), N.ptr(idx),
               N.stream_ptr(unknowns.device))
    return [dist2, idx]


def three_interpolate(points, idx, weight):
    """interpolate.cpp:42-70.  points (B,C,m), idx (B,n,3), weight (B,n,3) -> (B,C,n)."""
    N.check_tensor(points, "points", _f32)
    N.check_tensor(idx, "idx", _i32)
    N.check_tensor(weight, "weight", _f32)
    b, c, m = points.shape
    n = idx.size(1)
    out = torch.empty((b, c, n), dtype=_f32, device=points.device)
    with _guard(points):
        N.call("bqa_three_interpolate", b, c, m, n, N.ptr(points), N.ptr(idx), N.ptr(weight),
               N.ptr(out), N.stream_ptr(points.device))
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    """interpolate.cpp:71-99.  grad_out (B,C,n) -> (B,C,m)."""
    N.check_tensor(grad_out, "grad_out", _f32)
    N.check_tensor(idx, "idx", _i32)
    N.check_tensor(weight, "weight", _f32)
    b, c, n = grad_out.shape
    out = torch.empty((b, c, int(m)), dtype=_f32, device=grad_out.device)
    with _guard(grad_out):
        N.call("bqa_three_interpolate_grad", b, c, n, int(m), N.ptr(grad_out), N.ptr(idx),
               N.ptr(weight), N.ptr(out), N.stream_ptr(grad_out.device))
    return out


def transpose_to_point_major(features):
    """(B,C,N) -> (B,N,C) contiguous."""
    N.check_tensor(features, "features", _f32)
    b, c, n = features.shape
    out = torch.empty((b, n, c), dtype=_f32, device=features.device)
    with _guard(features):
        N.call("bqa_transpose_to_point_major", b, c, n, N.ptr(features), N.ptr(out),
               N.stream_ptr(features.device))
    return out
