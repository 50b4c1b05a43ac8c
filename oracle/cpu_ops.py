"""ctypes front-end of oracle/pointnet2_oracle.c (numpy in, numpy out).

TEST INFRASTRUCTURE ONLY -- the checker and the reported CPU baseline, never
the product path.  Function names and argument order follow the reference's
pybind module (`/root/reference/lib/pointnet2/_ext_src/src/bindings.cpp:6-19`)
so that `as_ext_module()` can stand in for `pointnet2._ext` under the
reference's own Python layer on a GPU-less machine.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "pointnet2_oracle.c")
SO = os.path.join(HERE, "libbqa_oracle.so")

_lib = None


def build(force=False):
    """gcc -O2 -ffp-contract=off (every fused multiply-add is spelled fmaf())."""
    if (not force and os.path.exists(SO)
            and (not os.path.exists(SRC) or os.path.getmtime(SO) >= os.path.getmtime(SRC))):
        return SO
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC",
           "-shared", "-std=gnu11", "-o", SO, SRC, "-lm"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout)
    return SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(SO)
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def num_threads():
    return int(lib().bqa_oracle_num_threads())


def set_num_threads(t):
    lib().bqa_oracle_set_num_threads(int(t))


def opt_n_threads(n):
    return int(lib().bqa_oracle_opt_n_threads(int(n)))


def furthest_point_sampling(xyz, npoint):
    xyz, px = _f(xyz)
    b, n, _ = xyz.shape
    out = np.zeros((b, npoint), dtype=np.int32)
    lib().bqa_oracle_furthest_point_sampling(b, n, int(npoint), px, out.ctypes.data_as(ctypes.c_void_p))
    return out


def gather_points(points, idx):
    points, pp = _f(points)
    idx, pi = _i(idx)
    b, c, n = points.shape
    m = idx.shape[1]
    out = np.zeros((b, c, m), dtype=np.float32)
    lib().bqa_oracle_gather_points(b, c, n, m, pp, pi, out.ctypes.data_as(ctypes.c_void_p))
    return out


def gather_points_grad(grad_out, idx, n):
    grad_out, pg = _f(grad_out)
    idx, pi = _i(idx)
    b, c, m = grad_out.shape
    out = np.zeros((b, c, n), dtype=np.float32)
    lib().bqa_oracle_gather_points_grad(b, c, int(n), m, pg, pi, out.ctypes.data_as(ctypes.c_void_p))
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    new_xyz, pq = _f(new_xyz)
    xyz, px = _f(xyz)
    b, m, _ = new_xyz.shape
    n = xyz.shape[1]
    out = np.zeros((b, m, nsample), dtype=np.int32)
    lib().bqa_oracle_ball_query(b, n, m, ctypes.c_float(radius), int(nsample), pq, px,
                                out.ctypes.data_as(ctypes.c_void_p))
    return out


def group_points(points, idx):
    points, pp = _f(points)
    idx, pi = _i(idx)
    b, c, n = points.shape
    _, npoints, nsample = idx.shape
    out = np.zeros((b, c, npoints, nsample), dtype=np.float32)
    lib().bqa_oracle_group_points(b, c, n, npoints, nsample, pp, pi, out.ctypes.data_as(ctypes.c_void_p))
    return out


def group_points_grad(grad_out, idx, n):
    grad_out, pg = _f(grad_out)
    idx, pi = _i(idx)
    b, c, npoints, nsample = grad_out.shape
    out = np.zeros((b, c, n), dtype=np.float32)
    lib().bqa_oracle_group_points_grad(b, c, int(n), npoints, nsample, pg, pi,
                                       out.ctypes.data_as(ctypes.c_void_p))
    return out


def three_nn(unknown, known):
    unknown, pu = _f(unknown)
    known, pk = _f(known)
    b, n, _ = unknown.shape
    m = known.shape[1]
    dist2 = np.zeros((b, n, 3), dtype=np.float32)
    idx = np.zeros((b, n, 3), dtype=np.int32)
    lib().bqa_oracle_three_nn(b, n, m, pu, pk, dist2.ctypes.data_as(ctypes.c_void_p),
                              idx.ctypes.data_as(ctypes.c_void_p))
    return dist2, idx


def three_interpolate(points, idx, weight):
    points, pp = _f(points)
    idx, pi = _i(idx)
    weight, pw = _f(weight)
    b, c, m = points.shape
    n = idx.shape[1]
    out = np.zeros((b, c, n), dtype=np.float32)
    lib().bqa_oracle_three_interpolate(b, c, m, n, pp, pi, pw, out.ctypes.data_as(ctypes.c_void_p))
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    grad_out, pg = _f(grad_out)
    idx, pi = _i(idx)
    weight, pw = _f(weight)
    b, c, n = grad_out.shape
    out = np.zeros((b, c, m), dtype=np.float32)
    lib().bqa_oracle_three_interpolate_grad(b, c, n, int(m), pg, pi, pw,
                                            out.ctypes.data_as(ctypes.c_void_p))
    return out


class _TorchExt:
    """`pointnet2._ext` look-alike over CPU torch tensors (bindings.cpp:6-19)."""

    def __getattr__(self, name):
        import torch
        fn = globals()[name]

        def call(*args):
            conv = [a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a for a in args]
            out = fn(*conv)
            if isinstance(out, tuple):
                return [torch.from_numpy(o) for o in out]
            return torch.from_numpy(out)

        return call


def as_ext_module():
    return _TorchExt()
