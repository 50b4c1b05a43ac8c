"""`nn_distance` / `huber_loss` of the reference's utils/nn_distance.py on one sm_100a kernel.

Drop-in for `from utils.nn_distance import nn_distance, huber_loss` (lib/loss_helper.py:13): same
signature, same outputs (dist float32, idx int64), differentiable w.r.t. both point sets like the
torch expression (the minimum routes the gradient to the selected pair).  Forward is bit-identical
to the reference expression on the GPU; it never builds the (B,N,M,3) difference tensor.
"""
import ctypes

import torch
from torch.autograd import Function

from . import _native as N

_f32 = torch.float32


def huber_loss(error, delta=1.0):
    """utils/nn_distance.py:6-23 (elementwise; plain torch, it is not on the distance path)."""
    abs_error = torch.abs(error)
    quadratic = torch.clamp(abs_error, max=delta)
    linear = abs_error - quadratic
    return 0.5 * quadratic ** 2 + delta * linear


def _dterm(diff, mode, delta):
    if mode == 0:
        return 2.0 * diff
    if mode == 1:
        return torch.sign(diff)
    return torch.clamp(diff, min=-delta, max=delta)


class _NNDistance(Function):
    @staticmethod
    def forward(ctx, pc1, pc2, mode, delta):
        N.check_tensor(pc1, "pc1", _f32)
        N.check_tensor(pc2, "pc2", _f32)
        if pc1.dim() != 3 or pc2.dim() != 3 or pc1.size(2) != 3 or pc2.size(2) != 3 or pc1.size(0) != pc2.size(0):
            raise RuntimeError("nn_distance: pc1 (B,N,3) and pc2 (B,M,3) expected")
        b, n, _ = pc1.shape
        m = pc2.size(1)
        dev = pc1.device
        dist1 = torch.empty((b, n), dtype=_f32, device=dev)
        dist2 = torch.empty((b, m), dtype=_f32, device=dev)
        idx1 = torch.empty((b, n), dtype=torch.int64, device=dev)
        idx2 = torch.empty((b, m), dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            N.call("bqa_nn_distance", b, n, m, N.ptr(pc1), N.ptr(pc2), int(mode), ctypes.c_float(delta),
                   N.ptr(dist1), N.ptr(idx1), N.ptr(dist2), N.ptr(idx2), N.stream_ptr(dev))
        ctx.save_for_backward(pc1, pc2, idx1, idx2)
        ctx.mode, ctx.delta = int(mode), float(delta)
        ctx.mark_non_differentiable(idx1, idx2)
        return dist1, idx1, dist2, idx2

    @staticmethod
    def backward(ctx, g1, _gi1, g2, _gi2):
        pc1, pc2, idx1, idx2 = ctx.saved_tensors
        e1 = idx1.unsqueeze(-1).expand(-1, -1, 3)
        e2 = idx2.unsqueeze(-1).expand(-1, -1, 3)
        # dist1[b,i] = f(pc1[b,i] - pc2[b,idx1[b,i]]);  dist2[b,j] = f(pc1[b,idx2[b,j]] - pc2[b,j])
        d1 = _dterm(pc1 - torch.gather(pc2, 1, e1), ctx.mode, ctx.delta) * g1.unsqueeze(-1)
        d2 = _dterm(torch.gather(pc1, 1, e2) - pc2, ctx.mode, ctx.delta) * g2.unsqueeze(-1)
        gpc1 = d1.clone().scatter_add_(1, e2, d2)
        gpc2 = (-d2).scatter_add_(1, e1, -d1)
        return gpc1, gpc2, None, None


def nn_distance(pc1, pc2, l1smooth=False, delta=1.0, l1=False):
    """utils/nn_distance.py:25-52.  pc1 (B,N,3), pc2 (B,M,3) ->
    dist1 (B,N), idx1 (B,N) int64, dist2 (B,M), idx2 (B,M) int64."""
    mode = 2 if l1smooth else (1 if l1 else 0)
    return _NNDistance.apply(pc1.contiguous(), pc2.contiguous(), mode, delta)
