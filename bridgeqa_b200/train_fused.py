"""Train-mode BatchNorm + ReLU (+ max over nsample) of a SharedMLP block as ONE autograd node
on the streaming kernels of csrc/bn_relu.cu.

The reference runs, after every 1x1 conv of an SA / FP layer in model.train(), nn.BatchNorm2d
(batch statistics) and the shared nn.ReLU(inplace=True) (lib/pointnet2/pytorch_utils.py:11-36,
73-80), and F.max_pool2d over nsample after the last block of an SA layer
(pointnet2_modules.py:259-262).  Those passes (cuDNN bn_fw_tr / bn_bw, max_pool fwd/bwd,
threshold_backward) are 25 of the 42 ms of a DET training step on B200.  Same math here
(biased variance for the normalisation, unbiased for running_var, momentum update, first-max
argmax like max_pool2d); parity with the torch modules is tested to 1e-5 / 1e-4 (gradients).
"""
import ctypes

import torch
from torch.autograd import Function

from . import _native as N

_f32 = torch.float32
_state = {"enabled": True}


def set_enabled(flag):
    _state["enabled"] = bool(flag)


def enabled():
    return _state["enabled"]


def _dims(y):
    b, c = y.shape[0], y.shape[1]
    return b, c, y.numel() // max(b * c, 1)


def _stats(y, norm):
    """Batch mean / invstd (+ in-place running-stat update, like F.batch_norm(training=True))."""
    b, c, l = _dims(y)
    dev = y.device
    mean = torch.empty((c,), dtype=_f32, device=dev)
    invstd = torch.empty((c,), dtype=_f32, device=dev)
    sums = torch.empty((2 * c,), dtype=torch.float64, device=dev)
    track = norm.track_running_stats and norm.running_mean is not None
    with torch.cuda.device(dev):
        N.call("bqa_bn_train_stats", b, c, l, N.ptr(y), N.ptr(sums), ctypes.c_float(norm.eps),
               ctypes.c_float(norm.momentum), N.ptr(mean), N.ptr(invstd),
               N.ptr(norm.running_mean if track else None), N.ptr(norm.running_var if track else None),
               N.stream_ptr(dev))
    if track and norm.num_batches_tracked is not None:
        norm.num_batches_tracked.add_(1)
    return mean, invstd


class _BNReLU(Function):
    @staticmethod
    def forward(ctx, y, weight, bias, norm):
        y = y.contiguous()
        mean, invstd = _stats(y, norm)
        b, c, l = _dims(y)
        x = torch.empty_like(y)
        with torch.cuda.device(y.device):
            N.call("bqa_bn_relu_forward", b, c, l, N.ptr(y), N.ptr(mean), N.ptr(invstd), N.ptr(weight),
                   N.ptr(bias), N.ptr(x), N.stream_ptr(y.device))
        ctx.save_for_backward(y, weight, bias, mean, invstd)
        return x

    @staticmethod
    def backward(ctx, dx):
        y, weight, bias, mean, invstd = ctx.saved_tensors
        dx = dx.contiguous()
        b, c, l = _dims(y)
        dev = y.device
        dy = torch.empty_like(y)
        dgamma = torch.empty_like(weight)
        dbeta = torch.empty_like(bias)
        sums = torch.empty((2 * c,), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            N.call("bqa_bn_relu_backward", b, c, l, N.ptr(dx), N.ptr(y), N.ptr(mean), N.ptr(invstd),
                   N.ptr(weight), N.ptr(bias), N.ptr(sums), N.ptr(dy), N.ptr(dgamma), N.ptr(dbeta),
                   N.stream_ptr(dev))
        return dy, dgamma, dbeta, None


class _BNReLUMax(Function):
    @staticmethod
    def forward(ctx, y, weight, bias, norm):
        y = y.contiguous()                         # (B, C, npoint, nsample)
        mean, invstd = _stats(y, norm)
        b, c, npoint, ns = y.shape
        out = torch.empty((b, c, npoint), dtype=_f32, device=y.device)
        argmax = torch.empty((b, c, npoint), dtype=torch.int32, device=y.device)
        with torch.cuda.device(y.device):
            N.call("bqa_bn_relu_max_forward", b, c, npoint, ns, N.ptr(y), N.ptr(mean), N.ptr(invstd),
                   N.ptr(weight), N.ptr(bias), N.ptr(out), N.ptr(argmax), N.stream_ptr(y.device))
        ctx.save_for_backward(y, weight, bias, mean, invstd, argmax)
        ctx.mark_non_differentiable(argmax)
        return out, argmax

    @staticmethod
    def backward(ctx, dout, _dargmax=None):
        y, weight, bias, mean, invstd, argmax = ctx.saved_tensors
        dout = dout.contiguous()
        b, c, npoint, ns = y.shape
        dev = y.device
        dy = torch.empty_like(y)
        dgamma = torch.empty_like(weight)
        dbeta = torch.empty_like(bias)
        sums = torch.empty((2 * c,), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            N.call("bqa_bn_relu_max_backward", b, c, npoint, ns, N.ptr(dout), N.ptr(argmax), N.ptr(y),
                   N.ptr(mean), N.ptr(invstd), N.ptr(weight), N.ptr(bias), N.ptr(sums), N.ptr(dy),
                   N.ptr(dgamma), N.ptr(dbeta), N.stream_ptr(dev))
        return dy, dgamma, dbeta, None


def norm_supported(norm, y):
    return (enabled() and norm.training and y.is_cuda and y.dtype == _f32 and norm.affine
            and norm.momentum is not None and y.dim() >= 3 and y.numel() > 0
            and y.numel() // y.shape[1] > 1 and y.shape[0] <= 65535 and y.shape[1] <= 65535)


def bn_relu(y, norm):
    """relu(batch_norm(y)) in training mode; y (B, C, ...)."""
    return _BNReLU.apply(y, norm.weight, norm.bias, norm)


def max_supported(y):
    # (the kernels index one (b, c) row per blockIdx.y)
    return (y.dim() == 4 and y.shape[0] * y.shape[1] <= 65535
            and bool(N.lib().bqa_bn_relu_max_supported(int(y.shape[3]))))


def bn_relu_max(y, norm):
    """max over the last axis of relu(batch_norm(y)); y (B, C, npoint, nsample) -> (B, C, npoint)."""
    return _BNReLUMax.apply(y, norm.weight, norm.bias, norm)[0]
