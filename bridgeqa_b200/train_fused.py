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
import os

import torch
from torch.autograd import Function

from . import _native as N

_f32 = torch.float32
_state = {"enabled": True, "conv": os.environ.get("BQA_TRAIN_CONV", "1") != "0"}


def set_enabled(flag):
    _state["enabled"] = bool(flag)


def enabled():
    return _state["enabled"]


def set_conv_enabled(flag):
    """tcgen05 TF32 1x1 convolutions (csrc/conv_tf32.cu) instead of cuDNN in the training path."""
    _state["conv"] = bool(flag)


def conv_enabled():
    """The tcgen05 path computes with TF32 operands, so it follows torch's own switch for TF32 convolutions
    (torch.backends.cudnn.allow_tf32, default True -- the reference's arithmetic): with the switch off the
    convolutions stay on cuDNN in fp32."""
    return _state["conv"] and bool(torch.backends.cudnn.allow_tf32)


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


# ---- 1x1 convolution on tcgen05 (TF32) + BatchNorm statistics from its epilogue ---------------------
# The reference's block is nn.Conv2d(k=1, bias=False) -> BatchNorm2d -> ReLU (pytorch_utils.py:104-157):
# cuDNN conv, then the BatchNorm reads y again for its statistics.  Here the conv kernel's epilogue sums
# y and y^2 per channel while it stores y, so the block is conv(+stats) -> finalize -> apply, ONE autograd
# node whose backward is BatchNorm/ReLU backward -> wgrad -> dgrad, all on the sm_100a kernels.

def _pad_rows(w2):
    """(rows, k) -> contiguous (rows, ld) with ld % 4 == 0 (16-byte aligned rows for the tensor map)."""
    k = w2.shape[1]
    ld = (k + 3) // 4 * 4
    if ld != k:
        w2 = torch.nn.functional.pad(w2, (0, ld - k))
    return w2.contiguous(), ld


def _conv_forward(x, w2, shift=None, sums=None):
    """y (B, Cout, L...) = w2 (Cout, Cin) . x (B, Cin, L...)"""
    b, cin = x.shape[0], x.shape[1]
    cout = w2.shape[0]
    p = x.numel() // max(b * cin, 1)
    wp, ld = _pad_rows(w2)
    y = torch.empty((b, cout) + tuple(x.shape[2:]), dtype=_f32, device=x.device)
    with torch.cuda.device(x.device):
        N.call("bqa_conv1x1_tf32_forward", b, cin, cout, p, N.ptr(x), N.ptr(wp), ld, N.ptr(y), N.ptr(shift),
               N.ptr(sums), N.stream_ptr(x.device))
    return y


def _conv_wgrad(x, dy, cout, cin):
    b = x.shape[0]
    p = x.numel() // max(b * cin, 1)
    dw = torch.zeros((cout, cin), dtype=_f32, device=x.device)
    with torch.cuda.device(x.device):
        N.call("bqa_conv1x1_tf32_wgrad", b, cin, cout, p, N.ptr(x), N.ptr(dy), N.ptr(dw), N.stream_ptr(x.device))
    return dw


def conv_supported(conv, x):
    if not (conv_enabled() and x.is_cuda and x.dtype == _f32 and x.dim() >= 3 and x.numel() > 0):
        return False
    if conv.bias is not None or conv.groups != 1 or conv.weight.dtype != _f32:
        return False
    one = lambda t: all(int(v) == 1 for v in t)
    zero = lambda t: all(int(v) == 0 for v in t)
    if not (one(conv.kernel_size) and one(conv.stride) and one(conv.dilation) and zero(conv.padding)):
        return False
    p = x.numel() // (x.shape[0] * x.shape[1])
    return p % 4 == 0 and conv.out_channels <= 256 and x.shape[1] == conv.in_channels


def _finalize(sums, count, norm, c, dev):
    mean = torch.empty((c,), dtype=_f32, device=dev)
    invstd = torch.empty((c,), dtype=_f32, device=dev)
    track = norm.track_running_stats and norm.running_mean is not None
    with torch.cuda.device(dev):
        N.call("bqa_bn_finalize_shifted", c, ctypes.c_double(count), N.ptr(sums),
               N.ptr(norm.running_mean if track else None), ctypes.c_float(norm.eps),
               ctypes.c_float(norm.momentum), N.ptr(mean), N.ptr(invstd),
               N.ptr(norm.running_mean if track else None), N.ptr(norm.running_var if track else None),
               N.stream_ptr(dev))
    if track and norm.num_batches_tracked is not None:
        norm.num_batches_tracked.add_(1)
    return mean, invstd


class _ConvBNReLU(Function):
    """x -> relu(batch_norm(conv1x1(x))) [-> max over nsample]; pool: 0 = none, 1 = max over the last axis."""

    @staticmethod
    def forward(ctx, x, weight, gamma, beta, norm, pool):
        x = x.contiguous()
        cout, cin = weight.shape[0], weight.shape[1]
        w2 = weight.reshape(cout, cin)
        dev = x.device
        track = norm.track_running_stats and norm.running_mean is not None
        sums = torch.zeros((2 * cout,), dtype=torch.float64, device=dev)
        y = _conv_forward(x, w2, norm.running_mean if track else None, sums)
        b, c, l = _dims(y)
        mean, invstd = _finalize(sums, float(b) * float(l), norm, c, dev)
        ctx.pool = bool(pool)
        if pool:
            npoint, ns = y.shape[2], y.shape[3]
            out = torch.empty((b, c, npoint), dtype=_f32, device=dev)
            argmax = torch.empty((b, c, npoint), dtype=torch.int32, device=dev)
            with torch.cuda.device(dev):
                N.call("bqa_bn_relu_max_forward", b, c, npoint, ns, N.ptr(y), N.ptr(mean), N.ptr(invstd),
                       N.ptr(gamma), N.ptr(beta), N.ptr(out), N.ptr(argmax), N.stream_ptr(dev))
            ctx.save_for_backward(x, weight, y, gamma, beta, mean, invstd, argmax)
            return out
        out = torch.empty_like(y)
        with torch.cuda.device(dev):
            N.call("bqa_bn_relu_forward", b, c, l, N.ptr(y), N.ptr(mean), N.ptr(invstd), N.ptr(gamma),
                   N.ptr(beta), N.ptr(out), N.stream_ptr(dev))
        ctx.save_for_backward(x, weight, y, gamma, beta, mean, invstd)
        return out

    @staticmethod
    def backward(ctx, dout):
        saved = ctx.saved_tensors
        x, weight, y, gamma, beta, mean, invstd = saved[:7]
        dout = dout.contiguous()
        dev = y.device
        b, c, l = _dims(y)
        dy = torch.empty_like(y)
        dgamma = torch.empty_like(gamma)
        dbeta = torch.empty_like(beta)
        sums = torch.empty((2 * c,), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            if ctx.pool:
                npoint, ns = y.shape[2], y.shape[3]
                N.call("bqa_bn_relu_max_backward", b, c, npoint, ns, N.ptr(dout), N.ptr(saved[7]), N.ptr(y),
                       N.ptr(mean), N.ptr(invstd), N.ptr(gamma), N.ptr(beta), N.ptr(sums), N.ptr(dy),
                       N.ptr(dgamma), N.ptr(dbeta), N.stream_ptr(dev))
            else:
                N.call("bqa_bn_relu_backward", b, c, l, N.ptr(dout), N.ptr(y), N.ptr(mean), N.ptr(invstd),
                       N.ptr(gamma), N.ptr(beta), N.ptr(sums), N.ptr(dy), N.ptr(dgamma), N.ptr(dbeta),
                       N.stream_ptr(dev))
        cout, cin = weight.shape[0], weight.shape[1]
        dw = _conv_wgrad(x, dy, cout, cin).reshape(weight.shape) if ctx.needs_input_grad[1] else None
        dx = None
        if ctx.needs_input_grad[0]:
            dx = _conv_forward(dy, weight.reshape(cout, cin).t())       # dx = W^T . dy
        return dx, dw, dgamma, dbeta, None, None


def conv_bn_relu(x, conv, norm):
    """relu(batch_norm(conv(x))) in training mode on the tcgen05 conv kernels."""
    return _ConvBNReLU.apply(x, conv.weight, norm.weight, norm.bias, norm, 0)


def conv_bn_relu_max(x, conv, norm):
    """max over the last axis of relu(batch_norm(conv(x))); x (B, Cin, npoint, nsample)."""
    return _ConvBNReLU.apply(x, conv.weight, norm.weight, norm.bias, norm, 1)


def conv_norm_supported(conv, norm, x):
    """the fused node needs what conv_supported and norm_supported need (the latter on the conv's output shape)"""
    if not (enabled() and conv_supported(conv, x) and norm.training and norm.affine and norm.momentum is not None):
        return False
    b, l = x.shape[0], x.numel() // (x.shape[0] * x.shape[1])
    return b * l > 1 and b <= 65535 and conv.out_channels <= 65535
