"""Loader-facing input staging.

The reference's loader emits `point_clouds` (B, N, 3 + C) fp32 on the host (lib/dataset.py:372-415) and the
solver copies it to the device as is (lib/solver.py:477-484); the backbone then splits xyz from the features
and transposes them (models/backbone_module.py:74-78).  The fused kernels consume the features as 16-bit
operands anyway (fp16 by default), so a loader can emit them that way and cut the host -> device bytes:

    xyz      (B, N, 3)         fp32   -- coordinates keep every bit: FPS / ball-query indices are bit-exact
    feat16   (B, N, roundup8(C)) 16-bit point-major, zero padded -- exactly the buffer the fused SA1 kernel
                                      gathers from (no transpose, no conversion pass on the device)

    host = staging.stage_host(point_clouds_cpu)          # pinned; the loader's job, once per batch
    dev = staging.StagedCloud.empty_like(host, "cuda")   # device buffers, reused across batches
    dev.copy_(host, non_blocking=True)                   # 28 instead of 40 bytes per point at C = 7
    net({"point_clouds": dev})                           # Pointnet2Backbone / VoteNetDetector, eval mode

The conversion rounds the features ONCE to the operand format, which is the same rounding the fused kernel
would do on the device: the forward is bit-identical to feeding the fp32 cloud (tested).  A StagedCloud is
only accepted by the fused inference path (training needs the fp32 features).
"""
import torch

from . import fused

_f32 = torch.float32


class StagedCloud(object):
    """xyz (B,N,3) fp32 + feat16 (B,N,roundup8(C)) int16 holding `precision` bits; quacks enough like the
    (B,N,3+C) tensor for the graph / in-flight plumbing (device, shape, data_ptr, record_stream, copy_)."""

    def __init__(self, xyz, feat16, c, precision):
        self.xyz, self.feat16, self.c, self.precision = xyz, feat16, int(c), precision
        if xyz.dim() != 3 or xyz.size(2) != 3 or xyz.dtype != _f32:
            raise RuntimeError("xyz must be (B, N, 3) float32")
        if self.c > 0 and (feat16 is None or feat16.dtype != torch.int16 or feat16.dim() != 3
                           or feat16.shape[:2] != xyz.shape[:2] or feat16.size(2) != (self.c + 7) // 8 * 8):
            raise RuntimeError("feat16 must be (B, N, roundup8(C)) int16")

    # ---- the bits of the tensor interface the forward plumbing uses
    is_staged = True
    dtype = _f32

    @property
    def shape(self):
        return torch.Size((self.xyz.size(0), self.xyz.size(1), 3 + self.c))

    def size(self, d=None):
        return self.shape if d is None else self.shape[d]

    def dim(self):
        return 3

    @property
    def device(self):
        return self.xyz.device

    @property
    def is_cuda(self):
        return self.xyz.is_cuda

    def is_contiguous(self):
        return self.xyz.is_contiguous() and (self.feat16 is None or self.feat16.is_contiguous())

    def data_ptr(self):
        # identifies the PAIR of buffers (a graph bound to this input reads both in place)
        return self.xyz.data_ptr() ^ ((self.feat16.data_ptr() * 31) if self.feat16 is not None else 0)

    @property
    def _version(self):
        return self.xyz._version + (self.feat16._version if self.feat16 is not None else 0)

    def record_stream(self, stream):
        self.xyz.record_stream(stream)
        if self.feat16 is not None:
            self.feat16.record_stream(stream)

    def detach(self):
        return self

    def clone(self, memory_format=None):
        return StagedCloud(self.xyz.clone(), None if self.feat16 is None else self.feat16.clone(), self.c,
                           self.precision)

    def contiguous(self):
        return self

    def copy_(self, other, non_blocking=False):
        if not isinstance(other, StagedCloud) or other.shape != self.shape or other.precision != self.precision:
            raise RuntimeError("copy_ needs a StagedCloud of the same shape and precision")
        self.xyz.copy_(other.xyz, non_blocking=non_blocking)
        if self.feat16 is not None:
            self.feat16.copy_(other.feat16, non_blocking=non_blocking)
        return self

    def to(self, device, non_blocking=False):
        return StagedCloud(self.xyz.to(device, non_blocking=non_blocking),
                           None if self.feat16 is None else self.feat16.to(device, non_blocking=non_blocking),
                           self.c, self.precision)

    def cuda(self):
        return self.to("cuda")

    def pin_memory(self):
        return StagedCloud(self.xyz.pin_memory(), None if self.feat16 is None else self.feat16.pin_memory(),
                           self.c, self.precision)

    def nbytes(self):
        return self.xyz.numel() * 4 + (self.feat16.numel() * 2 if self.feat16 is not None else 0)

    @staticmethod
    def empty_like(other, device):
        return StagedCloud(torch.empty_like(other.xyz, device=device),
                           None if other.feat16 is None else torch.empty_like(other.feat16, device=device),
                           other.c, other.precision)

    # ---- what the backbone takes out of it
    def features_view(self):
        """A (B, C, N) stand-in (4 bytes of storage, all strides 0) that carries the 16-bit twin the fused
        SA kernel gathers from; its VALUES must never be read (only the fused path accepts it)."""
        if self.c == 0:
            return None
        b, n = self.xyz.size(0), self.xyz.size(1)
        f = torch.zeros(1, dtype=_f32, device=self.xyz.device).expand(b, self.c, n)
        f._bqa_pm16, f._bqa_pm16_prec, f._bqa_staged = self.feat16, fused._PRECISIONS[self.precision], True
        return f


def stage_host(point_clouds, precision=None, pin=True):
    """(B, N, 3 + C) fp32 CPU tensor -> StagedCloud on the host (pinned by default).  This is loader work
    (one pass over the batch on the host); the rounding is the device kernels' own
    (round-to-nearest-even to fp16, saturating at +-65504, or to bf16)."""
    pc = torch.as_tensor(point_clouds)
    if pc.is_cuda or pc.dtype != _f32 or pc.dim() != 3 or pc.size(2) < 3:
        raise RuntimeError("stage_host takes a (B, N, 3 + C) float32 CPU tensor")
    precision = precision or fused.precision()
    b, n, w = pc.shape
    c = w - 3
    xyz = pc[..., :3].contiguous()
    feat16 = None
    if c > 0:
        stride = (c + 7) // 8 * 8
        f = pc[..., 3:]
        if precision == "fp16":
            h = f.clamp(-65504.0, 65504.0).to(torch.float16)
        else:
            h = f.to(torch.bfloat16)
        buf = torch.zeros((b, n, stride), dtype=h.dtype)
        buf[..., :c] = h
        feat16 = buf.view(torch.int16)
    out = StagedCloud(xyz, feat16, c, precision)
    return out.pin_memory() if pin else out


def read_back(data_dict, keys, out=None, half=(), stream=None):
    """Device -> (pinned) host copies of `keys` of a forward's data_dict on `stream` (default: current).
    Keys listed in `half` are converted to fp16 on the device first (e.g. fp2_features: 16.8 -> 8.4 MB per
    batch of 16).  Returns the dict of host tensors (allocated pinned on first use, pass it back as `out`)."""
    out = {} if out is None else out
    for k in keys:
        t = data_dict[k]
        if k in half:
            t = t.to(torch.float16)
        if k not in out:
            out[k] = torch.empty(t.shape, dtype=t.dtype).pin_memory()
        out[k].copy_(t, non_blocking=True)
    return out
