"""Host side of the fused inference kernels (SA: gather + 3x[conv+BN+ReLU] + max-pool on
tcgen05 tensor cores).

The module classes call in here only in eval mode with autograd off; training always
takes the un-fused, differentiable operators.  Nothing here computes on the host: folding
BN into the convolution weights is a handful of tiny torch device ops done once per module
(cached until .train() is called), packing to the kernel's bf16 shared-memory image is a
kernel of the C ABI.
"""
import contextlib
import ctypes
import os
import threading

import torch

from . import _native as N
from . import ext as _ext
from . import pytorch_utils as pt_utils

_state = {"enabled": True, "precision": "fp16"}
_PRECISIONS = {"bf16": 0, "fp16": 1}
_f32 = torch.float32


def set_fused(flag):
    """Globally allow (default) or forbid the fused inference kernels."""
    _state["enabled"] = bool(flag)


def enabled():
    return _state["enabled"]


def set_precision(name):
    """Operand format of the fused tensor-core kernels: "fp16" (default; 11-bit mantissa like
    TF32, the precision class of the reference's cuDNN convs; saturates at 65504) or "bf16"
    (8-bit mantissa, fp32 range).  Accumulation is always fp32."""
    if name not in _PRECISIONS:
        raise ValueError("precision must be one of %s" % sorted(_PRECISIONS))
    _state["precision"] = name


def precision():
    return _state["precision"]


def _widths(mlp_module):
    blocks = list(mlp_module)
    if len(blocks) != 3:
        return None
    ws = []
    for blk in blocks:
        conv = getattr(blk, "conv", None)
        if conv is None or not isinstance(conv, torch.nn.Conv2d) or conv.kernel_size != (1, 1):
            return None
        if not isinstance(getattr(blk, "activation", None), torch.nn.ReLU):
            return None
        ws.append((conv.in_channels, conv.out_channels))
    if ws[0][1] != ws[1][0] or ws[1][1] != ws[2][0]:
        return None
    return ws


def sa_supported(mlp_module, nsample, npoint, c_feat):
    ws = _widths(mlp_module)
    if ws is None or ws[0][0] != c_feat + 3:
        return False
    return bool(N.lib().bqa_sa_mlp_max_supported(int(nsample), int(npoint), int(c_feat),
                                                ws[0][1], ws[1][1], ws[2][1]))


def _fp_widths(mlp_module):
    blocks = list(mlp_module)
    if len(blocks) != 2:
        return None
    ws = []
    for blk in blocks:
        conv = getattr(blk, "conv", None)
        if conv is None or not isinstance(conv, torch.nn.Conv2d) or conv.kernel_size != (1, 1):
            return None
        if not isinstance(getattr(blk, "activation", None), torch.nn.ReLU):
            return None
        ws.append((conv.in_channels, conv.out_channels))
    return ws if ws[0][1] == ws[1][0] else None


def fp_supported(mlp_module, n, m, c_known, c_skip):
    ws = _fp_widths(mlp_module)
    if ws is None or ws[0][0] != c_known + c_skip:
        return False
    return bool(N.lib().bqa_fp_mlp_supported(int(n), int(m), int(c_known), int(c_skip), ws[0][1], ws[1][1]))


class PackedSA(object):
    """bf16 weight images + fp32 biases of one folded 3-layer SharedMLP, on one device."""

    def __init__(self, mlp_module):
        self.precision = _PRECISIONS[_state["precision"]]
        ws = _widths(mlp_module)
        self.c_in = ws[0][0]
        self.c1, self.c2, self.c3 = ws[0][1], ws[1][1], ws[2][1]
        self.packed, self.bias = [], []
        for li, blk in enumerate(mlp_module):
            w, b = pt_utils.fold_conv_bn(blk)
            c_out, c_in = w.shape
            kpad = (c_in + 15) // 16 * 16
            img = torch.empty(c_out * kpad, dtype=torch.int16, device=w.device)
            with torch.cuda.device(w.device):
                N.call("bqa_pack_weight_16", c_out, c_in, kpad, 1 if li == 0 else 0, self.precision, N.ptr(w),
                       N.ptr(img), N.stream_ptr(w.device))
            self.packed.append(img)
            self.bias.append(b)
        self.device = self.packed[0].device
        # images of the warp-specialised kernel (sa_fused_v2.cu): biases of layers 1-2 folded into K
        self.packed_v2 = []
        folded = [pt_utils.fold_conv_bn(blk) for blk in mlp_module]
        c = self.c_in - 3
        for li, (w, b) in enumerate(folded):
            c_out, c_in = w.shape
            kpad = ((c + 5 + 15) // 16 * 16, c_in + 16, c_in)[li]
            img = torch.empty(c_out * kpad, dtype=torch.int16, device=w.device)
            with torch.cuda.device(w.device):
                N.call("bqa_pack_weight_16_v2", c_out, c_in, kpad, (1, 2, 0)[li], self.precision, N.ptr(w),
                       N.ptr(b.contiguous()), N.ptr(img), N.stream_ptr(w.device))
            self.packed_v2.append(img)


def weights_signature(module):
    """Changes whenever a parameter / BN buffer of `module` is replaced, moved or written in
    place (load_state_dict, optimizer step, .to()), so folded weights are never stale."""
    return (_state["precision"],) + tuple(
        (t.data_ptr(), t._version) for t in list(module.parameters()) + list(module.buffers()))


def fold_sa_mlp(mlp_module):
    return PackedSA(mlp_module)


class PackedFP(object):
    """Layer-1 and layer-2 weight images back to back (the kernel streams them as 32 KB
    slices) + fp32 biases of one folded 2-layer SharedMLP."""

    def __init__(self, mlp_module):
        self.precision = _PRECISIONS[_state["precision"]]
        ws = _fp_widths(mlp_module)
        self.c1, self.c2 = ws[0][1], ws[1][1]
        folded = [pt_utils.fold_conv_bn(blk) for blk in mlp_module]
        sizes = [w.shape[0] * w.shape[1] for w, _ in folded]
        dev = folded[0][0].device
        self.image = torch.empty(sum(sizes), dtype=torch.int16, device=dev)
        off = 0
        with torch.cuda.device(dev):
            for (w, _), sz in zip(folded, sizes):
                c_out, c_in = w.shape
                N.call("bqa_pack_weight_16", c_out, c_in, c_in, 0, self.precision, N.ptr(w),
                       ctypes.c_void_p(self.image.data_ptr() + 2 * off), N.stream_ptr(dev))
                off += sz
        self.bias = [b for _, b in folded]


def fold_fp_mlp(mlp_module):
    return PackedFP(mlp_module)


def point_major(features):
    """(B,C,N) channel-major features -> a (B,N,C)-indexable fp32 view with unit channel stride.
    Layers that produced `features` attach the point-major twin they already have
    (`_bqa_pm`), so no transpose kernel runs between fused layers."""
    if getattr(features, "_bqa_staged", False):
        raise RuntimeError("a staged 16-bit cloud carries no fp32 features: only the fused SA kernel "
                           "(sa_fused_v2) can consume it")
    pm = getattr(features, "_bqa_pm", None)
    if (pm is not None and pm.dtype == _f32 and pm.device == features.device and pm.dim() == 3
            and pm.size(0) == features.size(0) and pm.size(1) == features.size(2)
            and pm.size(2) == features.size(1) and pm.stride(2) == 1
            and pm.stride(0) == pm.size(1) * pm.stride(1)):
        return pm
    return _ext.transpose_to_point_major(features.contiguous())


def point_major_16(features):
    """(B,C,N) channel-major features -> ((B,N,stride) 16-bit point-major tensor in the current
    operand precision, rows zero padded to `stride` = roundup8(C)).  Fused SA layers attach the twin
    they wrote (`_bqa_pm16`); the backbone's input features carry the (B,N,3+C) cloud they are a view
    of (`_bqa_cloud`), which is converted row-wise without a transpose; anything else goes through
    one transpose-convert kernel."""
    prec = _PRECISIONS[_state["precision"]]
    pm = getattr(features, "_bqa_pm16", None)
    if (pm is not None and getattr(features, "_bqa_pm16_prec", None) == prec and pm.device == features.device
            and pm.dim() == 3 and pm.size(0) == features.size(0) and pm.size(1) == features.size(2)
            and pm.size(2) >= features.size(1) and pm.is_contiguous()):
        return pm
    b, c, n = features.shape
    stride = (c + 7) // 8 * 8
    out = torch.empty((b, n, stride), dtype=torch.int16, device=features.device)
    cloud = getattr(features, "_bqa_cloud", None)
    with torch.cuda.device(features.device):
        if (cloud is not None and cloud.is_contiguous() and cloud.dtype == _f32 and cloud.device == features.device
                and cloud.dim() == 3 and cloud.size(0) == b and cloud.size(1) == n and cloud.size(2) == c + 3):
            N.call("bqa_rows_to_16", ctypes.c_longlong(b * n), c, c + 3, 3, stride, prec, N.ptr(cloud), N.ptr(out),
                   N.stream_ptr(features.device))
        else:
            f = features.contiguous()
            N.check_tensor(f, "features", _f32)
            N.call("bqa_to_point_major_16", b, c, n, stride, prec, N.ptr(f), N.ptr(out),
                   N.stream_ptr(features.device))
    return out


def sa_v2_enabled():
    """The warp-specialised SA kernel (sa_fused_v2.cu) is the default; BQA_SA_V2=0 selects the
    round-1 kernel (sa_fused.cu) for A/B measurements."""
    return os.environ.get("BQA_SA_V2", "1") != "0"


GRID_MIN_POINTS = 256      # prebuilt grids: below this a scan is as fast as the search


def prebuild_ball_query_grid(xyz, radius, after=None, inline=False):
    """Bin `xyz` (B,N,3) into the ball query's cell grid.  By default on a side stream -- the
    build only needs the coordinates, so it runs underneath the sampling that produces the
    centres; `after`: event that marks xyz complete when another stream produced it (default:
    the current stream's position).  inline=True builds on the current stream instead (the
    sampling itself is going to use the grid: furthest_point_sample_grid).
    Returns (grid, event | None) for sa_forward(..., grid=...), or None for scenes too small
    to bother."""
    N.check_tensor(xyz, "xyz", _f32)
    b, n, _ = xyz.shape
    if n < GRID_MIN_POINTS:
        return None
    dev = xyz.device
    grid = torch.empty((N.lib().bqa_ball_query_grid_bytes(b, n),), dtype=torch.uint8, device=dev)
    if inline:
        with torch.cuda.device(dev):
            N.call("bqa_ball_query_grid_build", b, n, ctypes.c_float(radius), N.ptr(xyz), N.ptr(grid),
                   N.stream_ptr(dev))
        return grid, None
    main = torch.cuda.current_stream(dev)
    side = side_stream(dev, "bq_grid")
    ready = after
    if ready is None:
        ready = torch.cuda.Event()
        ready.record(main)
    with torch.cuda.device(dev), torch.cuda.stream(side):
        side.wait_event(ready)
        N.call("bqa_ball_query_grid_build", b, n, ctypes.c_float(radius), N.ptr(xyz), N.ptr(grid),
               N.stream_ptr(dev))
        built = torch.cuda.Event()
        built.record(side)
    grid.record_stream(side)
    xyz.record_stream(side)
    return grid, built


def fps_grid_supported(n, npoint):
    return bool(N.lib().bqa_fps_grid_supported(int(n), int(npoint)))


_mode = threading.local()


def lean_sampling_enabled():
    return bool(getattr(_mode, "lean", False))


@contextlib.contextmanager
def lean_sampling(on=True):
    """Within this block furthest_point_sample_grid defaults to the throughput variant of the
    sorted sampling kernel (fps_sorted.cu, kLean): half the SMs per scene, ~35 % more latency.
    graphs.InFlight wraps its forwards in it; a single forward at a time should not."""
    old = lean_sampling_enabled()
    _mode.lean = bool(on)
    try:
        yield
    finally:
        _mode.lean = old


def furthest_point_sample_grid(xyz, npoint, grid, lean=None):
    """(inds, new_xyz) == furthest_point_sample_with_xyz(xyz, npoint), computed over the cell grid
    `grid` = prebuild_ball_query_grid(xyz, ..., inline=True) with warp-level pruning
    (fps_sorted.cu): bit-identical; 7-11 % less kernel time at 20k-100k points on B200.
    lean=True: the throughput variant (coordinates in shared memory, half the SMs per scene,
    more latency) for forwards that run several batches in flight; None = lean_sampling_enabled()."""
    if lean is None:
        lean = lean_sampling_enabled()
    N.check_tensor(xyz, "xyz", _f32)
    b, n, _ = xyz.shape
    m = int(npoint)
    if grid[1] is not None:
        torch.cuda.current_stream(xyz.device).wait_event(grid[1])
    inds = torch.empty((b, m), dtype=torch.int32, device=xyz.device)
    new_xyz = torch.empty((b, m, 3), dtype=_f32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        N.call("bqa_furthest_point_sampling_grid_lean" if lean else "bqa_furthest_point_sampling_grid",
               b, n, m, N.ptr(xyz), N.ptr(grid[0]), N.ptr(inds),
               N.ptr(new_xyz), N.stream_ptr(xyz.device))
    return inds, new_xyz


def ball_query_on_grid(new_xyz, xyz, radius, nsample, grid):
    """ball_query(new_xyz, xyz, radius, nsample) over a prebuilt cell grid (buffer, event | None)."""
    N.check_tensor(xyz, "xyz", _f32)
    N.check_tensor(new_xyz, "new_xyz", _f32)
    b, n, _ = xyz.shape
    npoint = new_xyz.size(1)
    if grid[1] is not None:
        torch.cuda.current_stream(xyz.device).wait_event(grid[1])
    idx = torch.empty((b, npoint, int(nsample)), dtype=torch.int32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        N.call("bqa_ball_query_grid_search", b, n, npoint, 0, npoint, ctypes.c_float(radius),
               int(nsample), N.ptr(new_xyz), N.ptr(xyz), N.ptr(idx), N.ptr(grid[0]),
               N.stream_ptr(xyz.device))
    return idx


def sa_forward(xyz, new_xyz, features, radius, nsample, normalize_xyz, packed, grid=None, pm32=True):
    """-> new_features (B, C3, npoint) fp32, with point-major twins attached: ._bqa_pm16 (16-bit, what
    the next fused SA layer gathers from) and, when pm32, ._bqa_pm (fp32, what the fused FP layers
    interpolate from).  `grid`: optional (buffer, event) from prebuild_ball_query_grid(xyz, ...)."""
    N.check_tensor(xyz, "xyz", _f32)
    N.check_tensor(new_xyz, "new_xyz", _f32)
    b, n, _ = xyz.shape
    npoint = new_xyz.size(1)
    if grid is not None:
        idx = ball_query_on_grid(new_xyz, xyz, radius, nsample, grid)
    else:
        idx = _ext.ball_query(new_xyz, xyz, radius, nsample)
    c = features.size(1) if features is not None else 0
    out_cm = torch.empty((b, packed.c3, npoint), dtype=_f32, device=xyz.device)
    if sa_v2_enabled() and N.lib().bqa_sa_mlp_max_v2_supported(int(nsample), npoint, c, packed.c1, packed.c2,
                                                               packed.c3):
        pm16 = point_major_16(features) if features is not None else None
        out_pm = torch.empty((b, npoint, packed.c3), dtype=_f32, device=xyz.device) if pm32 else None
        out_pm16 = torch.empty((b, npoint, packed.c3), dtype=torch.int16, device=xyz.device)
        with torch.cuda.device(xyz.device):
            N.call("bqa_sa_mlp_max_forward_v2", b, n, npoint, int(nsample), c, N.ptr(xyz), N.ptr(new_xyz),
                   N.ptr(pm16), pm16.size(2) if pm16 is not None else 0, N.ptr(idx), ctypes.c_float(radius),
                   1 if normalize_xyz else 0, packed.c1, packed.c2, packed.c3, N.ptr(packed.packed_v2[0]),
                   N.ptr(packed.packed_v2[1]), N.ptr(packed.packed_v2[2]), N.ptr(packed.bias[2]),
                   N.ptr(out_cm), N.ptr(out_pm), N.ptr(out_pm16), packed.precision, N.stream_ptr(xyz.device))
        if out_pm is not None:
            out_cm._bqa_pm = out_pm
        out_cm._bqa_pm16, out_cm._bqa_pm16_prec = out_pm16, packed.precision
        return out_cm
    if features is not None:
        pm = point_major(features)
        c, stride = pm.size(2), pm.stride(1)
    else:
        pm, c, stride = None, 0, 0
    out_pm = torch.empty((b, npoint, packed.c3), dtype=_f32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        N.call("bqa_sa_mlp_max_forward", b, n, npoint, int(nsample), c, N.ptr(xyz), N.ptr(new_xyz),
               N.ptr(pm), stride, N.ptr(idx), ctypes.c_float(radius), 1 if normalize_xyz else 0,
               packed.c1, packed.c2, packed.c3,
               N.ptr(packed.packed[0]), N.ptr(packed.bias[0]), N.ptr(packed.packed[1]),
               N.ptr(packed.bias[1]), N.ptr(packed.packed[2]), N.ptr(packed.bias[2]),
               N.ptr(out_cm), N.ptr(out_pm), packed.precision, N.stream_ptr(xyz.device))
    out_cm._bqa_pm = out_pm
    return out_cm


def fps_prefix_check(xyz, npoint):
    """run_flags (B,) int32 for a cloud that should already be in sampling order (the previous
    level's centres): 0 = re-sampling any prefix of it to <= npoint points provably returns the
    identity prefix, 1 = that scene needs the real chain (bqa_fps_prefix_check)."""
    N.check_tensor(xyz, "xyz", _f32)
    b, n, _ = xyz.shape
    flags = torch.empty((b,), dtype=torch.int32, device=xyz.device)
    scratch = torch.empty((b, int(npoint)), dtype=_f32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        N.call("bqa_fps_prefix_check", b, n, int(npoint), N.ptr(xyz), N.ptr(scratch), N.ptr(flags),
               N.stream_ptr(xyz.device))
    return flags


def furthest_point_sample_cond(xyz, npoint, run_flags):
    """(inds, new_xyz) == furthest_point_sample_with_xyz(xyz, npoint); scenes whose run_flag is 0
    get the (proven) identity prefix without running the serial chain."""
    N.check_tensor(xyz, "xyz", _f32)
    b, n, _ = xyz.shape
    m = int(npoint)
    inds = torch.empty((b, m), dtype=torch.int32, device=xyz.device)
    new_xyz = torch.empty((b, m, 3), dtype=_f32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        nbytes = N.lib().bqa_fps_scratch_bytes(b, n)
        scratch = torch.empty((nbytes // 4,), dtype=_f32, device=xyz.device) if nbytes else None
        N.call("bqa_furthest_point_sampling_cond", b, n, m, N.ptr(xyz), N.ptr(run_flags), N.ptr(inds),
               N.ptr(new_xyz), N.ptr(scratch), N.stream_ptr(xyz.device))
    return inds, new_xyz


_side_streams = {}


def side_stream(device, tag):
    """High-priority side streams: their kernels are short dependencies of the main stream's next
    layer (grid builds, the sampling-order check, lower-level sampling); without priority their
    CTAs queue behind the persistent SA kernels that fill every SM."""
    key = (device, tag)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device, priority=-1)
    return _side_streams[key]


def fp_forward(unknown, known, unknow_feats, known_feats, packed, pm32=True):
    """-> new_features (B, C2, n) fp32, with a point-major fp32 twin attached as ._bqa_pm when pm32 (the
    next FP layer interpolates from it).  The skip features are taken from their 16-bit twin when the
    layer that produced them attached one (bits identical to converting the fp32 values here)."""
    N.check_tensor(unknown, "unknown", _f32)
    N.check_tensor(known, "known", _f32)
    b, n, _ = unknown.shape
    m = known.size(1)
    kpm = point_major(known_feats)
    s16 = getattr(unknow_feats, "_bqa_pm16", None)
    if not (s16 is not None and getattr(unknow_feats, "_bqa_pm16_prec", None) == packed.precision
            and s16.device == unknown.device and s16.is_contiguous() and s16.dim() == 3
            and s16.size(0) == b and s16.size(1) == n and s16.size(2) >= unknow_feats.size(1)
            and s16.size(2) % 8 == 0):
        s16 = None
    spm = point_major(unknow_feats) if s16 is None else None
    c_skip = unknow_feats.size(1)
    out_cm = torch.empty((b, packed.c2, n), dtype=_f32, device=unknown.device)
    out_pm = torch.empty((b, n, packed.c2), dtype=_f32, device=unknown.device) if pm32 else None
    with torch.cuda.device(unknown.device):
        N.call("bqa_fp_mlp_forward", b, n, m, kpm.size(2), c_skip, N.ptr(unknown), N.ptr(known),
               N.ptr(kpm), kpm.stride(1), N.ptr(spm), spm.stride(1) if spm is not None else 0, packed.c1, packed.c2,
               N.ptr(packed.image), N.ptr(packed.bias[0]), N.ptr(packed.bias[1]), N.ptr(out_cm),
               N.ptr(out_pm), packed.precision, N.stream_ptr(unknown.device),
               N.ptr(s16), s16.size(2) if s16 is not None else 0)
    if out_pm is not None:
        out_cm._bqa_pm = out_pm
    return out_cm
