"""ctypes binding of libbqa_pointnet2.so (include/bqa_pointnet2.h).

This is the only way the package reaches a GPU kernel: there is no eager / CPU
fallback.  If the shared library is missing or a symbol is absent, import-time use
raises; if a tensor is not a contiguous CUDA tensor of the right dtype the shim raises
RuntimeError just like the reference's CHECK_* macros
(/root/reference/lib/pointnet2/_ext_src/include/utils.h:5-25).
"""
import ctypes
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "lib", "libbqa_pointnet2.so")

_I = ctypes.c_int
_F = ctypes.c_float
_D = ctypes.c_double
_P = ctypes.c_void_p
_LL = ctypes.c_longlong

# name -> argtypes, in header order
_SIGNATURES = {
    "bqa_abi_version": ([], _I),
    "bqa_last_error": ([], ctypes.c_char_p),
    "bqa_launch_count": ([], _LL),
    "bqa_fps_scratch_bytes": ([_I, _I], _LL),
    "bqa_furthest_point_sampling": ([_I, _I, _I, _P, _P, _P, _P, _P], _I),
    "bqa_furthest_point_sampling_slice": ([_I, _I, _I, _I, _I, _P, _P, _P, _P, _I, _P], _I),
    "bqa_group_points_grad_workspace_bytes": ([_I, _I, _I], _LL),
    "bqa_group_points_grad_ws": ([_I, _I, _I, _I, _I, _P, _P, _P, _P, _P], _I),
    "bqa_group_concat_point_major": ([_I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _F, _I, _P, _P], _I),
    "bqa_nms3d": ([_I, _I, _P, _P, ctypes.c_double, _I, _I, _P, _P, _P], _I),
    "bqa_count_points_in_boxes": ([_I, _I, _I, _P, _P, _P, _P], _I),
    "bqa_nn_distance": ([_I, _I, _I, _P, _P, _I, _F, _P, _P, _P, _P, _P], _I),
    "bqa_bn_relu_max_supported": ([_I], _I),
    "bqa_bn_train_stats": ([_I, _I, _LL, _P, _P, _F, _F, _P, _P, _P, _P, _P], _I),
    "bqa_bn_finalize_shifted": ([_I, _D, _P, _P, _F, _F, _P, _P, _P, _P, _P], _I),
    "bqa_conv1x1_tf32_supported": ([_I, _I, _I, _LL, _I], _I),
    "bqa_conv1x1_tf32_forward": ([_I, _I, _I, _LL, _P, _P, _I, _P, _P, _P, _P], _I),
    "bqa_conv1x1_tf32_wgrad": ([_I, _I, _I, _LL, _P, _P, _P, _P], _I),
    "bqa_bn_relu_forward": ([_I, _I, _LL, _P, _P, _P, _P, _P, _P, _P], _I),
    "bqa_bn_relu_max_forward": ([_I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P], _I),
    "bqa_bn_relu_backward": ([_I, _I, _LL, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P], _I),
    "bqa_bn_relu_max_backward": ([_I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P], _I),
    "bqa_fps_grid_supported": ([_I, _I], _I),
    "bqa_furthest_point_sampling_grid": ([_I, _I, _I, _P, _P, _P, _P, _P], _I),
    "bqa_furthest_point_sampling_grid_lean": ([_I, _I, _I, _P, _P, _P, _P, _P], _I),
    "bqa_fps_prefix_check": ([_I, _I, _I, _P, _P, _P, _P], _I),
    "bqa_furthest_point_sampling_cond": ([_I, _I, _I, _P, _P, _P, _P, _P, _P], _I),
    "bqa_gather_points": ([_I, _I, _I, _I, _P, _P, _P, _P], _I),
    "bqa_gather_points_grad": ([_I, _I, _I, _I, _P, _P, _P, _P], _I),
    "bqa_ball_query_workspace_bytes": ([_I, _I, _I, _I], _LL),
    "bqa_ball_query": ([_I, _I, _I, _F, _I, _P, _P, _P, _P, _P], _I),
    "bqa_ball_query_slice": ([_I, _I, _I, _I, _I, _F, _I, _P, _P, _P, _P, _P], _I),
    "bqa_ball_query_grid_bytes": ([_I, _I], _LL),
    "bqa_ball_query_grid_build": ([_I, _I, _F, _P, _P, _P], _I),
    "bqa_ball_query_grid_search": ([_I, _I, _I, _I, _I, _F, _I, _P, _P, _P, _P, _P], _I),
    "bqa_group_points": ([_I, _I, _I, _I, _I, _P, _P, _P, _P], _I),
    "bqa_group_points_grad": ([_I, _I, _I, _I, _I, _P, _P, _P, _P], _I),
    "bqa_three_nn": ([_I, _I, _I, _P, _P, _P, _P, _P], _I),
    "bqa_three_interpolate": ([_I, _I, _I, _I, _P, _P, _P, _P, _P], _I),
    "bqa_three_interpolate_grad": ([_I, _I, _I, _I, _P, _P, _P, _P, _P], _I),
    "bqa_transpose_to_point_major": ([_I, _I, _I, _P, _P, _P], _I),
    "bqa_fp_mlp_supported": ([_I, _I, _I, _I, _I, _I], _I),
    "bqa_fp_mlp_forward": ([_I, _I, _I, _I, _I, _P, _P, _P, _I, _P, _I, _I, _I, _P, _P, _P, _P, _P, _I, _P,
                            _P, _I], _I),
    "bqa_pack_weight_16": ([_I, _I, _I, _I, _I, _P, _P, _P], _I),
    "bqa_sa_mlp_max_supported": ([_I, _I, _I, _I, _I, _I], _I),
    "bqa_pack_weight_16_v2": ([_I, _I, _I, _I, _I, _P, _P, _P, _P], _I),
    "bqa_to_point_major_16": ([_I, _I, _I, _I, _I, _P, _P, _P], _I),
    "bqa_rows_to_16": ([_LL, _I, _I, _I, _I, _I, _P, _P, _P], _I),
    "bqa_sa_mlp_max_v2_supported": ([_I, _I, _I, _I, _I, _I], _I),
    "bqa_sa_mlp_max_forward_v2": ([_I, _I, _I, _I, _I, _P, _P, _P, _I, _P, _F, _I, _I, _I, _I,
                                   _P, _P, _P, _P, _P, _P, _P, _I, _P], _I),
    "bqa_sa_mlp_max_forward": ([_I, _I, _I, _I, _I, _P, _P, _P, _I, _P, _F, _I, _I, _I, _I,
                                _P, _P, _P, _P, _P, _P, _P, _P, _I, _P], _I),
}

_lib = None


def exported_symbols():
    return sorted(_SIGNATURES)


def lib():
    """Load the library (building it first when nvcc and the sources are at hand)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        from . import build as _build
        _build.build()
    if not os.path.exists(SO_PATH):
        raise RuntimeError(
            "bridgeqa_b200: %s is missing; run `python -m bridgeqa_b200.build`. "
            "There is no CPU or eager fallback." % SO_PATH)
    handle = ctypes.CDLL(SO_PATH)
    for name, (argtypes, restype) in _SIGNATURES.items():
        fn = getattr(handle, name)  # AttributeError if the .so is stale: fail loudly
        fn.argtypes = argtypes
        fn.restype = restype
    if handle.bqa_abi_version() != 1:
        raise RuntimeError("bridgeqa_b200: ABI version mismatch in %s" % SO_PATH)
    _lib = handle
    return _lib


def launch_count():
    return int(lib().bqa_launch_count())


def check_tensor(t, name, dtype):
    """utils.h:5-25 -- CHECK_CUDA / CHECK_CONTIGUOUS / CHECK_IS_FLOAT / CHECK_IS_INT."""
    if not isinstance(t, torch.Tensor):
        raise RuntimeError("%s must be a tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (CPU not supported)" % name)
    if not t.is_contiguous():
        raise RuntimeError("%s must be a contiguous tensor" % name)
    if t.dtype != dtype:
        raise RuntimeError("%s must be a %s tensor" % (name, "float" if dtype == torch.float32 else "int"))
    return t


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def call(name, *args):
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        msg = lib().bqa_last_error().decode("utf-8", "replace")
        raise RuntimeError("%s failed (status %d): %s" % (name, rc, msg))
