"""Host side of the fused inference kernels (SA: gather + 3x[conv+BN+ReLU] + max-pool;
FP: three_nn + interpolate + concat + 2x[conv+BN+ReLU]).

The module classes call in here only in eval mode with autograd off; training always
takes the un-fused, differentiable operators.
"""
import torch

from . import _native as N
from . import pytorch_utils as pt_utils

_state = {"enabled": True}


def set_fused(flag):
    """Globally allow (default) or forbid the fused inference kernels."""
    _state["enabled"] = bool(flag)


def enabled():
    return _state["enabled"]


def _has(symbol):
    return symbol in N._SIGNATURES


def sa_supported(mlp_module, nsample, c_feat):
    return False


def fp_supported(mlp_module, c_known, c_skip):
    return False


def fold_sa_mlp(mlp_module):
    return [pt_utils.fold_conv_bn(block) for block in mlp_module]


def fold_fp_mlp(mlp_module):
    return [pt_utils.fold_conv_bn(block) for block in mlp_module]


def sa_forward(xyz, new_xyz, features, radius, nsample, normalize_xyz, folded):
    raise RuntimeError("fused SA kernel not built")


def fp_forward(unknown, known, unknow_feats, known_feats, folded):
    raise RuntimeError("fused FP kernel not built")
