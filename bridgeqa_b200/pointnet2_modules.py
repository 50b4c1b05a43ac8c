"""Set-abstraction / feature-propagation layers -- drop-in for the reference's
`lib/pointnet2/pointnet2_modules.py` (same class names, keyword-only constructors,
forward signatures, return tuples and state_dict keys).

Used by BridgeQA: PointnetSAModuleVotes (pointnet2_modules.py:164-277) and
PointnetFPModule (:361-421).  The other classes of the reference file
(_PointnetSAModuleBase, PointnetSAModuleMSG, PointnetSAModule, PointnetSAModuleMSGVotes,
PointnetLFPModuleMSG) are kept working on the un-fused operators.

Two execution paths, chosen per call:
  * un-fused  -- FPS(+gather) -> ball query -> grouping -> SharedMLP (torch) -> max-pool.
    Differentiable; used for training and as the structure the parity tests compare.
  * fused     -- eval mode, no autograd: FPS(+gather) -> ball query -> ONE kernel that
    gathers neighbours, runs the three folded conv+BN+ReLU layers on tensor cores and
    max-pools over nsample, never materialising the (B, C+3, npoint, nsample) tensor.
    Enabled with `bridgeqa_b200.set_fused(True)` (default on when the kernel is built).
"""
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import pointnet2_utils
from . import pytorch_utils as pt_utils
from . import fused


def _centres(xyz, npoint, inds):
    """(inds, new_xyz) for the sampled centres.  When nothing needs a gradient the
    coordinates come out of the FPS kernel's epilogue; otherwise the reference's
    gather_operation(xyz^T, inds)^T chain is used so xyz receives its gradient
    (pointnet2_modules.py:233-240)."""
    if inds is None and not (torch.is_grad_enabled() and xyz.requires_grad):
        return pointnet2_utils.furthest_point_sample_with_xyz(xyz, npoint)
    if inds is None:
        inds = pointnet2_utils.furthest_point_sample(xyz, npoint)
    new_xyz = pointnet2_utils.gather_operation(
        xyz.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()
    return inds, new_xyz


def _pool_max(x):
    # max over the nsample axis, (B,C,npoint,nsample) -> (B,C,npoint)
    return F.max_pool2d(x, kernel_size=[1, x.size(3)]).squeeze(-1)


def _make_scales(npoint, radii, nsamples, mlps, bn, use_xyz, sample_uniformly):
    groupers, nets = nn.ModuleList(), nn.ModuleList()
    for radius, nsample, spec in zip(radii, nsamples, mlps):
        if npoint is not None:
            groupers.append(pointnet2_utils.QueryAndGroup(
                radius, nsample, use_xyz=use_xyz, sample_uniformly=sample_uniformly))
        else:
            groupers.append(pointnet2_utils.GroupAll(use_xyz))
        if use_xyz:
            spec[0] += 3          # in place, like the reference (callers see the +3)
        nets.append(pt_utils.SharedMLP(spec, bn=bn))
    return groupers, nets


class _PointnetSAModuleBase(nn.Module):
    def __init__(self):
        super().__init__()
        self.npoint = None
        self.groupers = None
        self.mlps = None

    def _multi_scale(self, xyz, new_xyz, features):
        outs = [mlp.forward_pooled(grouper(xyz, new_xyz, features))
                for grouper, mlp in zip(self.groupers, self.mlps)]
        return torch.cat(outs, dim=1)

    def forward(self, xyz, features=None):
        """xyz (B,N,3), features (B,C,N) -> new_xyz (B,npoint,3), new_features (B,sum mlp[-1],npoint)"""
        new_xyz = _centres(xyz, self.npoint, None)[1] if self.npoint is not None else None
        return new_xyz, self._multi_scale(xyz, new_xyz, features)


class PointnetSAModuleMSG(_PointnetSAModuleBase):
    """Multi-scale grouping SA layer (pointnet2_modules.py:78-125)."""

    def __init__(self, *, npoint: int, radii: List[float], nsamples: List[int],
                 mlps: List[List[int]], bn: bool = True, use_xyz: bool = True,
                 sample_uniformly: bool = False):
        super().__init__()
        assert len(radii) == len(nsamples) == len(mlps)
        self.npoint = npoint
        self.groupers, self.mlps = _make_scales(npoint, radii, nsamples, mlps, bn, use_xyz,
                                                sample_uniformly)


class PointnetSAModule(PointnetSAModuleMSG):
    """Single-scale SA layer (pointnet2_modules.py:128-161)."""

    def __init__(self, *, mlp: List[int], npoint: int = None, radius: float = None,
                 nsample: int = None, bn: bool = True, use_xyz: bool = True):
        super().__init__(mlps=[mlp], npoint=npoint, radii=[radius], nsamples=[nsample], bn=bn,
                         use_xyz=use_xyz)


class PointnetSAModuleVotes(nn.Module):
    """SA layer that also returns the sampled indices (VoteNet needs them for the GT
    votes).  forward(xyz (B,N,3), features (B,C,N) | None, inds (B,npoint) int32 | None)
    -> (new_xyz (B,npoint,3), new_features (B,mlp[-1],npoint), inds[, unique_cnt])."""

    def __init__(self, *, mlp: List[int], npoint: int = None, radius: float = None,
                 nsample: int = None, bn: bool = True, use_xyz: bool = True,
                 pooling: str = 'max', sigma: float = None, normalize_xyz: bool = False,
                 sample_uniformly: bool = False, ret_unique_cnt: bool = False):
        super().__init__()
        self.npoint = npoint
        self.radius = radius
        self.nsample = nsample
        self.pooling = pooling
        self.use_xyz = use_xyz
        self.sigma = sigma if sigma is not None else (radius / 2 if radius is not None else None)
        self.normalize_xyz = normalize_xyz
        self.sample_uniformly = sample_uniformly
        self.ret_unique_cnt = ret_unique_cnt
        if npoint is not None:
            self.grouper = pointnet2_utils.QueryAndGroup(
                radius, nsample, use_xyz=use_xyz, ret_grouped_xyz=True,
                normalize_xyz=normalize_xyz, sample_uniformly=sample_uniformly,
                ret_unique_cnt=ret_unique_cnt)
        else:
            self.grouper = pointnet2_utils.GroupAll(use_xyz, ret_grouped_xyz=True)
        mlp_spec = mlp
        if use_xyz and len(mlp_spec) > 0:
            mlp_spec[0] += 3
        self.mlp_module = pt_utils.SharedMLP(mlp_spec, bn=bn)
        self._fused_cache = None

    # -- fused inference path -----------------------------------------------------
    def _can_fuse(self, xyz, features):
        return (fused.enabled() and self.npoint is not None and self.pooling == 'max'
                and self.use_xyz and not self.sample_uniformly and not self.training
                and not torch.is_grad_enabled() and xyz.is_cuda
                and fused.sa_supported(self.mlp_module, self.nsample, self.npoint,
                                       0 if features is None else features.size(1)))

    def train(self, mode=True):
        self._fused_cache = None       # folded weights are stale once BN stats can move
        return super().train(mode)

    def _packed(self):
        sig = fused.weights_signature(self.mlp_module)
        if self._fused_cache is None or self._fused_cache[0] != sig:
            self._fused_cache = (sig, fused.fold_sa_mlp(self.mlp_module))
        return self._fused_cache[1]

    def forward(self, xyz, features=None, inds=None, new_xyz=None, grid=None):
        """`new_xyz` (optional, beyond the reference signature): the centres xyz[inds] when the
        caller already has them (e.g. from the FPS kernel's epilogue), skipping the gather.
        `grid` (optional): fused.prebuild_ball_query_grid(xyz, radius), built while sampling."""
        if inds is not None:
            assert inds.shape[1] == self.npoint
        if self.npoint is None:
            new_xyz = None
        elif inds is None or new_xyz is None:
            inds, new_xyz = _centres(xyz, self.npoint, inds)

        if self._can_fuse(xyz, features):
            new_features = fused.sa_forward(xyz, new_xyz, features, self.radius, self.nsample,
                                            self.normalize_xyz, self._packed(), grid=grid)
            return new_xyz, new_features, inds

        if getattr(features, "_bqa_staged", False):
            raise RuntimeError("a staged 16-bit cloud is only accepted by the fused inference path")
        if grid is not None and isinstance(self.grouper, pointnet2_utils.QueryAndGroup):
            grouped = self.grouper(xyz, new_xyz, features, grid=grid)
        else:
            grouped = self.grouper(xyz, new_xyz, features)
        unique_cnt = grouped[2] if self.ret_unique_cnt else None
        grouped_features, grouped_xyz = grouped[0], grouped[1]

        if self.pooling == 'max':
            # (B, mlp[-1], npoint, nsample) -> max over nsample, pointnet2_modules.py:259-262
            new_features = self.mlp_module.forward_pooled(grouped_features)
            if self.ret_unique_cnt:
                return new_xyz, new_features, inds, unique_cnt
            return new_xyz, new_features, inds
        new_features = self.mlp_module(grouped_features)       # (B, mlp[-1], npoint, nsample)
        if self.pooling == 'avg':
            new_features = F.avg_pool2d(new_features, kernel_size=[1, new_features.size(3)]).squeeze(-1)
        elif self.pooling == 'rbf':
            # RBF-weighted sum over the ball, normalised by nsample (pointnet2_modules.py:268-271)
            rbf = torch.exp(-1 * grouped_xyz.pow(2).sum(1, keepdim=False) / (self.sigma ** 2) / 2)
            new_features = torch.sum(new_features * rbf.unsqueeze(1), -1) / float(self.nsample)
        else:
            raise ValueError("unknown pooling %r" % (self.pooling,))

        if self.ret_unique_cnt:
            return new_xyz, new_features, inds, unique_cnt
        return new_xyz, new_features, inds


class PointnetSAModuleMSGVotes(nn.Module):
    """Multi-scale variant that returns inds (pointnet2_modules.py:279-358)."""

    def __init__(self, *, mlps: List[List[int]], npoint: int, radii: List[float],
                 nsamples: List[int], bn: bool = True, use_xyz: bool = True,
                 sample_uniformly: bool = False):
        super().__init__()
        assert len(mlps) == len(nsamples) == len(radii)
        self.npoint = npoint
        self.groupers, self.mlps = _make_scales(npoint, radii, nsamples, mlps, bn, use_xyz,
                                                sample_uniformly)

    def forward(self, xyz, features=None, inds=None):
        if self.npoint is not None:
            inds, new_xyz = _centres(xyz, self.npoint, inds)
        else:
            new_xyz = None
        outs = [mlp.forward_pooled(grouper(xyz, new_xyz, features))
                for grouper, mlp in zip(self.groupers, self.mlps)]
        return new_xyz, torch.cat(outs, dim=1), inds


class PointnetFPModule(nn.Module):
    """Feature propagation: inverse-distance interpolation from `known` to `unknown`,
    concat with the skip features, SharedMLP (pointnet2_modules.py:361-421).

    forward(unknown (B,n,3), known (B,m,3), unknow_feats (B,C1,n), known_feats (B,C2,m))
      -> (B, mlp[-1], n)"""

    def __init__(self, *, mlp: List[int], bn: bool = True):
        super().__init__()
        self.mlp = pt_utils.SharedMLP(mlp, bn=bn)
        self._fused_cache = None
        # fused inference path: also emit the point-major fp32 copy of the output that a following
        # fused FP layer interpolates from (the last FP layer of a backbone can switch it off)
        self.emit_point_major = True

    def train(self, mode=True):
        self._fused_cache = None
        return super().train(mode)

    def forward(self, unknown, known, unknow_feats, known_feats):
        if (known is not None and unknow_feats is not None and fused.enabled()
                and not self.training and not torch.is_grad_enabled() and unknown.is_cuda
                and fused.fp_supported(self.mlp, unknown.size(1), known.size(1),
                                       known_feats.size(1), unknow_feats.size(1))):
            sig = fused.weights_signature(self.mlp)
            if self._fused_cache is None or self._fused_cache[0] != sig:
                self._fused_cache = (sig, fused.fold_fp_mlp(self.mlp))
            return fused.fp_forward(unknown, known, unknow_feats, known_feats, self._fused_cache[1],
                                    pm32=self.emit_point_major)

        if known is not None:
            dist, idx = pointnet2_utils.three_nn(unknown, known)
            dist_recip = 1.0 / (dist + 1e-8)
            weight = dist_recip / torch.sum(dist_recip, dim=2, keepdim=True)
            interpolated = pointnet2_utils.three_interpolate(known_feats, idx, weight)
        else:
            interpolated = known_feats.expand(*known_feats.size()[0:2], unknown.size(1))
        if unknow_feats is not None:
            interpolated = torch.cat([interpolated, unknow_feats], dim=1)   # (B, C2+C1, n)
        return self.mlp(interpolated.unsqueeze(-1)).squeeze(-1)


class PointnetLFPModuleMSG(nn.Module):
    """Learnable feature propagation (pointnet2_modules.py:423-501): group features1 around
    xyz2, per-scale SharedMLP + max-pool, concat features2, post-MLP."""

    def __init__(self, *, mlps: List[List[int]], radii: List[float], nsamples: List[int],
                 post_mlp: List[int], bn: bool = True, use_xyz: bool = True,
                 sample_uniformly: bool = False):
        super().__init__()
        assert len(mlps) == len(nsamples) == len(radii)
        self.post_mlp = pt_utils.SharedMLP(post_mlp, bn=bn)
        self.groupers, self.mlps = _make_scales(0, radii, nsamples, mlps, bn, use_xyz,
                                                sample_uniformly)

    def forward(self, xyz2, xyz1, features2, features1):
        outs = []
        for grouper, mlp in zip(self.groupers, self.mlps):
            f = _pool_max(mlp(grouper(xyz1, xyz2, features1)))          # (B, mlp[-1], N2)
            if features2 is not None:
                f = torch.cat([f, features2], dim=1)
            outs.append(self.post_mlp(f.unsqueeze(-1)))
        return torch.cat(outs, dim=1).squeeze(-1)
