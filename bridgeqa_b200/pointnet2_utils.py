"""Operator API of the point-cloud hot path -- drop-in for the reference's
`lib/pointnet2/pointnet2_utils.py` (same public names, argument order, return
contracts and autograd behaviour), running on the sm_100a kernels behind
include/bqa_pointnet2.h.

    reference name (pointnet2_utils.py)       here
    furthest_point_sample       :51-80        FurthestPointSampling.apply
    gather_operation            :83-117       GatherOperation.apply
    three_nn                    :120-149      ThreeNN.apply            (returns sqrt'ed distances)
    three_interpolate           :152-206      ThreeInterpolate.apply
    grouping_operation          :209-257      GroupingOperation.apply
    ball_query                  :260-291      BallQuery.apply          (radius, nsample, xyz, new_xyz)
    QueryAndGroup / GroupAll    :294-425      nn.Modules below

Additions that the fused layers use: `furthest_point_sample_with_xyz`.
"""
import torch
import torch.nn as nn
from torch.autograd import Function

from . import ext as _ext


class FurthestPointSampling(Function):
    """xyz (B,N,3) f32, npoint -> (B,npoint) int32 indices; not differentiable."""

    @staticmethod
    def forward(ctx, xyz, npoint):
        inds = _ext.furthest_point_sampling(xyz, npoint)
        ctx.mark_non_differentiable(inds)
        return inds

    @staticmethod
    def backward(ctx, grad=None):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


def furthest_point_sample_with_xyz(xyz, npoint):
    """FPS whose epilogue also emits the sampled coordinates:
    returns (inds (B,npoint) int32, new_xyz (B,npoint,3) f32).  Equivalent to
    furthest_point_sample followed by gather_operation on xyz^T and a transpose
    (pointnet2_modules.py:233-240), in one kernel.  Not differentiable w.r.t. xyz."""
    return _ext.furthest_point_sampling(xyz.detach(), npoint, return_xyz=True)


class GatherOperation(Function):
    """features (B,C,N), idx (B,npoint) int32 -> (B,C,npoint)."""

    @staticmethod
    def forward(ctx, features, idx):
        ctx.saved = (idx, features.size(2))
        return _ext.gather_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        idx, n = ctx.saved
        return _ext.gather_points_grad(grad_out.contiguous(), idx, n), None


gather_operation = GatherOperation.apply


class ThreeNN(Function):
    """unknown (B,n,3), known (B,m,3) -> (dist (B,n,3) euclidean, idx (B,n,3) int32)."""

    @staticmethod
    def forward(ctx, unknown, known):
        dist2, idx = _ext.three_nn(unknown, known)
        ctx.mark_non_differentiable(idx)
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    """features (B,c,m), idx (B,n,3) int32, weight (B,n,3) -> (B,c,n)."""

    @staticmethod
    def forward(ctx, features, idx, weight):
        ctx.saved = (idx, weight, features.size(2))
        return _ext.three_interpolate(features, idx, weight)

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight, m = ctx.saved
        return _ext.three_interpolate_grad(grad_out.contiguous(), idx, weight, m), None, None


three_interpolate = ThreeInterpolate.apply


class GroupingOperation(Function):
    """features (B,C,N), idx (B,npoint,nsample) int32 -> (B,C,npoint,nsample)."""

    @staticmethod
    def forward(ctx, features, idx):
        ctx.saved = (idx, features.size(2))
        return _ext.group_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        idx, n = ctx.saved
        return _ext.group_points_grad(grad_out.contiguous(), idx, n), None


grouping_operation = GroupingOperation.apply


class BallQuery(Function):
    """(radius, nsample, xyz (B,N,3), new_xyz (B,npoint,3)) -> (B,npoint,nsample) int32."""

    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz):
        inds = _ext.ball_query(new_xyz, xyz, radius, nsample)
        ctx.mark_non_differentiable(inds)
        return inds

    @staticmethod
    def backward(ctx, grad=None):
        return None, None, None, None


ball_query = BallQuery.apply


_state = {"group_concat": True}


def set_group_concat(flag):
    """Developer switch: False makes QueryAndGroup always run the reference's op sequence (two
    group_points, subtract, divide, cat) instead of the one-pass kernel."""
    _state["group_concat"] = bool(flag)


class QueryAndGroup(nn.Module):
    """Ball query + grouping, un-fused (this is what training and the parity tests use;
    inference goes through the fused SA kernel and never builds this tensor).

    forward(xyz (B,N,3), new_xyz (B,npoint,3), features (B,C,N) | None)
      -> new_features (B, 3+C, npoint, nsample)            [, grouped_xyz][, unique_cnt]
    """

    def __init__(self, radius, nsample, use_xyz=True, ret_grouped_xyz=False,
                 normalize_xyz=False, sample_uniformly=False, ret_unique_cnt=False):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz
        self.ret_grouped_xyz = ret_grouped_xyz
        self.normalize_xyz = normalize_xyz
        self.sample_uniformly = sample_uniformly
        self.ret_unique_cnt = ret_unique_cnt
        if ret_unique_cnt and not sample_uniformly:
            raise AssertionError("ret_unique_cnt requires sample_uniformly")

    def _resample_uniformly(self, idx):
        # pointnet2_utils.py:336-345: per ball, keep the unique indices and pad by sampling
        # them with replacement (host RNG, like the reference; never enabled by BridgeQA).
        counts = torch.zeros(idx.shape[:2])
        for b in range(idx.size(0)):
            for r in range(idx.size(1)):
                uniq = torch.unique(idx[b, r])
                k = uniq.numel()
                counts[b, r] = k
                pick = torch.randint(0, k, (self.nsample - k,), dtype=torch.long)
                idx[b, r] = torch.cat((uniq, uniq[pick]))
        return counts

    def forward(self, xyz, new_xyz, features=None, grid=None):
        """`grid` (optional, beyond the reference signature): a cell grid already built for `xyz`
        (fused.prebuild_ball_query_grid), e.g. the one the sampling just used."""
        if grid is not None:
            from . import fused
            idx = fused.ball_query_on_grid(new_xyz.detach(), xyz.detach(), self.radius, self.nsample, grid)
        else:
            idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        unique_cnt = self._resample_uniformly(idx) if self.sample_uniformly else None

        # one-pass grouping + centring + cat straight from a point-major twin of the features
        # (the input cloud itself at SA1), when nothing here needs a gradient
        pm = getattr(features, "_bqa_pm", None) if features is not None else None
        if (pm is not None and _state["group_concat"] and self.use_xyz and xyz.is_cuda and pm.is_cuda and pm.dim() == 3
                and pm.size(0) == features.size(0) and pm.size(1) == features.size(2)
                and pm.size(2) == features.size(1) and pm.stride(2) == 1
                and pm.stride(0) == pm.size(1) * pm.stride(1)
                and not (torch.is_grad_enabled() and (xyz.requires_grad or new_xyz.requires_grad
                                                      or features.requires_grad))):
            new_features = _ext.group_concat_point_major(xyz, new_xyz, pm, idx, self.radius,
                                                         self.normalize_xyz)
            out = [new_features]
            if self.ret_grouped_xyz:
                out.append(new_features[:, :3])
            if self.ret_unique_cnt:
                out.append(unique_cnt)
            return out[0] if len(out) == 1 else tuple(out)

        if features is not None and not features.is_contiguous():
            features = features.contiguous()
        grouped_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if self.normalize_xyz:
            grouped_xyz = grouped_xyz / self.radius   # true division, as the reference

        if features is None:
            if not self.use_xyz:
                raise AssertionError("Cannot have not features and not use xyz as a feature!")
            new_features = grouped_xyz
        else:
            grouped = grouping_operation(features, idx)
            new_features = torch.cat([grouped_xyz, grouped], dim=1) if self.use_xyz else grouped

        out = [new_features]
        if self.ret_grouped_xyz:
            out.append(grouped_xyz)
        if self.ret_unique_cnt:
            out.append(unique_cnt)
        return out[0] if len(out) == 1 else tuple(out)


class GroupAll(nn.Module):
    """Group every point into one region: (B, 3+C, 1, N).  (The reference's version drops
    ret_grouped_xyz in __init__, pointnet2_utils.py:387-390; here it is honoured.)"""

    def __init__(self, use_xyz=True, ret_grouped_xyz=False):
        super().__init__()
        self.use_xyz = use_xyz
        self.ret_grouped_xyz = ret_grouped_xyz

    def forward(self, xyz, new_xyz, features=None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            new_features = grouped_xyz
        else:
            grouped = features.unsqueeze(2)
            new_features = torch.cat([grouped_xyz, grouped], dim=1) if self.use_xyz else grouped
        return (new_features, grouped_xyz) if self.ret_grouped_xyz else new_features


class RandomDropout(nn.Module):
    """Present for API completeness.  The reference's forward calls a function that does
    not exist (pt_utils.feature_dropout_no_scaling, pointnet2_utils.py:40-48) and is never
    used by BridgeQA; this one implements the evident intent (unscaled feature dropout
    with a random rate in [0, p))."""

    def __init__(self, p=0.5, inplace=False):
        super().__init__()
        self.p = p
        self.inplace = inplace

    def forward(self, X):
        if not self.training:
            return X
        theta = float(torch.empty(1).uniform_(0, self.p))
        keep = (torch.rand(X.shape[:2], device=X.device) >= theta).to(X.dtype)
        keep = keep.reshape(keep.shape + (1,) * (X.dim() - 2))
        return X.mul_(keep) if self.inplace else X * keep
