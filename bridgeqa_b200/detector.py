"""VoteNet detector front-end built on the B200 point-cloud operators: host mirror of the
three reference model files that sit directly on the hot path,

    Pointnet2Backbone   /root/reference/models/backbone_module.py:11-131
    VotingModule        /root/reference/models/voting_module.py:11-60
    ProposalModule      /root/reference/models/proposal_module.py:20-151

with identical constructor arguments, data_dict keys and state_dict keys, so a BridgeQA /
VoteNet checkpoint's `detection_backbone.*`, `voting_net.*` and `proposal_net.*` entries
load with strict=True.  (The reference files themselves also run unchanged on top of
`bridgeqa_b200.compat`; see INTEGRATION.md.  These classes exist because the reference
tree does not travel to the benchmark machine.)

`detect()` strings them together the way ScanQA.forward does
(/root/reference/models/qa_module.py:438-459).
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import fused, pointnet2_utils
from .pointnet2_modules import PointnetFPModule, PointnetSAModuleVotes


class Pointnet2Backbone(nn.Module):
    """PointNet++ single-scale-grouping backbone: 4 set-abstraction + 2 feature-propagation
    layers.  Input data_dict["point_clouds"] is (B, N, 3 + input_feature_dim)."""

    # (npoint, radius, nsample, in_width_units, hidden_units, out_units) in units of `width`
    _SA = (
        (2048, 0.2, 64, None, 64, 128),
        (1024, 0.4, 32, 128, 128, 256),
        (512, 0.8, 16, 256, 128, 256),
        (256, 1.2, 16, 256, 128, 256),
    )

    def __init__(self, input_feature_dim=0, width=1, depth=2, seed_feat_dim=256):
        super().__init__()
        self.input_feature_dim = input_feature_dim
        for i, (npoint, radius, nsample, cin, hid, cout) in enumerate(self._SA, start=1):
            first = input_feature_dim if cin is None else cin * width
            spec = [first] + [hid * width] * depth + [cout * width]
            setattr(self, "sa%d" % i, PointnetSAModuleVotes(
                npoint=npoint, radius=radius, nsample=nsample, mlp=spec, use_xyz=True,
                normalize_xyz=True))
        c = 256 * width
        import os
        self.prefix_check = os.environ.get("BQA_FPS_PREFIX_CHECK", "1") != "0"
        # sampling over the cell-sorted points with warp-level pruning (fps_sorted.cu): bit-identical;
        # at 16 x 40000 the kernel is 7 % faster (1.60 -> 1.48 ms) but has to wait for the grid
        # build (0.09 ms) that otherwise hides under the sampling: +1.6 % on the whole forward
        self.fps_grid = os.environ.get("BQA_FPS_GRID", "1") != "0"
        self.fp1 = PointnetFPModule(mlp=[c + c, c, c])
        self.fp2 = PointnetFPModule(mlp=[c + c, c, seed_feat_dim])
        self.fp2.emit_point_major = False      # nothing downstream gathers from fp2's output rows

    @staticmethod
    def _break_up_pc(pc):
        if getattr(pc, "is_staged", False):
            # loader-facing staged input (staging.StagedCloud): fp32 coordinates + the 16-bit point-major
            # features the fused SA1 kernel gathers from; no split / transpose / conversion on the device
            if pc.precision != fused.precision():
                raise RuntimeError("staged cloud holds %s features but the operand precision is %s"
                                   % (pc.precision, fused.precision()))
            return pc.xyz, pc.features_view()
        xyz = pc[..., :3].contiguous()
        # (B,C,N) VIEW of the cloud (the reference materialises it, backbone_module.py:74-78): every
        # consumer on the hot path gathers from the point-major twin below instead, so the
        # transposed copy (338 MB at C=132) is only made by a consumer that really needs it
        features = pc[..., 3:].transpose(1, 2) if pc.size(-1) > 3 else None
        if features is not None:
            # the input cloud already is the point-major layout the fused SA kernel gathers from;
            # when the channel count allows 16-byte loads (C % 4 == 0, e.g. the 132-d multiview
            # config) one packed copy makes the rows aligned (the view starts 12 bytes in)
            c = pc.size(-1) - 3
            features._bqa_pm = pc[..., 3:].contiguous() if (c % 4 == 0 and c >= 16 and not fused.sa_v2_enabled()) \
                else pc[..., 3:]
            features._bqa_cloud = pc        # fused.point_major_16 converts the feature columns row-wise
        return xyz, features

    def _sample_lower_levels(self, xyz1):
        """Inference only.  The FPS stages depend on coordinates alone, so levels 2-4 run on a
        side stream underneath SA1's ball query + MLP instead of in front of SA2/3/4.  Their
        input is a cloud already in sampling order, so one parallel check (fused.fps_prefix_check)
        usually proves that all three samplings are identity prefixes and the serial chains
        (~0.7 ms) are skipped; scenes it cannot prove go through the chain as before.
        Returns [(inds, new_xyz, event)] for SA2..SA4."""
        main = torch.cuda.current_stream(xyz1.device)
        side = fused.side_stream(xyz1.device, "fps_levels")
        out = []
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(side):
            side.wait_event(ready)
            cur = xyz1
            covered = 0 < self.sa2.npoint <= min(xyz1.size(1), 8192) and self.prefix_check
            flags = fused.fps_prefix_check(xyz1, self.sa2.npoint) if covered else None
            for sa in (self.sa2, self.sa3, self.sa4):
                if flags is not None and sa.npoint <= min(self.sa2.npoint, cur.size(1)):
                    # cur is a prefix of xyz1 for flagged scenes (previous level = identity prefix)
                    inds, new_xyz = fused.furthest_point_sample_cond(cur, sa.npoint, flags)
                else:
                    flags = None
                    inds, new_xyz = pointnet2_utils.furthest_point_sample_with_xyz(cur, sa.npoint)
                done = torch.cuda.Event()
                done.record(side)
                for t in (inds, new_xyz):
                    t.record_stream(main)
                out.append((inds, new_xyz, done))
                cur = new_xyz
        return out

    def prefetch_sampling(self, point_clouds):
        """SA1's sampling of a FUTURE batch (cell grid + furthest point sampling: coordinates
        only, no weights involved) issued now on a side stream, so that it runs underneath
        whatever the current stream does next -- in training, the current step's backward.  The
        forward of exactly this tensor (same storage, same version) then picks the result up
        instead of sampling in front of SA1.  Uses the throughput variant of the sampling
        kernel (3 SMs per 40k-point scene); indices are the same bits either way."""
        pc = point_clouds
        sa = self.sa1
        if not (torch.is_tensor(pc) and pc.is_cuda and fused.enabled() and self.fps_grid
                and fused.fps_grid_supported(pc.size(1), sa.npoint)):
            return False
        dev = pc.device
        main = torch.cuda.current_stream(dev)
        side = fused.side_stream(dev, "prefetch")
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(side), torch.no_grad(), fused.lean_sampling(True):
            side.wait_event(ready)
            xyz = pc[..., :3].contiguous()
            grid = fused.prebuild_ball_query_grid(xyz, sa.radius, inline=True)
            inds, new_xyz = fused.furthest_point_sample_grid(xyz, sa.npoint, grid)
            done = torch.cuda.Event()
            done.record(side)
        pc.record_stream(side)
        # the tensor itself is kept (not its address): a freed block can come back from the caching
        # allocator at the same address with version 0 and would otherwise match a stale entry
        self._prefetched = {"pc": pc, "version": pc._version, "xyz": xyz, "grid": grid,
                            "inds": inds, "new_xyz": new_xyz, "done": done}
        return True

    def _take_prefetched(self, pc):
        pre, self._prefetched = getattr(self, "_prefetched", None), None
        if pre is None or pre["pc"] is not pc or pre["version"] != pc._version:
            return None
        main = torch.cuda.current_stream(pc.device)
        main.wait_event(pre["done"])
        for t in (pre["xyz"], pre["grid"][0], pre["inds"], pre["new_xyz"]):
            t.record_stream(main)
        return pre

    def enable_cuda_graph(self, on=True, bind_inputs=False):
        """Inference forwards of a fixed input shape replay a captured CUDA graph (graphs.py);
        the returned tensors are then static buffers the next call overwrites."""
        from . import graphs
        if on and (getattr(self, "_graph_runner", None) is None
                   or self._graph_runner.bind_inputs != bind_inputs):
            self._graph_runner = graphs.GraphedForward(self, self._forward_impl, bind_inputs)
        self._graphed = self._graph_runner if on else None
        return self

    def in_flight(self, depth=2, lean_sampling=None):
        """Queue that keeps `depth` inference forwards in flight on their own streams, so that
        the next batch's sampling chain runs under this batch's SA/FP kernels (graphs.InFlight)."""
        from . import graphs
        return graphs.InFlight(self, depth, lean_sampling)

    def forward(self, data_dict):
        g = getattr(self, "_graphed", None)
        if g is not None and g.applicable(data_dict):
            return g(data_dict)
        return self._forward_impl(data_dict)

    def train(self, mode=True):
        self._prefetched = None
        return super().train(mode)

    def _forward_impl(self, data_dict):
        xyz, features = self._break_up_pc(data_dict["point_clouds"])

        overlap = (fused.enabled() and xyz.is_cuda and not self.training
                   and not torch.is_grad_enabled())
        if overlap:
            self._prefetched = None        # a prefetch is only ever consumed by the training branch
            main = torch.cuda.current_stream(xyz.device)
            # SA1's cell grid serves both the sampling (pruned FPS over the cell-sorted points) and
            # the ball query; without the sorted sampling it is built on a side stream underneath
            # FPS1.  Levels 2-4 are sampled on a second side stream as soon as SA1's centres exist.
            grid1 = None
            if self.sa1._can_fuse(xyz, features):
                sorted_fps = self.fps_grid and fused.fps_grid_supported(xyz.size(1), self.sa1.npoint)
                grid1 = fused.prebuild_ball_query_grid(xyz, self.sa1.radius, inline=sorted_fps)
            if grid1 is not None and grid1[1] is None:
                inds1, xyz1 = fused.furthest_point_sample_grid(xyz, self.sa1.npoint, grid1)
            else:
                inds1, xyz1 = pointnet2_utils.furthest_point_sample_with_xyz(xyz, self.sa1.npoint)
            levels = self._sample_lower_levels(xyz1)
            # grids of levels 2-4: each needs the coordinates one sampling level up
            grids = [fused.prebuild_ball_query_grid(xyz1, self.sa2.radius),
                     fused.prebuild_ball_query_grid(levels[0][1], self.sa3.radius, after=levels[0][2]),
                     fused.prebuild_ball_query_grid(levels[1][1], self.sa4.radius, after=levels[1][2])]
            xyz1, feats1, inds1 = self.sa1(xyz, features, inds1, new_xyz=xyz1, grid=grid1)
            outs = [(xyz1, feats1, inds1)]
            xyz, features = xyz1, feats1
            for sa, (inds, new_xyz, event), grid in zip((self.sa2, self.sa3, self.sa4), levels, grids):
                main.wait_event(event)
                xyz, features, inds = sa(xyz, features, inds, new_xyz=new_xyz,
                                         grid=grid if sa._can_fuse(xyz, features) else None)
                outs.append((xyz, features, inds))
        else:
            if getattr(data_dict["point_clouds"], "is_staged", False):
                raise RuntimeError("a staged 16-bit cloud is only accepted by the fused inference path "
                                   "(eval mode, no autograd, fused kernels enabled)")
            # training / un-fused path: same layer sequence as the reference
            # (models/backbone_module.py:97-121); on the GPU the re-sampling of sampled clouds at
            # levels 2-4 still goes through the parallel identity-prefix proof
            outs = []
            flags = None
            pre = self._take_prefetched(data_dict["point_clouds"]) if xyz.is_cuda and not xyz.requires_grad else None
            for k, sa in enumerate((self.sa1, self.sa2, self.sa3, self.sa4)):
                inds = new_xyz = grid = None
                plain = xyz.is_cuda and not xyz.requires_grad
                if k == 0 and pre is not None:
                    # sampled ahead of time under the previous step's backward (prefetch_sampling)
                    grid, inds, new_xyz = pre["grid"], pre["inds"], pre["new_xyz"]
                elif k == 0 and plain and self.fps_grid and fused.enabled() \
                        and fused.fps_grid_supported(xyz.size(1), sa.npoint):
                    # SA1: one cell grid serves the pruned sampling and the ball query
                    grid = fused.prebuild_ball_query_grid(xyz, sa.radius, inline=True)
                    inds, new_xyz = fused.furthest_point_sample_grid(xyz, sa.npoint, grid)
                if k >= 1 and plain and self.prefix_check:
                    if k == 1 and 0 < sa.npoint <= min(xyz.size(1), 8192):
                        flags = fused.fps_prefix_check(xyz, sa.npoint)
                    if flags is not None and sa.npoint <= min(self.sa2.npoint, xyz.size(1)):
                        inds, new_xyz = fused.furthest_point_sample_cond(xyz, sa.npoint, flags)
                    else:
                        flags = None
                xyz, features, inds = sa(xyz, features, inds, new_xyz=new_xyz, grid=grid)
                outs.append((xyz, features, inds))
        data_dict["sa1_xyz"], data_dict["sa1_features"], data_dict["sa1_inds"] = outs[0]
        data_dict["sa2_xyz"], data_dict["sa2_features"], data_dict["sa2_inds"] = outs[1]
        data_dict["sa3_xyz"], data_dict["sa3_features"] = outs[2][0], outs[2][1]
        data_dict["sa4_xyz"], data_dict["sa4_features"] = outs[3][0], outs[3][1]

        features = self.fp1(data_dict["sa3_xyz"], data_dict["sa4_xyz"],
                            data_dict["sa3_features"], data_dict["sa4_features"])
        features = self.fp2(data_dict["sa2_xyz"], data_dict["sa3_xyz"],
                            data_dict["sa2_features"], features)
        data_dict["fp2_features"] = features
        data_dict["fp2_xyz"] = data_dict["sa2_xyz"]
        num_seed = data_dict["fp2_xyz"].shape[1]
        # seeds index the ORIGINAL cloud: the first num_seed SA1 samples (backbone_module.py:130)
        data_dict["fp2_inds"] = data_dict["sa1_inds"][:, 0:num_seed]
        return data_dict


class VotingModule(nn.Module):
    """Seeds -> votes: 3 pointwise convs; output = seed + predicted (xyz offset, feature residual)."""

    def __init__(self, vote_factor, seed_feature_dim):
        super().__init__()
        self.vote_factor = vote_factor
        self.in_dim = seed_feature_dim
        self.out_dim = self.in_dim
        self.conv1 = nn.Conv1d(self.in_dim, self.in_dim, 1)
        self.conv2 = nn.Conv1d(self.in_dim, self.in_dim, 1)
        self.conv3 = nn.Conv1d(self.in_dim, (3 + self.out_dim) * self.vote_factor, 1)
        self.bn1 = nn.BatchNorm1d(self.in_dim)
        self.bn2 = nn.BatchNorm1d(self.in_dim)

    def forward(self, seed_xyz, seed_features):
        b, num_seed = seed_xyz.shape[0], seed_xyz.shape[1]
        vf, d = self.vote_factor, self.out_dim
        net = F.relu(self.bn1(self.conv1(seed_features)))
        net = F.relu(self.bn2(self.conv2(net)))
        net = self.conv3(net)                                   # (B, (3+d)*vf, num_seed)
        net = net.transpose(2, 1).reshape(b, num_seed, vf, 3 + d)
        vote_xyz = (seed_xyz.unsqueeze(2) + net[..., 0:3]).reshape(b, num_seed * vf, 3)
        vote_features = seed_features.transpose(2, 1).unsqueeze(2) + net[..., 3:]
        vote_features = vote_features.reshape(b, num_seed * vf, d).transpose(2, 1).contiguous()
        return vote_xyz.contiguous(), vote_features


class DatasetConfig(object):
    """Stand-in for `data.scannet.model_util_scannet.ScannetDatasetConfig`, which the
    reference imports (models/proposal_module.py:11) from an un-vendored data tree (dangling
    symlink in /root/reference).  ScanNet boxes are axis-aligned: one heading bin whose angle
    is always 0 (heading_mode="zero", the VoteNet/ScanRefer ScanNet config: class2angle_batch
    returns zeros); heading_mode="bins" decodes `class * 2pi/num_heading_bin + residual`, wrapped
    to (-pi, pi] (the SUN RGB-D style config of the same code base).  18 classes = 18 size
    clusters.  mean_size_arr is data (a .npz in the original); here it defaults to ones and can
    be replaced."""

    def __init__(self, num_class=18, num_heading_bin=1, num_size_cluster=18, mean_size_arr=None,
                 heading_mode="zero"):
        if heading_mode not in ("zero", "bins"):
            raise ValueError("heading_mode must be 'zero' or 'bins'")
        self.num_class = num_class
        self.num_heading_bin = num_heading_bin
        self.num_size_cluster = num_size_cluster
        self.heading_mode = heading_mode
        self.mean_size_arr = (np.ones((num_size_cluster, 3), dtype=np.float32)
                              if mean_size_arr is None else np.asarray(mean_size_arr, dtype=np.float32))

    def class2angle(self, pred_cls, residual, to_label_format=True):
        if self.heading_mode == "zero":
            return 0
        angle = pred_cls * (2 * np.pi / float(self.num_heading_bin)) + residual
        if to_label_format and angle > np.pi:
            angle = angle - 2 * np.pi
        return angle

    def class2size(self, pred_cls, residual):
        return self.mean_size_arr[pred_cls] + residual


def decode_heading(heading_scores, heading_residuals, heading_mode):
    """Heading angle (B,K) as `param2obb_batch` hands it to get_3d_box_batch: the NEGATED
    class2angle_batch(argmax class, its residual); all on the device."""
    cls = torch.argmax(heading_scores, -1)
    if heading_mode == "zero":
        return torch.zeros(cls.shape, dtype=heading_residuals.dtype, device=cls.device)
    nh = heading_scores.size(-1)
    res = torch.gather(heading_residuals, 2, cls.unsqueeze(-1)).squeeze(-1)
    angle = cls.to(res.dtype) * (2 * math.pi / float(nh)) + res
    angle = torch.where(angle > math.pi, angle - 2 * math.pi, angle)
    return -angle


def box_corners(box_size, heading_angle, center):
    """Torch / on-device version of utils/box_util.py:302-325 (get_3d_box_batch):
    box_size (...,3), heading (...), center (...,3) -> (...,8,3)."""
    l, w, h = box_size[..., 0:1] / 2, box_size[..., 1:2] / 2, box_size[..., 2:3] / 2
    cx = torch.cat((l, l, -l, -l, l, l, -l, -l), -1)
    cy = torch.cat((w, -w, -w, w, w, -w, -w, w), -1)
    cz = torch.cat((h, h, h, h, -h, -h, -h, -h), -1)
    c, s = torch.cos(heading_angle).unsqueeze(-1), torch.sin(heading_angle).unsqueeze(-1)
    # corners @ R^T with R = roty(t) = [[c,0,s],[0,1,0],[-s,0,c]]
    x = cx * c + cz * s
    z = -cx * s + cz * c
    return torch.stack((x, cy, z), -1) + center.unsqueeze(-2)


class ProposalModule(nn.Module):
    """Vote aggregation (an SA layer over the votes) + proposal head + score decoding."""

    def __init__(self, num_class, num_heading_bin, num_size_cluster, mean_size_arr, num_proposal,
                 sampling, seed_feat_dim=256, proposal_size=128, radius=0.3, nsample=16,
                 heading_mode="zero"):
        super().__init__()
        if heading_mode not in ("zero", "bins"):
            raise ValueError("heading_mode must be 'zero' (ScanNet) or 'bins'")
        self.heading_mode = heading_mode
        self.num_class = num_class
        self.num_heading_bin = num_heading_bin
        self.num_size_cluster = num_size_cluster
        self.mean_size_arr = mean_size_arr
        self.num_proposal = num_proposal
        self.sampling = sampling
        self.seed_feat_dim = seed_feat_dim
        self.votenet_hidden_size = proposal_size
        self.vote_aggregation = PointnetSAModuleVotes(
            npoint=num_proposal, radius=radius, nsample=nsample,
            mlp=[seed_feat_dim, proposal_size, proposal_size, proposal_size],
            use_xyz=True, normalize_xyz=True)
        out_ch = 2 + 3 + num_heading_bin * 2 + num_size_cluster * 4 + num_class
        self.proposal = nn.Sequential(
            nn.Conv1d(proposal_size, proposal_size, 1, bias=False),
            nn.BatchNorm1d(proposal_size),
            nn.ReLU(),
            nn.Conv1d(proposal_size, proposal_size, 1, bias=False),
            nn.BatchNorm1d(proposal_size),
            nn.ReLU(),
            nn.Conv1d(proposal_size, out_ch, 1))
        self.register_buffer("_mean_size", torch.as_tensor(np.asarray(mean_size_arr, dtype=np.float32)),
                             persistent=False)

    def forward(self, xyz, features, data_dict):
        xyz, features, fps_inds = self.vote_aggregation(xyz, features)
        data_dict["aggregated_vote_xyz"] = xyz
        data_dict["aggregated_vote_features"] = features.permute(0, 2, 1).contiguous()
        data_dict["aggregated_vote_inds"] = fps_inds
        net = self.proposal(features)                         # (B, 97, num_proposal)
        return self.decode_scores(net, data_dict)

    def decode_scores(self, net, data_dict):
        nh, ns = self.num_heading_bin, self.num_size_cluster
        t = net.transpose(2, 1).contiguous()                  # (B, K, 97)
        b, k = t.shape[0], t.shape[1]
        o = 5 + 2 * nh
        data_dict["objectness_scores"] = t[:, :, 0:2]
        data_dict["center"] = data_dict["aggregated_vote_xyz"] + t[:, :, 2:5]
        data_dict["heading_scores"] = t[:, :, 5:5 + nh]
        data_dict["heading_residuals_normalized"] = t[:, :, 5 + nh:o]
        data_dict["heading_residuals"] = data_dict["heading_residuals_normalized"] * (math.pi / nh)
        data_dict["size_scores"] = t[:, :, o:o + ns]
        srn = t[:, :, o + ns:o + 4 * ns].view(b, k, ns, 3)
        data_dict["size_residuals_normalized"] = srn
        data_dict["size_residuals"] = srn * self._mean_size.to(srn.device)[None, None]
        data_dict["sem_cls_scores"] = t[:, :, o + 4 * ns:]
        data_dict["bbox_corner"] = self.decode_pred_box(data_dict)
        data_dict["bbox_feature"] = data_dict["aggregated_vote_features"]
        data_dict["bbox_mask"] = data_dict["objectness_scores"].argmax(-1)
        data_dict["bbox_sems"] = data_dict["sem_cls_scores"].argmax(-1)
        return data_dict

    def decode_pred_box(self, data_dict):
        """Box corners (B, K, 8, 3) on the device.  The reference does this through
        .cpu().numpy() + param2obb_batch + get_3d_box_batch + .cuda()
        (proposal_module.py:87-108; utils/box_util.py:302-325), a host sync inside the forward.
        Same decode here in torch ops: size = mean_size[argmax size class] + its residual, heading
        = -class2angle(argmax heading class, its residual) (0 for ScanNet's single bin), corners =
        the 8 sign patterns rotated about y and shifted to the centre."""
        size_cls = torch.argmax(data_dict["size_scores"], -1)                       # (B,K)
        res = torch.gather(data_dict["size_residuals"], 2,
                           size_cls[..., None, None].expand(-1, -1, 1, 3)).squeeze(2)
        box_size = self._mean_size.to(res.device)[size_cls] + res
        heading = decode_heading(data_dict["heading_scores"], data_dict["heading_residuals"],
                                 self.heading_mode)
        return box_corners(box_size, heading, data_dict["center"])


class VoteNetDetector(nn.Module):
    """backbone -> voting -> L2-normalise vote features -> proposal, i.e. the detection
    branch of ScanQA.forward (models/qa_module.py:438-459) with the reference's default
    hyper-parameters (scripts/train.py:71-86)."""

    def __init__(self, input_feature_dim, config=None, num_proposal=256, vote_factor=1,
                 sampling="vote_fps", seed_feat_dim=256, proposal_size=128, pointnet_width=1,
                 pointnet_depth=2, vote_radius=0.3, vote_nsample=16):
        super().__init__()
        cfg = config or DatasetConfig()
        self.detection_backbone = Pointnet2Backbone(
            input_feature_dim=input_feature_dim, width=pointnet_width, depth=pointnet_depth,
            seed_feat_dim=seed_feat_dim)
        self.voting_net = VotingModule(vote_factor, seed_feat_dim)
        self.proposal_net = ProposalModule(
            cfg.num_class, cfg.num_heading_bin, cfg.num_size_cluster, cfg.mean_size_arr,
            num_proposal, sampling, seed_feat_dim=seed_feat_dim, proposal_size=proposal_size,
            radius=vote_radius, nsample=vote_nsample, heading_mode=getattr(cfg, "heading_mode", "zero"))

    def enable_cuda_graph(self, on=True, bind_inputs=False):
        """See Pointnet2Backbone.enable_cuda_graph; here the whole detector forward is one graph."""
        from . import graphs
        if on and (getattr(self, "_graph_runner", None) is None
                   or self._graph_runner.bind_inputs != bind_inputs):
            self._graph_runner = graphs.GraphedForward(self, self._forward_impl, bind_inputs)
        self._graphed = self._graph_runner if on else None
        return self

    def in_flight(self, depth=2, lean_sampling=None):
        """Queue that keeps `depth` inference forwards in flight on their own streams, so that
        the next batch's sampling chain runs under this batch's SA/FP kernels (graphs.InFlight)."""
        from . import graphs
        return graphs.InFlight(self, depth, lean_sampling)

    def forward(self, data_dict):
        g = getattr(self, "_graphed", None)
        if g is not None and g.applicable(data_dict):
            return g(data_dict)
        return self._forward_impl(data_dict)

    def _forward_impl(self, data_dict):
        bb = self.detection_backbone
        # (a backbone with its own graph must not replay it inside this module's capture)
        data_dict = bb._forward_impl(data_dict) if torch.cuda.is_current_stream_capturing() else bb(data_dict)
        xyz, features = data_dict["fp2_xyz"], data_dict["fp2_features"]
        data_dict["seed_inds"] = data_dict["fp2_inds"]
        data_dict["seed_xyz"] = xyz
        data_dict["seed_features"] = features
        xyz, features = self.voting_net(xyz, features)
        features = features.div(torch.norm(features, p=2, dim=1).unsqueeze(1))
        data_dict["vote_xyz"] = xyz
        data_dict["vote_features"] = features
        return self.proposal_net(xyz, features, data_dict)
