"""CPU restatement of the reference's SA / FP / backbone / voting / proposal forward
(eval mode), composed from the C oracle ops (oracle/cpu_ops.py) and torch-CPU fp32
conv / batch-norm / relu / max -- the reference's own dense path, run on the host.

TEST INFRASTRUCTURE ONLY (parity checker + the reported CPU baseline).

Follows, line by line:
  PointnetSAModuleVotes.forward   /root/reference/lib/pointnet2/pointnet2_modules.py:210-277
  QueryAndGroup.forward           /root/reference/lib/pointnet2/pointnet2_utils.py:317-376
  PointnetFPModule.forward        /root/reference/lib/pointnet2/pointnet2_modules.py:376-421
  Pointnet2Backbone.forward       /root/reference/models/backbone_module.py:80-131
  VotingModule.forward            /root/reference/models/voting_module.py:33-60
  ScanQA.forward (detector part)  /root/reference/models/qa_module.py:438-459
  ProposalModule.forward          /root/reference/models/proposal_module.py:58-85

Weights come from a state_dict with the reference's key names, so the same dict drives
the reference modules, this file and the product.  Pinned against the real reference
Python layer by tests/golden/make_golden_cpu.py (run in the build container, where
/root/reference is importable).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import cpu_ops as ops

BN_EPS = 1e-5


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def shared_mlp(x, sd, prefix, nlayers):
    """x (B,C,H,W) torch; layers `{prefix}layer{i}.conv.weight` + `.bn.bn.*`; eval-mode BN."""
    for i in range(nlayers):
        p = "%slayer%d." % (prefix, i)
        x = F.conv2d(x, sd[p + "conv.weight"], sd.get(p + "conv.bias"))
        if p + "bn.bn.weight" in sd:
            x = F.batch_norm(x, sd[p + "bn.bn.running_mean"], sd[p + "bn.bn.running_var"],
                             sd[p + "bn.bn.weight"], sd[p + "bn.bn.bias"], False, 0.0, BN_EPS)
        x = F.relu(x)
    return x


def _count_layers(sd, prefix):
    n = 0
    while "%slayer%d.conv.weight" % (prefix, n) in sd:
        n += 1
    return n


def sa_layer(xyz, features, sd, prefix, npoint, radius, nsample, normalize_xyz=True, inds=None):
    """xyz (B,N,3) np, features (B,C,N) np | None -> new_xyz (B,npoint,3), new_features
    (B,Cout,npoint), inds (B,npoint) int32, plus the ball-query idx for index parity."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    if inds is None:
        inds = ops.furthest_point_sampling(xyz, npoint)                   # :235
    xyz_flipped = np.ascontiguousarray(xyz.transpose(0, 2, 1))            # :233
    new_xyz = np.ascontiguousarray(ops.gather_points(xyz_flipped, inds).transpose(0, 2, 1))  # :238-240
    idx = ops.ball_query(new_xyz, xyz, radius, nsample)                   # utils:334
    grouped_xyz = ops.group_points(xyz_flipped, idx)                      # utils:349
    grouped_xyz = grouped_xyz - new_xyz.transpose(0, 2, 1)[..., None]     # utils:350
    if normalize_xyz:
        grouped_xyz = grouped_xyz / np.float32(radius)                    # utils:351-352
    if features is not None:
        grouped = ops.group_points(np.ascontiguousarray(features, dtype=np.float32), idx)
        new_features = np.concatenate([grouped_xyz, grouped], axis=1)     # utils:357-359
    else:
        new_features = grouped_xyz
    y = shared_mlp(_t(new_features.astype(np.float32)), sd, prefix + "mlp_module.",
                   _count_layers(sd, prefix + "mlp_module."))             # :251
    y = F.max_pool2d(y, kernel_size=[1, y.size(3)]).squeeze(-1)           # :259-262
    return new_xyz, y.numpy(), inds, idx


def fp_layer(unknown, known, unknow_feats, known_feats, sd, prefix):
    """-> (B, Cout, n) np, plus three_nn idx for index parity."""
    dist2, idx = ops.three_nn(unknown, known)                             # :399
    dist = np.sqrt(dist2)                                                 # utils:142
    dist_recip = (np.float32(1.0) / (dist + np.float32(1e-8))).astype(np.float32)   # :400
    norm = np.sum(dist_recip, axis=2, keepdims=True, dtype=np.float32)    # :401
    weight = (dist_recip / norm).astype(np.float32)                       # :402
    interpolated = ops.three_interpolate(known_feats, idx, weight)        # :404
    if unknow_feats is not None:
        new_features = np.concatenate([interpolated, unknow_feats], axis=1)   # :413
    else:
        new_features = interpolated
    y = shared_mlp(_t(new_features.astype(np.float32))[..., None], sd, prefix + "mlp.",
                   _count_layers(sd, prefix + "mlp."))
    return y.squeeze(-1).numpy(), idx


SA_CFG = ((2048, 0.2, 64), (1024, 0.4, 32), (512, 0.8, 16), (256, 1.2, 16))


def backbone(point_clouds, sd, prefix="", sa_cfg=SA_CFG):
    """point_clouds (B,N,3+C) np -> dict with the reference's data_dict keys (numpy)."""
    pc = np.asarray(point_clouds, dtype=np.float32)
    xyz = np.ascontiguousarray(pc[..., :3])
    features = np.ascontiguousarray(pc[..., 3:].transpose(0, 2, 1)) if pc.shape[-1] > 3 else None
    out = {}
    for i, (npoint, radius, nsample) in enumerate(sa_cfg, start=1):
        xyz, features, inds, bq = sa_layer(xyz, features, sd, "%ssa%d." % (prefix, i), npoint,
                                           radius, nsample)
        out["sa%d_xyz" % i], out["sa%d_features" % i] = xyz, features
        out["sa%d_inds" % i], out["sa%d_ball_idx" % i] = inds, bq
    f, out["fp1_nn_idx"] = fp_layer(out["sa3_xyz"], out["sa4_xyz"], out["sa3_features"],
                                    out["sa4_features"], sd, prefix + "fp1.")
    f, out["fp2_nn_idx"] = fp_layer(out["sa2_xyz"], out["sa3_xyz"], out["sa2_features"], f, sd,
                                    prefix + "fp2.")
    out["fp2_features"] = f
    out["fp2_xyz"] = out["sa2_xyz"]
    out["fp2_inds"] = out["sa1_inds"][:, :out["fp2_xyz"].shape[1]]
    return out


def voting(seed_xyz, seed_features, sd, prefix, vote_factor=1):
    x = _t(seed_features)

    def bn(y, name):
        return F.batch_norm(y, sd[prefix + name + ".running_mean"], sd[prefix + name + ".running_var"],
                            sd[prefix + name + ".weight"], sd[prefix + name + ".bias"], False, 0.0, BN_EPS)

    net = F.relu(bn(F.conv1d(x, sd[prefix + "conv1.weight"], sd[prefix + "conv1.bias"]), "bn1"))
    net = F.relu(bn(F.conv1d(net, sd[prefix + "conv2.weight"], sd[prefix + "conv2.bias"]), "bn2"))
    net = F.conv1d(net, sd[prefix + "conv3.weight"], sd[prefix + "conv3.bias"])
    b, _, num_seed = x.shape
    d = x.shape[1]
    net = net.transpose(2, 1).reshape(b, num_seed, vote_factor, 3 + d)
    vote_xyz = (_t(seed_xyz).unsqueeze(2) + net[..., 0:3]).reshape(b, num_seed * vote_factor, 3)
    vote_features = x.transpose(2, 1).unsqueeze(2) + net[..., 3:]
    vote_features = vote_features.reshape(b, num_seed * vote_factor, d).transpose(2, 1).contiguous()
    return vote_xyz.contiguous().numpy(), vote_features.numpy()


def detector(point_clouds, sd, num_proposal=256, vote_radius=0.3, vote_nsample=16):
    """Detector forward up to the proposal head's raw output (before box decoding)."""
    out = backbone(point_clouds, sd, "detection_backbone.")
    vote_xyz, vote_features = voting(out["fp2_xyz"], out["fp2_features"], sd, "voting_net.")
    vf = _t(vote_features)
    vf = vf.div(torch.norm(vf, p=2, dim=1).unsqueeze(1)).numpy()
    out["vote_xyz"], out["vote_features"] = vote_xyz, vf
    xyz, feats, inds, bq = sa_layer(vote_xyz, vf, sd, "proposal_net.vote_aggregation.", num_proposal,
                                    vote_radius, vote_nsample)
    out["aggregated_vote_xyz"] = xyz
    out["aggregated_vote_features"] = np.ascontiguousarray(feats.transpose(0, 2, 1))
    out["aggregated_vote_inds"] = inds
    p = "proposal_net.proposal."
    x = _t(feats)

    def bn(y, i):
        return F.batch_norm(y, sd["%s%d.running_mean" % (p, i)], sd["%s%d.running_var" % (p, i)],
                            sd["%s%d.weight" % (p, i)], sd["%s%d.bias" % (p, i)], False, 0.0, BN_EPS)

    x = F.relu(bn(F.conv1d(x, sd[p + "0.weight"]), 1))
    x = F.relu(bn(F.conv1d(x, sd[p + "3.weight"]), 4))
    x = F.conv1d(x, sd[p + "6.weight"], sd[p + "6.bias"])
    out["proposal_scores"] = x.numpy()
    return out


def proposal_head(aggregated_features, sd, prefix="proposal_net.proposal."):
    """(B,128,K) np -> raw head output (B, 2+3+2nh+4ns+nc, K)  (models/proposal_module.py:43-56, 81)."""
    p = prefix
    x = _t(aggregated_features)

    def bn(y, i):
        return F.batch_norm(y, sd["%s%d.running_mean" % (p, i)], sd["%s%d.running_var" % (p, i)],
                            sd["%s%d.weight" % (p, i)], sd["%s%d.bias" % (p, i)], False, 0.0, BN_EPS)

    x = F.relu(bn(F.conv1d(x, sd[p + "0.weight"]), 1))
    x = F.relu(bn(F.conv1d(x, sd[p + "3.weight"]), 4))
    return F.conv1d(x, sd[p + "6.weight"], sd[p + "6.bias"]).numpy()


def decode_scores(net, base_xyz, mean_size_arr, num_heading_bin=1, num_size_cluster=18, heading_mode="zero"):
    """models/proposal_module.py:110-151 (decode_scores) + :87-108 (decode_pred_box) in NumPy:
    slices of the transposed head output in float32 as torch does, then the box through
    param2obb_batch (un-vendored data/scannet/model_util_scannet.py; its published form: size =
    mean_size[class] + residual, heading = -class2angle_batch, 0 for ScanNet) and
    utils/box_util.py:302-325 get_3d_box_batch in float64 (the reference's NumPy dtype)."""
    nh, ns = num_heading_bin, num_size_cluster
    t = np.ascontiguousarray(np.transpose(net, (0, 2, 1))).astype(np.float32)      # :115
    b, k = t.shape[:2]
    o = 5 + 2 * nh
    out = {"objectness_scores": t[:, :, 0:2],                                       # :119
           "center": (base_xyz.astype(np.float32) + t[:, :, 2:5]).astype(np.float32),   # :121-122
           "heading_scores": t[:, :, 5:5 + nh],
           "heading_residuals": (t[:, :, 5 + nh:o] * np.float32(np.pi / nh)).astype(np.float32),   # :136
           "size_scores": t[:, :, o:o + ns],
           "sem_cls_scores": t[:, :, o + 4 * ns:]}
    srn = t[:, :, o + ns:o + 4 * ns].reshape(b, k, ns, 3)
    mean_size = np.asarray(mean_size_arr, dtype=np.float32)
    out["size_residuals"] = (srn * mean_size[None, None]).astype(np.float32)         # :139
    hcls = out["heading_scores"].argmax(-1)                                         # :90
    hres = np.take_along_axis(out["heading_residuals"], hcls[..., None], 2)[..., 0]
    scls = out["size_scores"].argmax(-1)                                            # :94
    sres = np.take_along_axis(out["size_residuals"], scls[..., None, None].repeat(3, -1), 2)[:, :, 0]
    if heading_mode == "zero":
        angle = np.zeros(hcls.shape)
    else:
        angle = hcls * (2 * np.pi / float(nh)) + hres
        angle = np.where(angle > np.pi, angle - 2 * np.pi, angle)
    heading = -angle.astype(np.float64)
    size = (mean_size[scls] + sres).astype(np.float64)
    centre = out["center"].astype(np.float64)
    l, w, h = size[..., 0:1] / 2, size[..., 1:2] / 2, size[..., 2:3] / 2            # box_util.py:311-320
    c3 = np.stack([np.concatenate((l, l, -l, -l, l, l, -l, -l), -1),
                   np.concatenate((w, -w, -w, w, w, -w, -w, w), -1),
                   np.concatenate((h, h, h, h, -h, -h, -h, -h), -1)], -1)
    R = np.zeros(heading.shape + (3, 3))                                            # roty_batch :255-268
    R[..., 0, 0] = np.cos(heading); R[..., 0, 2] = np.sin(heading); R[..., 1, 1] = 1
    R[..., 2, 0] = -np.sin(heading); R[..., 2, 2] = np.cos(heading)
    out["bbox_corner"] = np.matmul(c3, np.swapaxes(R, -1, -2)) + centre[..., None, :]   # :321-324
    out["bbox_mask"] = out["objectness_scores"].argmax(-1)
    out["bbox_sems"] = out["sem_cls_scores"].argmax(-1)
    return out
