"""GPU parity of the layer / model level: bridgeqa_b200 modules vs oracle/modules_cpu.py
(the CPU restatement of the reference's SA / FP / backbone / voting / proposal forward,
itself pinned to the reference's Python layer by tests/golden/ref_python_layer.npz).

Tolerances (stated by BASELINE.json's north_star): indices bit-exact; features within
rel 1e-4 for the fp32/TF32-class paths and 1e-2 for the bf16 tensor-core path, measured
as max|a-b| / max|b| over the tensor.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import bridgeqa_b200  # noqa: E402
from bridgeqa_b200 import detector, pointnet2_modules as pm, synthetic  # noqa: E402
from oracle import modules_cpu  # noqa: E402


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(autouse=True)
def _fp32_convs():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _sd_cpu(module):
    return {k: v.detach().cpu() for k, v in module.state_dict().items()}


# (fused?, operand precision) -> tolerance on max|a-b| / max|b| for the WHOLE backbone (errors of
# 4 SA + 2 FP layers compound): fp32 torch path 1e-4; fp16 operands (TF32-class mantissa) 2e-3;
# bf16 operands 2e-2 (1e-2 per layer, see test_fused_sa_kernel_matches_unfused_fp32)
PATHS = [(False, "fp16", 1e-4), (True, "fp16", 2e-3), (True, "bf16", 2e-2)]


@pytest.fixture
def _restore_fused():
    yield
    bridgeqa_b200.set_fused(True)
    bridgeqa_b200.set_precision("fp16")


@pytest.mark.parametrize("fused,precision,tol", PATHS)
def test_backbone_matches_oracle(fused, precision, tol, _restore_fused):
    B, N, C = 2, 8192, 7
    pc = synthetic.make_batch(B, N, C, first_scene=40)
    net = synthetic.fill_state_dict(detector.Pointnet2Backbone(input_feature_dim=C), seed=3).cuda().eval()
    want = modules_cpu.backbone(pc.numpy(), _sd_cpu(net))
    bridgeqa_b200.set_fused(fused)
    bridgeqa_b200.set_precision(precision)
    with torch.no_grad():
        got = net({"point_clouds": pc.cuda()})
    for k in ("sa1_inds", "sa2_inds", "fp2_inds"):
        np.testing.assert_array_equal(got[k].cpu().numpy(), want[k], err_msg=k)
    for k in ("sa1_xyz", "sa2_xyz", "sa3_xyz", "sa4_xyz", "fp2_xyz"):
        np.testing.assert_array_equal(got[k].cpu().numpy(), want[k], err_msg=k)
    for k in ("sa1_features", "sa2_features", "sa3_features", "sa4_features", "fp2_features"):
        assert relerr(got[k].cpu().numpy(), want[k]) < tol, (k, relerr(got[k].cpu().numpy(), want[k]))


@pytest.mark.parametrize("fused,precision,tol", PATHS)
def test_detector_matches_oracle(fused, precision, tol, _restore_fused):
    B, N, C = 2, 6000, 132
    pc = synthetic.make_batch(B, N, C, first_scene=60)
    net = synthetic.fill_state_dict(detector.VoteNetDetector(C), seed=4).cuda().eval()
    want = modules_cpu.detector(pc.numpy(), _sd_cpu(net))
    bridgeqa_b200.set_fused(fused)
    bridgeqa_b200.set_precision(precision)
    with torch.no_grad():
        got = net({"point_clouds": pc.cuda()})
    np.testing.assert_array_equal(got["seed_inds"].cpu().numpy(), want["fp2_inds"])
    tol = 2 * tol
    assert relerr(got["vote_xyz"].cpu().numpy(), want["vote_xyz"]) < tol
    assert relerr(got["vote_features"].cpu().numpy(), want["vote_features"]) < tol
    # vote_xyz differs from the CPU's in the last bits (different conv summation order), and FPS
    # on it is chaotic, so vote-aggregation indices are checked for validity, not equality;
    # the aggregation layer itself is pinned by feeding the oracle the GPU's votes.
    inds = got["aggregated_vote_inds"].cpu().numpy()
    assert inds.shape == (B, 256) and inds.min() >= 0 and inds.max() < 1024
    xyz_a, feat_a, inds_a, _ = modules_cpu.sa_layer(
        got["vote_xyz"].cpu().numpy(), got["vote_features"].cpu().numpy(), _sd_cpu(net),
        "proposal_net.vote_aggregation.", 256, 0.3, 16)
    np.testing.assert_array_equal(inds, inds_a)
    np.testing.assert_array_equal(got["aggregated_vote_xyz"].cpu().numpy(), xyz_a)
    assert relerr(got["aggregated_vote_features"].cpu().numpy(), feat_a.transpose(0, 2, 1)) < tol
    assert got["bbox_corner"].shape == (B, 256, 8, 3)
    assert got["objectness_scores"].shape == (B, 256, 2)
    assert got["sem_cls_scores"].shape == (B, 256, 18)


def test_sa_module_accepts_precomputed_inds_and_returns_contract():
    torch.manual_seed(0)
    sa = pm.PointnetSAModuleVotes(npoint=64, radius=0.4, nsample=8, mlp=[5, 16, 16, 32],
                                  use_xyz=True, normalize_xyz=True).cuda().eval()
    xyz = synthetic.make_batch(2, 1500, 0)[..., :3].contiguous().cuda()
    feats = torch.randn(2, 5, 1500, device="cuda")
    with torch.no_grad():
        new_xyz, new_feats, inds = sa(xyz, feats)
        new_xyz2, new_feats2, inds2 = sa(xyz, feats, inds)
    assert new_xyz.shape == (2, 64, 3) and new_feats.shape == (2, 32, 64)
    assert inds.dtype == torch.int32 and inds.shape == (2, 64)
    assert torch.equal(new_xyz, new_xyz2) and torch.equal(inds, inds2)
    torch.testing.assert_close(new_feats, new_feats2)


def test_training_path_backward_runs_and_matches_torch_autograd():
    """Un-fused differentiable path: gradients w.r.t. features and weights equal those of
    a pure-torch re-expression (gather via indexing) of the same layer."""
    torch.manual_seed(1)
    sa = pm.PointnetSAModuleVotes(npoint=32, radius=0.5, nsample=8, mlp=[4, 8, 8, 16],
                                  use_xyz=True, normalize_xyz=True).cuda().train()
    xyz = synthetic.make_batch(2, 700, 0)[..., :3].contiguous().cuda()
    feats = torch.randn(2, 4, 700, device="cuda", requires_grad=True)
    new_xyz, out, inds = sa(xyz, feats)
    loss = (out * torch.linspace(0, 1, out.numel(), device="cuda").view_as(out)).sum()
    loss.backward()
    g_feats = feats.grad.clone()
    g_w = sa.mlp_module.layer0.conv.weight.grad.clone()

    sa.zero_grad()
    feats2 = feats.detach().clone().requires_grad_(True)
    from bridgeqa_b200 import ext
    idx = ext.ball_query(new_xyz, xyz, 0.5, 8).long()
    def grp(t):      # (B,C,N) -> (B,C,npoint,nsample) with plain torch.gather
        return torch.gather(t.unsqueeze(2).expand(-1, -1, idx.size(1), -1), 3,
                            idx.unsqueeze(1).expand(-1, t.size(1), -1, -1))

    gx = (grp(xyz.transpose(1, 2)) - new_xyz.transpose(1, 2).unsqueeze(-1)) / 0.5
    gf = grp(feats2)
    y = sa.mlp_module(torch.cat([gx, gf], 1)).max(-1).values
    (y * torch.linspace(0, 1, y.numel(), device="cuda").view_as(y)).sum().backward()
    torch.testing.assert_close(out, y, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(g_feats, feats2.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(g_w, sa.mlp_module.layer0.conv.weight.grad, rtol=1e-4, atol=1e-5)


def test_fp_module_matches_oracle():
    rng = np.random.default_rng(0)
    fp = synthetic.fill_state_dict(pm.PointnetFPModule(mlp=[24 + 16, 32, 32]), seed=5).cuda().eval()
    xyz = synthetic.make_batch(2, 900, 0)[..., :3].contiguous().numpy()
    unknown, known = xyz[:, :600].copy(), xyz[:, 600:900].copy()
    uf = rng.standard_normal((2, 16, 600)).astype(np.float32)
    kf = rng.standard_normal((2, 24, 300)).astype(np.float32)
    want, _ = modules_cpu.fp_layer(unknown, known, uf, kf, _sd_cpu(fp), "")
    bridgeqa_b200.set_fused(False)
    try:
        with torch.no_grad():
            got = fp(*(torch.from_numpy(a).cuda() for a in (unknown, known, uf, kf)))
    finally:
        bridgeqa_b200.set_fused(True)
    assert relerr(got.cpu().numpy(), want) < 1e-4


FUSED_SA_CASES = [
    # (B, N, C, npoint, radius, nsample, mlp)  -- all three kernel configs, every nsample, C edge cases
    (2, 5000, 7, 256, 0.3, 64, [64, 64, 128]),
    (2, 5000, 0, 128, 0.3, 64, [64, 64, 128]),
    (1, 4000, 132, 64, 0.3, 64, [64, 64, 128]),
    (2, 2048, 128, 256, 0.4, 32, [128, 128, 256]),
    (2, 1024, 256, 128, 0.8, 16, [128, 128, 256]),
    (3, 512, 256, 64, 1.2, 16, [128, 128, 256]),
    (2, 1024, 256, 256, 0.3, 16, [128, 128, 128]),
    (1, 3000, 13, 8, 0.5, 128, [64, 64, 128]),
    # feature chunk counts that do not fill the 8-chunk swizzle blocks of the operand buffer (10 = 8 + 2) and a
    # second pass of 8 chunks (24 = 16 + 8)
    (1, 2000, 83, 32, 0.5, 32, [128, 128, 256]),
    (1, 2000, 200, 32, 0.5, 32, [128, 128, 256]),
]


@pytest.mark.parametrize("precision,tol", [("fp16", 1e-3), ("bf16", 1e-2)])
@pytest.mark.parametrize("B,N,C,npoint,radius,nsample,mlp", FUSED_SA_CASES)
def test_fused_sa_kernel_matches_unfused_fp32(B, N, C, npoint, radius, nsample, mlp, precision, tol,
                                              _restore_fused):
    """tcgen05 fused SA (16-bit operands, fp32 accumulate) vs the un-fused fp32 path of the same
    module (torch conv/BN/ReLU/max on the grouped tensor).  Tolerance per layer, as a fraction
    of the tensor's max magnitude: 1e-2 for bf16 operands (north_star's bf16 bound), 1e-3 for
    fp16 operands (TF32-class mantissa); indices/centres identical by construction."""
    bridgeqa_b200.set_precision(precision)
    sa = pm.PointnetSAModuleVotes(npoint=npoint, radius=radius, nsample=nsample, mlp=[C] + mlp,
                                  use_xyz=True, normalize_xyz=True)
    sa = synthetic.fill_state_dict(sa, seed=11).cuda().eval()
    pc = synthetic.make_batch(B, N, 0, first_scene=80)
    xyz = pc[..., :3].contiguous().cuda()
    feats = torch.randn(B, C, N, device="cuda") if C > 0 else None
    with torch.no_grad():
        bridgeqa_b200.set_fused(False)
        try:
            ref_xyz, ref_feats, ref_inds = sa(xyz, feats)
        finally:
            bridgeqa_b200.set_fused(True)
        from bridgeqa_b200 import _native
        before = _native.launch_count()
        got_xyz, got_feats, got_inds = sa(xyz, feats)
        assert _native.launch_count() - before >= 3          # fps, ball query, fused MLP (+pack/transposes)
    assert torch.equal(ref_inds, got_inds) and torch.equal(ref_xyz, got_xyz)
    assert got_feats.shape == ref_feats.shape == (B, mlp[-1], npoint)
    err = relerr(got_feats.cpu().numpy(), ref_feats.cpu().numpy())
    assert err < tol, err
    # the point-major twin the next layer consumes
    assert torch.equal(got_feats._bqa_pm, got_feats.transpose(1, 2).contiguous())


def test_fused_cache_follows_weight_updates():
    sa = pm.PointnetSAModuleVotes(npoint=64, radius=0.4, nsample=16, mlp=[4, 64, 64, 128],
                                  use_xyz=True, normalize_xyz=True)
    sa = synthetic.fill_state_dict(sa, seed=1).cuda().eval()
    xyz = synthetic.make_batch(1, 1000, 0)[..., :3].contiguous().cuda()
    feats = torch.randn(1, 4, 1000, device="cuda")
    with torch.no_grad():
        a = sa(xyz, feats)[1]
        synthetic.fill_state_dict(sa, seed=2)
        b = sa(xyz, feats)[1]
        bridgeqa_b200.set_fused(False)
        try:
            c = sa(xyz, feats)[1]
        finally:
            bridgeqa_b200.set_fused(True)
    assert not torch.allclose(a, b)
    assert relerr(b.cpu().numpy(), c.cpu().numpy()) < 1e-2


FUSED_FP_CASES = [
    # (B, n, m, c_known, c_skip)
    (2, 512, 256, 256, 256),
    (2, 1024, 512, 256, 256),
    (1, 300, 700, 128, 64),       # ragged tile, known set larger than one smem pass, other widths
    (2, 129, 2, 64, 64),          # fewer than 3 known points: infinite distances -> zero weights
]


@pytest.mark.parametrize("precision,tol", [("fp16", 1e-3), ("bf16", 1e-2)])
@pytest.mark.parametrize("B,n,m,ck,cs", FUSED_FP_CASES)
def test_fused_fp_kernel_matches_unfused_fp32(B, n, m, ck, cs, precision, tol, _restore_fused):
    """Fused FP (three_nn + interpolate + concat + 2-layer MLP on tcgen05) vs the un-fused fp32
    path of the same module.  Neighbour indices are bit-identical by construction (same cascade);
    features within 1e-3 (fp16 operands) / 1e-2 (bf16) of the tensor's max."""
    bridgeqa_b200.set_precision(precision)
    fp = synthetic.fill_state_dict(pm.PointnetFPModule(mlp=[ck + cs, 256, 256]), seed=21).cuda().eval()
    xyz = synthetic.make_batch(B, n + m, 0, first_scene=85)[..., :3].contiguous().cuda()
    unknown, known = xyz[:, :n].contiguous(), xyz[:, n:].contiguous()
    uf = torch.randn(B, cs, n, device="cuda")
    kf = torch.randn(B, ck, m, device="cuda")
    with torch.no_grad():
        bridgeqa_b200.set_fused(False)
        ref = fp(unknown, known, uf, kf)
        bridgeqa_b200.set_fused(True)
        from bridgeqa_b200 import _native
        before = _native.launch_count()
        got = fp(unknown, known, uf, kf)
        assert _native.launch_count() - before <= 5      # transposes (no twins here) + packs + 1 fused kernel
    assert got.shape == ref.shape == (B, 256, n)
    err = relerr(got.cpu().numpy(), ref.cpu().numpy())
    assert err < tol, err
    assert torch.equal(got._bqa_pm, got.transpose(1, 2).contiguous())


def test_fused_fp_kernel_known_rows_not_32_byte_aligned(_restore_fused):
    """The FP kernel reads 32 bytes of a neighbour row per lane with one 256-bit load when the point-major known
    features allow it; rows that are only 16-byte aligned (a 260-float pitch, first row 16 bytes into the
    buffer) take two 128-bit loads.  Same bits either way."""
    bridgeqa_b200.set_precision("fp16")
    bridgeqa_b200.set_fused(True)
    B, n, m = 2, 300, 200
    fp = synthetic.fill_state_dict(pm.PointnetFPModule(mlp=[256 + 128, 256, 256]), seed=22).cuda().eval()
    xyz = synthetic.make_batch(B, n + m, 0, first_scene=87)[..., :3].contiguous().cuda()
    unknown, known = xyz[:, :n].contiguous(), xyz[:, n:].contiguous()
    uf = torch.randn(B, 128, n, device="cuda")
    kf = torch.randn(B, 256, m, device="cuda")
    with torch.no_grad():
        want = fp(unknown, known, uf, kf)                  # contiguous point-major twin: 256-bit loads
        pitched = torch.zeros(B, m, 260, device="cuda")
        twin = pitched[..., 4:]                            # (B, m, 256) view, pitch 260 floats, +16 bytes
        twin.copy_(kf.transpose(1, 2))
        assert twin.data_ptr() % 32 == 16 and twin.stride(1) % 8 == 4
        kf2 = kf.clone()
        kf2._bqa_pm = twin
        got = fp(unknown, known, uf, kf2)
    assert torch.equal(got, want)


def test_detector_train_step_backward_reaches_every_parameter():
    """configs[3] on one GPU: train-mode forward + backward through gather / group /
    three_interpolate gradient kernels and the MLPs; all parameters get finite gradients and
    BN running statistics move."""
    from bridgeqa_b200 import training
    torch.manual_seed(0)
    net = synthetic.fill_state_dict(detector.VoteNetDetector(7), seed=6).cuda()
    loss_fn = training.ProjectionLoss().cuda()
    pc = synthetic.make_batch(2, 5000, 7, first_scene=33).cuda()
    rm0 = net.detection_backbone.sa1.mlp_module.layer0.bn.bn.running_mean.clone()
    from bridgeqa_b200 import _native
    before = _native.launch_count()
    loss = training.train_step(net, loss_fn, pc)
    assert torch.isfinite(loss)
    assert _native.launch_count() - before >= 25       # fwd ops + the *_grad kernels
    missing = [n for n, p in net.named_parameters() if p.grad is None]
    assert not missing, missing
    assert all(torch.isfinite(p.grad).all() for p in net.parameters())
    assert sum(float(p.grad.abs().sum()) for p in net.parameters()) > 0
    assert not torch.equal(rm0, net.detection_backbone.sa1.mlp_module.layer0.bn.bn.running_mean)


def test_fp_module_backward_matches_torch_autograd():
    torch.manual_seed(2)
    fp = pm.PointnetFPModule(mlp=[12 + 8, 16, 16]).cuda().train()
    xyz = synthetic.make_batch(2, 500, 0)[..., :3].contiguous().cuda()
    unknown, known = xyz[:, :300].contiguous(), xyz[:, 300:].contiguous()
    uf = torch.randn(2, 8, 300, device="cuda", requires_grad=True)
    kf = torch.randn(2, 12, 200, device="cuda", requires_grad=True)
    out = fp(unknown, known, uf, kf)
    w = torch.linspace(-1, 1, out.numel(), device="cuda").view_as(out)
    (out * w).sum().backward()
    g_uf, g_kf = uf.grad.clone(), kf.grad.clone()
    # torch-only re-expression of the interpolation
    from bridgeqa_b200 import pointnet2_utils as pu
    dist, idx = pu.three_nn(unknown, known)
    rec = 1.0 / (dist + 1e-8)
    wt = rec / rec.sum(2, keepdim=True)
    uf2, kf2 = uf.detach().clone().requires_grad_(True), kf.detach().clone().requires_grad_(True)
    gathered = torch.gather(kf2.unsqueeze(2).expand(-1, -1, 300, -1), 3,
                            idx.long().unsqueeze(1).expand(-1, 12, -1, -1))          # (B,C,n,3)
    interp = (gathered * wt.unsqueeze(1)).sum(-1)
    fp.zero_grad()
    ref = fp.mlp(torch.cat([interp, uf2], 1).unsqueeze(-1)).squeeze(-1)
    (ref * w).sum().backward()
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(g_uf, uf2.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(g_kf, kf2.grad, rtol=1e-3, atol=1e-5)


@pytest.mark.parametrize("which,c", [("backbone", 7), ("detector", 132)])
def test_cuda_graph_replay_equals_eager(which, c, _restore_fused):
    """graphs.GraphedForward: replaying the captured forward (three streams, ~35 launches) gives
    bit-identical results to issuing it eagerly, follows new inputs, new input buffers
    (bind_inputs) and weight updates."""
    make = (lambda: detector.Pointnet2Backbone(input_feature_dim=c)) if which == "backbone" \
        else (lambda: detector.VoteNetDetector(c))
    keys = ["fp2_features", "fp2_inds", "sa4_features"] if which == "backbone" \
        else ["bbox_corner", "objectness_scores", "aggregated_vote_features", "aggregated_vote_inds"]
    net = synthetic.fill_state_dict(make(), seed=0).cuda().eval()
    pcs = [synthetic.make_batch(2, 20000, c, first_scene=3 * i).cuda() for i in range(3)]
    with torch.no_grad():
        eager = [{k: net({"point_clouds": p})[k].clone() for k in keys} for p in pcs]
        for bind in (False, True):
            net.enable_cuda_graph(True, bind_inputs=bind)
            for rep in range(2):
                for p, want in zip(pcs, eager):
                    out = net({"point_clouds": p})
                    for k in keys:
                        assert torch.equal(out[k], want[k]), (bind, rep, k)
            assert net._graphed.replays == 6
            assert len(net._graphed.cache) == (3 if bind else 1)
        # a weight update must not replay stale folded weights
        with torch.no_grad():
            net_params = [p for p in net.parameters()]
            net_params[0].mul_(1.5)
        out = {k: v.clone() for k, v in net({"point_clouds": pcs[0]}).items() if k in keys}
        net.enable_cuda_graph(False)
        ref = net({"point_clouds": pcs[0]})
        for k in keys:
            assert torch.equal(out[k], ref[k]), k
        assert not torch.equal(out[keys[0]], eager[0][keys[0]])


def test_training_sampling_prefetch_changes_nothing():
    """Pointnet2Backbone.prefetch_sampling / train_step(next_point_clouds=...): SA1's sampling of
    the next batch issued under this step's backward is picked up by the next forward (same
    tensor only) and gives the same indices, loss and gradients as sampling in place."""
    import copy
    from bridgeqa_b200 import training
    torch.manual_seed(0)
    net_a = synthetic.fill_state_dict(detector.VoteNetDetector(4), seed=0).cuda()
    net_b = copy.deepcopy(net_a)
    loss_fn = training.ProjectionLoss().cuda()
    pcs = [synthetic.make_batch(2, 20000, 4, first_scene=7 * i).cuda() for i in range(3)]
    for i, pc in enumerate(pcs):
        nxt = pcs[i + 1] if i + 1 < len(pcs) else None
        la = training.train_step(net_a, loss_fn, pc, next_point_clouds=nxt)
        bb = net_a.detection_backbone
        assert (getattr(bb, "_prefetched", None) is not None) == (nxt is not None)
        lb = training.train_step(net_b, loss_fn, pc)
        torch.testing.assert_close(la, lb, rtol=1e-4, atol=1e-6)
        for (n1, p1), (_, p2) in zip(net_a.named_parameters(), net_b.named_parameters()):
            if p1.grad is None:
                assert p2.grad is None
                continue
            scale = float(p2.grad.abs().max()) + 1e-12
            assert float((p1.grad - p2.grad).abs().max()) <= 2e-3 * scale, n1
    # a prefetch for another tensor is dropped, not used
    net_a.detection_backbone.prefetch_sampling(pcs[0])
    net_a.eval()
    with torch.no_grad():
        ia = net_a({"point_clouds": pcs[1]})["sa1_inds"]
        ib = net_b.eval()({"point_clouds": pcs[1]})["sa1_inds"]
    assert torch.equal(ia, ib)


@pytest.mark.parametrize("graph", [False, True])
@pytest.mark.parametrize("depth", [1, 2, 3])
def test_in_flight_forwards_equal_serial(graph, depth, _restore_fused):
    """graphs.InFlight: forwards issued on a ring of streams (the next batch's sampling chain
    under this batch's SA/FP kernels) return exactly what one-at-a-time forwards return."""
    net = synthetic.fill_state_dict(detector.Pointnet2Backbone(input_feature_dim=4), seed=0).cuda().eval()
    keys = ["fp2_features", "fp2_inds", "sa1_inds", "sa4_features"]
    pcs = [synthetic.make_batch(2, 20000, 4, first_scene=5 * i).cuda() for i in range(4)]
    with torch.no_grad():
        want = [{k: net({"point_clouds": p})[k].clone() for k in keys} for p in pcs]
    if graph:
        net.enable_cuda_graph(True, bind_inputs=True)
    q = net.in_flight(depth)
    for rep in range(2):
        tickets = [q.submit({"point_clouds": p}) for p in pcs]
        for t, w in zip(tickets, want):
            out = t.wait()
            for k in keys:
                assert torch.equal(out[k], w[k]), (graph, depth, rep, k)
    q.drain()
    torch.cuda.synchronize()
    assert q.submitted == 8
    with pytest.raises(RuntimeError):
        q.submit({"point_clouds": pcs[0].cpu()})
    net.enable_cuda_graph(False)


@pytest.mark.parametrize("B,cin,spec,npoint,ns", [
    (2, 6, [16, 16], 64, 16), (3, 10, [64, 64, 128], 128, 32), (2, 12, [128], 33, 64),
    (2, 5, [8, 8], 10, 4), (2, 7, [16, 24], 20, 6),          # ns = 6: pooled falls back to max_pool2d
])
def test_train_bn_relu_max_matches_torch_modules(B, cin, spec, npoint, ns):
    """csrc/bn_relu.cu behind SharedMLP in model.train(): outputs, running statistics and every
    gradient equal those of the torch BatchNorm2d + ReLU + max_pool2d path."""
    import copy
    import torch.nn.functional as F
    from bridgeqa_b200 import pytorch_utils as pt, train_fused
    torch.manual_seed(3)
    mlp_a = pt.SharedMLP([cin] + spec, bn=True).cuda().train()
    for m in mlp_a.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.uniform_(-0.5, 0.5)
            m.momentum = 0.3
    mlp_b = copy.deepcopy(mlp_a)
    x = torch.randn(B, cin, npoint, ns, device="cuda")
    x[..., ns // 2:] = x[..., :ns - ns // 2]                 # repeated neighbours: ties for the argmax
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    w_full = torch.randn(B, spec[-1], npoint, ns, device="cuda")
    w_pool = torch.randn(B, spec[-1], npoint, device="cuda")

    for pooled in (False, True):
        for net in (mlp_a, mlp_b):
            net.zero_grad()
        xa.grad = xb.grad = None
        assert train_fused.enabled()
        out_a = mlp_a.forward_pooled(xa) if pooled else mlp_a(xa)
        train_fused.set_enabled(False)
        try:
            full_b = mlp_b(xb)
            out_b = F.max_pool2d(full_b, kernel_size=[1, ns]).squeeze(-1) if pooled else full_b
        finally:
            train_fused.set_enabled(True)
        w = w_pool if pooled else w_full
        (out_a * w).sum().backward()
        (out_b * w).sum().backward()
        torch.testing.assert_close(out_a, out_b, rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(xa.grad, xb.grad, rtol=2e-3, atol=2e-5)
        for (na, pa), (nb, pb) in zip(mlp_a.named_parameters(), mlp_b.named_parameters()):
            torch.testing.assert_close(pa.grad, pb.grad, rtol=2e-3, atol=1e-4, msg=lambda m: na + ": " + m)
        for (na, ba), (nb, bb) in zip(mlp_a.named_buffers(), mlp_b.named_buffers()):
            torch.testing.assert_close(ba, bb, rtol=1e-5, atol=1e-6, msg=lambda m: na + ": " + m)


def test_train_bn_relu_on_1d_rows_and_odd_lengths():
    """FP-layer shape (B, C, n, 1) with n not a multiple of 4 (scalar path of the kernels)."""
    import copy
    from bridgeqa_b200 import pytorch_utils as pt, train_fused
    torch.manual_seed(4)
    a = pt.SharedMLP([9, 12, 12], bn=True).cuda().train()
    b = copy.deepcopy(a)
    x = torch.randn(3, 9, 77, 1, device="cuda")
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya = a(xa)
    train_fused.set_enabled(False)
    try:
        yb = b(xb)
    finally:
        train_fused.set_enabled(True)
    w = torch.randn_like(ya)
    (ya * w).sum().backward()
    (yb * w).sum().backward()
    torch.testing.assert_close(ya, yb, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(xa.grad, xb.grad, rtol=2e-3, atol=2e-5)
    for pa, pb in zip(a.parameters(), b.parameters()):
        torch.testing.assert_close(pa.grad, pb.grad, rtol=2e-3, atol=1e-4)
