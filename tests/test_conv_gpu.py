"""tcgen05 TF32 1x1 convolutions of the training path (csrc/conv_tf32.cu) against fp32 torch.

Reference call sites: lib/pointnet2/pytorch_utils.py:104-157 (nn.Conv2d, kernel 1, no bias, then
BatchNorm2d + ReLU).  Tolerance: TF32 operands (10-bit mantissa, rounded by the TMA unit), fp32
accumulation -- errors are measured against an fp64 contraction and compared with what cuDNN's TF32
convolution (torch's default, the reference's arithmetic) does on the same inputs.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

from bridgeqa_b200 import train_fused  # noqa: E402
from bridgeqa_b200 import pytorch_utils as pt_utils  # noqa: E402

CASES = [
    # b, cin, cout, shape of the positions
    (2, 135, 64, (64, 32)),       # SA1 layer 1 (xyz + 132 features): ragged K, padded weight rows
    (2, 64, 64, (128, 16)),
    (2, 64, 128, (256, 16)),
    (3, 131, 128, (32, 32)),
    (2, 128, 256, (32, 16)),      # two M halves
    (2, 259, 128, (16, 16)),      # K tail of 3
    (2, 512, 256, (512,)),        # FP layer 1 (1-D positions)
    (1, 256, 256, (100,)),        # partial tile (100 positions)
    (2, 7, 16, (12,)),            # tiny
]


def _rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("b,cin,cout,pos", CASES)
def test_conv_forward_stats_dgrad_wgrad(b, cin, cout, pos):
    g = torch.Generator(device="cpu").manual_seed(cin * 1000 + cout)
    x = torch.randn((b, cin) + pos, generator=g).cuda()
    w = (torch.randn(cout, cin, generator=g) / cin ** 0.5).cuda()
    dy = torch.randn((b, cout) + pos, generator=g).cuda()
    p = x[0, 0].numel()
    xr, dyr = x.reshape(b, cin, p).double(), dy.reshape(b, cout, p).double()
    want = torch.einsum("oc,bcp->bop", w.double(), xr)
    shift = torch.randn(cout, generator=g).cuda()
    sums = torch.zeros(2 * cout, dtype=torch.float64, device="cuda")
    y = train_fused._conv_forward(x, w, shift, sums)
    assert y.shape == (b, cout) + pos
    err = _rel(y.reshape(b, cout, p), want)
    # cuDNN / cuBLAS TF32 on the same inputs
    torch.backends.cuda.matmul.allow_tf32 = True
    ref_tf32 = torch.matmul(w, x.reshape(b, cin, p))
    torch.backends.cuda.matmul.allow_tf32 = False
    err_ref = _rel(ref_tf32, want)
    assert err < 2e-3, "forward rel err %g" % err
    # (the library falls back to fp32 for tiny shapes: then only the absolute bar applies)
    assert err <= max(2.0 * err_ref, 1e-3), "forward rel err %g vs library TF32 %g" % (err, err_ref)
    # statistics of the stored values, around the shift
    d = y.reshape(b, cout, p).double() - shift.double()[None, :, None]
    torch.testing.assert_close(sums[:cout], d.sum((0, 2)), rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(sums[cout:], (d * d).sum((0, 2)), rtol=1e-5, atol=1e-4)
    # data gradient: dx = W^T dy
    dx = train_fused._conv_forward(dy, w.t())
    want_dx = torch.einsum("oc,bop->bcp", w.double(), dyr)
    assert _rel(dx.reshape(b, cin, p), want_dx) < 2e-3
    # weight gradient
    dw = train_fused._conv_wgrad(x, dy, cout, cin)
    want_dw = torch.einsum("bop,bcp->oc", dyr, xr)
    assert _rel(dw, want_dw) < 2e-3


def tf32(t):
    """round to TF32 (10-bit mantissa, nearest, ties away) -- what the TFLOAT32 tensor map does on load"""
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1fff).view(torch.float32)


class _RefConv(torch.autograd.Function):
    """The arithmetic model of csrc/conv_tf32.cu in plain torch: operands rounded to TF32, fp32 math."""

    @staticmethod
    def forward(ctx, x, w):
        ctx.save_for_backward(x, w)
        return torch.einsum("oc,bcp->bop", tf32(w.reshape(w.shape[0], -1)), tf32(x).flatten(2)).reshape(
            (x.shape[0], w.shape[0]) + tuple(x.shape[2:]))

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        w2, dyr, xr = tf32(w.reshape(w.shape[0], -1)), tf32(dy).flatten(2), tf32(x).flatten(2)
        dx = torch.einsum("oc,bop->bcp", tf32(w2.t().contiguous()).t(), dyr).reshape(x.shape)
        dw = torch.einsum("bop,bcp->oc", dyr, xr).reshape(w.shape)
        return dx, dw


def _ref_forward(mlp, x, pooled):
    """SharedMLP in train(): conv (TF32 model) -> torch BatchNorm2d (batch statistics) -> ReLU [-> max]"""
    for blk in mlp.children():
        mods = list(blk.children())
        x = torch.relu(mods[1](_RefConv.apply(x, mods[0].weight)))
    return x.max(dim=3)[0] if pooled else x


def test_conv_block_matches_torch_modules_forward_and_backward():
    """A SharedMLP in train() on the fused conv + BatchNorm node vs the same torch modules around the
    arithmetic model of the conv (operands rounded to TF32, fp32 contraction): outputs, input gradient,
    every parameter gradient, running statistics.  Outputs and parameter gradients to 1e-3 of each tensor's
    max.  The input gradient in the L2 norm (1e-2) with at most 0.5 % of its elements off by more than 1e-2 of
    the max: what is left between the two is the accumulation order, and a pre-activation within ~1e-5 of zero
    (or of another sample under the max) can land on the other side of the ReLU / max decision, which switches
    that unit's whole contribution at that position on or off."""
    torch.manual_seed(0)
    torch.backends.cudnn.allow_tf32 = True           # the fused node follows torch's TF32 switch
    torch.backends.cuda.matmul.allow_tf32 = False    # the arithmetic model's einsums stay fp32
    mlp_a = pt_utils.SharedMLP([67, 64, 64, 128], bn=True).cuda().train()
    mlp_b = pt_utils.SharedMLP([67, 64, 64, 128], bn=True).cuda().train()
    mlp_b.load_state_dict(mlp_a.state_dict())
    for m in list(mlp_a.modules()) + list(mlp_b.modules()):
        if isinstance(m, torch.nn.BatchNorm2d):
            m.momentum = 0.3
    x = torch.randn(2, 67, 128, 16, device="cuda") * 2 + 0.5

    def close(a, b, what, tol=1e-3):
        err = float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))
        assert err < tol, "%s: %g of the tensor's max" % (what, err)
    for pooled in (False, True):
        mlp_a.zero_grad(); mlp_b.zero_grad()
        wgt = torch.randn((2, 128, 128) + (() if pooled else (16,)), device="cuda")
        xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        assert train_fused.conv_enabled() and train_fused.enabled()
        out_a = mlp_a.forward_pooled(xa) if pooled else mlp_a(xa)
        (out_a * wgt).sum().backward()
        out_b = _ref_forward(mlp_b, xb, pooled)
        (out_b * wgt).sum().backward()
        close(out_a, out_b, "output (pooled=%s)" % pooled)
        diff = (xa.grad - xb.grad)
        l2 = float(diff.norm() / xb.grad.norm())
        bad = float((diff.abs() > 1e-2 * xb.grad.abs().max()).float().mean())
        assert l2 < 1e-2 and bad < 5e-3, "input gradient (pooled=%s): L2 %g, %g of the elements off" % (pooled, l2, bad)
        # (measured: 4 of the 524288 output units sit on the other side of the ReLU; each switches one whole
        #  term of its channel's sums on or off -- up to 4 % of the largest weight gradient in 4 of 128 rows --
        #  so parameter gradients are held in the L2 norm, with few elements off)
        for (na, pa), (nb, pb) in zip(mlp_a.named_parameters(), mlp_b.named_parameters()):
            d = pa.grad - pb.grad
            l2p = float(d.norm() / pb.grad.norm())
            badp = float((d.abs() > 1e-2 * pb.grad.abs().max()).float().mean())
            assert l2p < 2e-2 and badp < 5e-2, "%s (pooled=%s): L2 %g, %g of the elements off" % (na, pooled, l2p, badp)
        for (na, ba), (nb, bb) in zip(mlp_a.named_buffers(), mlp_b.named_buffers()):
            torch.testing.assert_close(ba.float(), bb.float(), rtol=1e-4, atol=1e-4, msg=na)


def test_detector_train_step_runs_on_the_tcgen05_convolutions():
    """configs[3] on one GPU with torch's default TF32 switch: every SharedMLP block of the SA / FP layers
    goes through bqa_conv1x1_tf32_forward / _wgrad (counted), the loss agrees with the same step on cuDNN's
    TF32 convolutions, every parameter gets a finite gradient, BN running statistics move identically."""
    from bridgeqa_b200 import detector, profiler, synthetic, training
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = True
    try:
        results = []
        for fused_conv in (True, False):
            torch.manual_seed(0)
            net = synthetic.fill_state_dict(detector.VoteNetDetector(7), seed=6).cuda()
            loss_fn = training.ProjectionLoss().cuda()
            pc = synthetic.make_batch(2, 5000, 7, first_scene=33).cuda()
            train_fused.set_conv_enabled(fused_conv)
            net.train()
            with profiler.KernelTimer() as kt:
                out = net({"point_clouds": pc})
                loss = loss_fn(out)
                loss.backward()
            names = [r[0] for r in kt.records]
            results.append((float(loss), names, net, out["fp2_features"].detach()))
        (loss_a, names_a, net_a, feat_a), (loss_b, names_b, net_b, feat_b) = results
    finally:
        train_fused.set_conv_enabled(True)
        torch.backends.cudnn.allow_tf32 = old
    # SA1-4 (3 blocks each), FP1-2 (2 each), vote aggregation (3): 19 forward calls, 19 weight gradients,
    # and a data gradient for every block but SA1's first (its input is leaf data)
    assert names_a.count("bqa_conv1x1_tf32_wgrad") == 19
    assert names_a.count("bqa_conv1x1_tf32_forward") == 19 + 18
    assert "bqa_conv1x1_tf32_forward" not in names_b
    # backbone output: same arithmetic class as cuDNN's TF32 convolutions, ten layers deep.  (The loss also
    # sees the vote aggregation, whose sampling of the predicted votes can pick other indices when the votes
    # move by 1e-3: only its order of magnitude is compared.)
    err = float((feat_a - feat_b).abs().max() / feat_b.abs().max())
    assert err < 3e-2, "fp2_features: %g of the tensor's max vs cuDNN TF32" % err
    assert abs(loss_a - loss_b) <= 0.25 * abs(loss_b), (loss_a, loss_b)
    missing = [n for n, p in net_a.named_parameters() if p.grad is None]
    assert not missing, missing
    assert all(torch.isfinite(p.grad).all() for p in net_a.parameters())
    rm_a = net_a.detection_backbone.sa2.mlp_module.layer1.bn.bn.running_mean
    rm_b = net_b.detection_backbone.sa2.mlp_module.layer1.bn.bn.running_mean
    torch.testing.assert_close(rm_a, rm_b, rtol=2e-2, atol=2e-3)


def test_graphed_train_step_equals_eager_steps():
    """training.GraphedTrainStep (two alternating CUDA graphs, the next batch's sampling produced by the
    current replay) against the same steps issued eagerly: loss, gradients and BatchNorm running statistics
    over four steps that alternate between batches, announced and unannounced.  Backbone only: the detector's
    vote aggregation re-samples PREDICTED coordinates, where a 1e-6 difference (library algorithm choice
    under capture) can pick another vote."""
    from bridgeqa_b200 import detector, synthetic, training
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = True
    try:
        nets = []
        for _ in range(2):
            torch.manual_seed(0)
            nets.append(synthetic.fill_state_dict(detector.Pointnet2Backbone(input_feature_dim=7), seed=6).cuda())
        w = torch.randn(256, 1024, device="cuda") / 32.0

        def loss_fn(out):
            return (out["fp2_features"] * w).pow(2).mean() + out["sa4_features"].pow(2).mean()
        batches = [synthetic.make_batch(2, 5000, 7, first_scene=40 + 2 * i).cuda() for i in range(3)]
        order = [0, 1, 2, 1]
        graphed = training.GraphedTrainStep(nets[0], loss_fn, batches[0])
        for step, bi in enumerate(order):
            nxt = batches[order[step + 1]] if step + 1 < len(order) and step != 1 else None   # step 2 comes unannounced
            loss_g = graphed(batches[bi], nxt)
            loss_e = training.train_step(nets[1], loss_fn, batches[bi])
            assert abs(float(loss_g) - float(loss_e)) <= 1e-4 * abs(float(loss_e)) + 1e-6, (step, float(loss_g), float(loss_e))
            for (na, pa), (nb, pb) in zip(nets[0].named_parameters(), nets[1].named_parameters()):
                scale = float(pb.grad.abs().max()) + 1e-12
                assert float((pa.grad - pb.grad).abs().max()) <= 2e-3 * scale, (step, na)
        for (na, ba), (nb, bb) in zip(nets[0].named_buffers(), nets[1].named_buffers()):
            torch.testing.assert_close(ba.float(), bb.float(), rtol=1e-4, atol=1e-5, msg=na)
    finally:
        torch.backends.cudnn.allow_tf32 = old


def test_train_step_at_the_benchmarked_size_graphed_vs_eager_vs_cudnn():
    """configs[3] as bench.py times it: VoteNetDetector, C = 132, 16 scenes x 40000 points, train-mode BN,
    torch's default TF32 switch.  (a) The step replayed from the two alternating CUDA graphs gives the eager
    step's loss (1e-4) for two consecutive, announced batches; (b) the eager step runs every SharedMLP block
    on the tcgen05 kernels (call counts); (c) its backbone output agrees with the same step on cuDNN's TF32
    convolutions to 3e-2 of the tensor's max (ten TF32 layers deep, through BatchNorm)."""
    from bridgeqa_b200 import detector, profiler, synthetic, training
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = True
    try:
        feats, bsz = 132, 16
        batches = [synthetic.make_batch(bsz, 40000, feats, first_scene=300 + bsz * i).cuda() for i in range(2)]
        loss_fn = training.ProjectionLoss().cuda()

        def fresh():
            torch.manual_seed(0)
            return synthetic.fill_state_dict(detector.VoteNetDetector(feats), seed=0).cuda()
        # eager, tcgen05 convs
        net_e = fresh()
        net_e.train()
        with profiler.KernelTimer() as kt:
            out = net_e({"point_clouds": batches[0]})
            loss_e0 = loss_fn(out)
            loss_e0.backward()
        names = [r[0] for r in kt.records]
        assert names.count("bqa_conv1x1_tf32_wgrad") == 19 and names.count("bqa_conv1x1_tf32_forward") == 37
        feat_e = out["fp2_features"].detach().clone()
        loss_e1 = training.train_step(net_e, loss_fn, batches[1])
        # graphed
        net_g = fresh()
        graphed = training.GraphedTrainStep(net_g, loss_fn, batches[0])
        loss_g0 = graphed(batches[0], batches[1])
        loss_g1 = graphed(batches[1], None)
        for a, b_ in ((loss_g0, loss_e0), (loss_g1, loss_e1)):
            assert abs(float(a) - float(b_)) <= 1e-4 * abs(float(b_)) + 1e-7, (float(a), float(b_))
        del graphed, net_g
        # cuDNN TF32 convs
        net_c = fresh()
        net_c.train()
        train_fused.set_conv_enabled(False)
        try:
            out_c = net_c({"point_clouds": batches[0]})
        finally:
            train_fused.set_conv_enabled(True)
        err = float((feat_e - out_c["fp2_features"]).abs().max() / out_c["fp2_features"].abs().max())
        assert err < 3e-2, "fp2_features vs cuDNN TF32: %g of the tensor's max" % err
    finally:
        torch.backends.cudnn.allow_tf32 = old
        torch.cuda.empty_cache()


def test_graphed_train_step_recaptures_when_bn_momentum_changes():
    """BNMomentumScheduler (lib/solver.py:271-279, pytorch_utils.py:299-333) rewrites every BatchNorm's momentum
    per epoch; the value is baked into the captured graphs, so GraphedTrainStep must notice and re-capture:
    running statistics after the change follow the NEW momentum exactly like the eager step's."""
    from bridgeqa_b200 import detector, synthetic, training
    from bridgeqa_b200 import pytorch_utils as pu
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = True
    try:
        nets = []
        for _ in range(2):
            torch.manual_seed(0)
            nets.append(synthetic.fill_state_dict(detector.Pointnet2Backbone(input_feature_dim=7), seed=6).cuda())
        loss_fn = lambda out: out["fp2_features"].pow(2).mean()
        pc = synthetic.make_batch(2, 5000, 7, first_scene=77).cuda()
        graphed = training.GraphedTrainStep(nets[0], loss_fn, pc)
        scheds = [pu.BNMomentumScheduler(n, bn_lambda=lambda e: 0.5 * 0.5 ** e) for n in nets]
        for epoch in range(2):
            for s in scheds:
                s.step(epoch)
            graphed(pc, pc)
            training.train_step(nets[1], loss_fn, pc)
        rm_g = nets[0].sa1.mlp_module.layer0.bn.bn.running_mean
        rm_e = nets[1].sa1.mlp_module.layer0.bn.bn.running_mean
        assert float(nets[0].sa1.mlp_module.layer0.bn.bn.momentum) == 0.25
        torch.testing.assert_close(rm_g, rm_e, rtol=1e-4, atol=1e-5)
        nb = nets[0].sa1.mlp_module.layer0.bn.bn.num_batches_tracked
        assert int(nb) == int(nets[1].sa1.mlp_module.layer0.bn.bn.num_batches_tracked)
    finally:
        torch.backends.cudnn.allow_tf32 = old
