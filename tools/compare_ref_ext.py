"""Time the reference's own CUDA extension (oracle/_ref) and the sm_100a kernels on the
same inputs, op by op, at the BASELINE sizes (B=16, N=40000).  Measurement tool for
BASELINE.md section 5 / profiles/; not part of the product and not part of bench.py.

    gpurun -- 'python tools/compare_ref_ext.py > gpurun_out/compare_ref_ext.json'
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bridgeqa_b200 import ext, synthetic  # noqa: E402
from oracle import ref_ext  # noqa: E402


def timeit(fn, warm=3, it=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(it):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    times.sort()
    return times[len(times) // 2]


def main():
    ref = ref_ext.load()
    B = 16
    pc = synthetic.make_batch(B, 40000, 7).cuda()
    xyz = pc[..., :3].contiguous()
    feats = pc[..., 3:].transpose(1, 2).contiguous()
    rows = []

    def add(name, ours, theirs, equal=None):
        r = {"op": name, "b200_ms": round(timeit(ours), 4)}
        if ref is not None:
            r["ref_ext_ms"] = round(timeit(theirs, warm=1, it=3), 4)
            r["speedup"] = round(r["ref_ext_ms"] / r["b200_ms"], 2)
            if equal is not None:
                r["bit_exact"] = bool(equal())
        rows.append(r)
        print(json.dumps(r), flush=True)

    levels = [(40000, 2048, 0.2, 64), (2048, 1024, 0.4, 32), (1024, 512, 0.8, 16), (512, 256, 1.2, 16)]
    cur = xyz
    for n, m, r, ns in levels:
        x = cur
        add("fps %d->%d" % (n, m), lambda: ext.furthest_point_sampling(x, m),
            lambda: ref.furthest_point_sampling(x, m),
            lambda: torch.equal(ext.furthest_point_sampling(x, m), ref.furthest_point_sampling(x, m)))
        inds, centres = ext.furthest_point_sampling(x, m, return_xyz=True)
        add("ball_query n=%d m=%d r=%.1f ns=%d" % (n, m, r, ns), lambda: ext.ball_query(centres, x, r, ns),
            lambda: ref.ball_query(centres, x, r, ns),
            lambda: torch.equal(ext.ball_query(centres, x, r, ns), ref.ball_query(centres, x, r, ns)))
        if n == 40000:
            idx = ext.ball_query(centres, x, r, ns)
            add("group_points C=7", lambda: ext.group_points(feats, idx), lambda: ref.group_points(feats, idx),
                lambda: torch.equal(ext.group_points(feats, idx), ref.group_points(feats, idx)))
            xt = x.transpose(1, 2).contiguous()
            add("gather_points C=3", lambda: ext.gather_points(xt, inds), lambda: ref.gather_points(xt, inds),
                lambda: torch.equal(ext.gather_points(xt, inds), ref.gather_points(xt, inds)))
        cur = centres
    u, k = torch.rand(B, 1024, 3, device="cuda"), torch.rand(B, 512, 3, device="cuda")
    add("three_nn 1024<-512", lambda: ext.three_nn(u, k), lambda: ref.three_nn(u, k),
        lambda: torch.equal(ext.three_nn(u, k)[1], ref.three_nn(u, k)[1]))
    d2, i3 = ext.three_nn(u, k)
    w = torch.rand(B, 1024, 3, device="cuda")
    f = torch.randn(B, 256, 512, device="cuda")
    add("three_interpolate C=256", lambda: ext.three_interpolate(f, i3, w), lambda: ref.three_interpolate(f, i3, w),
        lambda: torch.equal(ext.three_interpolate(f, i3, w), ref.three_interpolate(f, i3, w)))
    print(json.dumps({"summary": rows, "ref_ext_available": ref is not None,
                      "gpu": torch.cuda.get_device_name(0)}))


if __name__ == "__main__" and "--pipeline" not in sys.argv:
    main()


def reference_pipeline_forward():
    """Whole-backbone 'kernel to beat': the same un-fused module structure the reference runs
    (FPS -> gather -> ball query -> group x2 -> cat -> cuDNN conv/BN/ReLU -> max_pool, FP via
    three_nn/three_interpolate + cuDNN), with the REFERENCE extension's kernels substituted for
    the nine ops and torch's default TF32 convs, vs this repo's default path."""
    import bridgeqa_b200
    from bridgeqa_b200 import detector, pointnet2_utils
    import bridgeqa_b200.ext as our_ext
    ref = ref_ext.load()
    pc = synthetic.make_batch(16, 40000, 7).cuda()
    net = synthetic.fill_state_dict(detector.Pointnet2Backbone(input_feature_dim=7), seed=0).cuda().eval()

    def fwd():
        with torch.no_grad():
            return net({"point_clouds": pc})["fp2_features"]

    out = {"ours_fused_ms": round(timeit(fwd, warm=3, it=10), 3)}
    net.enable_cuda_graph()
    out["ours_fused_cuda_graph_ms"] = round(timeit(fwd, warm=3, it=10), 3)
    net.enable_cuda_graph(False)
    bridgeqa_b200.set_fused(False)
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    out["ours_unfused_tf32_ms"] = round(timeit(fwd, warm=2, it=5), 3)
    if ref is not None:
        pointnet2_utils.set_group_concat(False)      # the reference's op sequence, on its kernels
        net.prefix_check = False
        saved = {}
        names = ["gather_points", "ball_query", "group_points", "three_nn", "three_interpolate"]
        for n in names:
            saved[n] = getattr(our_ext, n)
            setattr(our_ext, n, getattr(ref, n))
        saved["furthest_point_sampling"] = our_ext.furthest_point_sampling

        def ref_fps(points, nsamples, return_xyz=False):
            inds = ref.furthest_point_sampling(points, nsamples)
            if not return_xyz:
                return inds
            xyz = ref.gather_points(points.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()
            return inds, xyz
        our_ext.furthest_point_sampling = ref_fps
        try:
            out["reference_ext_pipeline_ms"] = round(timeit(fwd, warm=1, it=3), 3)
        finally:
            for n, f in saved.items():
                setattr(our_ext, n, f)
            pointnet2_utils.set_group_concat(True)
    # configs[3]: DET training step (C=132), this repo vs the reference's kernels + torch modules
    from bridgeqa_b200 import train_fused, training
    pc132 = synthetic.make_batch(16, 40000, 132).cuda()
    det = synthetic.fill_state_dict(detector.VoteNetDetector(132), seed=0).cuda()
    loss_fn = training.ProjectionLoss().cuda()
    out["ours_train_step_ms"] = round(timeit(lambda: training.train_step(det, loss_fn, pc132), warm=2, it=5), 3)
    if ref is not None:
        train_fused.set_enabled(False)
        pointnet2_utils.set_group_concat(False)
        det.detection_backbone.prefix_check = False
        for n in names:
            setattr(our_ext, n, getattr(ref, n))
        our_ext.furthest_point_sampling = ref_fps
        try:
            out["reference_ext_train_step_ms"] = round(
                timeit(lambda: training.train_step(det, loss_fn, pc132), warm=1, it=3), 3)
        finally:
            for n, f in saved.items():
                setattr(our_ext, n, f)
            train_fused.set_enabled(True)
            pointnet2_utils.set_group_concat(True)
    bridgeqa_b200.set_fused(True)
    out["speedup_vs_reference_ext_pipeline"] = (round(out["reference_ext_pipeline_ms"] / out["ours_fused_cuda_graph_ms"], 1)
                                                if "reference_ext_pipeline_ms" in out else None)
    out["train_speedup_vs_reference_ext"] = (round(out["reference_ext_train_step_ms"] / out["ours_train_step_ms"], 1)
                                             if "reference_ext_train_step_ms" in out else None)
    print(json.dumps({"backbone_forward_B16_N40000_C7": out}))


if __name__ == "__main__" and "--pipeline" in sys.argv:
    reference_pipeline_forward()
