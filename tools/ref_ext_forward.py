"""Times the STOCK reference (unmodified Python layer staged in oracle/_ref/ref_tree + the reference's
own CUDA extension oracle/_ref/pointnet2_ref/_ext.so, torch's default TF32 convs) on the bench's
headline workload, on this GPU.  Prints one JSON object.  Run by bench.py in a subprocess (its
`ref_ext` field) so that none of this is loaded into the product process.

    python tools/ref_ext_forward.py [--steps 5] [--warmup 2] [--features 7] [--batch 16] [--points 40000]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--features", type=int, default=7)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--points", type=int, default=40000)
    args = ap.parse_args()
    import torch
    from oracle import build_ref, ref_ext
    ref = ref_ext.load()
    if ref is None or not build_ref.tree_available():
        print(json.dumps({"unavailable": "oracle/_ref (reference extension + staged Python layer) not built"}))
        return 0
    from bridgeqa_b200 import synthetic            # input + weight generators only
    bm, _, utils = ref_ext.load_reference_modules(ref)
    assert utils._ext is ref
    torch.backends.cudnn.allow_tf32 = True          # torch's default; the reference never changes it
    torch.backends.cuda.matmul.allow_tf32 = True
    net = synthetic.fill_state_dict(bm.Pointnet2Backbone(input_feature_dim=args.features), seed=0).cuda().eval()
    pc = synthetic.make_batch(args.batch, args.points, args.features, first_scene=0).cuda()
    with torch.no_grad():
        for _ in range(args.warmup):
            net({"point_clouds": pc})
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            out = net({"point_clouds": pc})
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({
        "value": args.batch / (ms / 1e3), "unit": "scenes/s", "ms_per_step": ms, "steps": args.steps,
        "warmup": args.warmup,
        "what": "UNMODIFIED reference models/backbone_module.py + lib/pointnet2/*.py on the reference's own CUDA "
                "extension (sm_100a build of _ext_src, oracle/_ref) + cuDNN TF32 convs, same GPU, same synthetic "
                "batch (%d x %d points, C=%d), device-resident input, eager (the reference has no graph path)"
                % (args.batch, args.points, args.features),
        "checksum_fp2_features": float(out["fp2_features"].double().abs().mean())}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
