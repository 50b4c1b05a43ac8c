"""Generate tests/golden/ref_ext_*.npz: outputs of the reference's OWN CUDA extension
(oracle/_ref, built unmodified from /root/reference/lib/pointnet2/_ext_src by
oracle/build_ref.py) on seeded synthetic scenes.  Needs a GPU:

    gpurun -- 'python tests/golden/make_golden_gpu.py gpurun_out/golden'

then copy gpurun_out/golden/ref_ext_*.npz into tests/golden/ and commit.  The CPU-only
suite (tests/test_oracle_cpu.py) replays them against oracle/pointnet2_oracle.c; inputs
are regenerated from the seeds stored in each file (bridgeqa_b200.synthetic is
deterministic on the CPU).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from bridgeqa_b200 import synthetic  # noqa: E402
from oracle import ref_ext  # noqa: E402

# (name, B, N, first_scene, npoint, radius, nsample, nn_m)
CASES = [
    ("sa1_small", 2, 4096, 70, 256, 0.2, 64, 128),
    ("sa2_like", 2, 2048, 71, 1024, 0.4, 32, 512),
    ("tiny_lt512", 3, 300, 72, 64, 0.8, 16, 5),
    ("ragged", 2, 5001, 73, 77, 0.3, 16, 40),
]


def main(outdir):
    ext = ref_ext.load()
    if ext is None:
        raise SystemExit("oracle/_ref is not built")
    os.makedirs(outdir, exist_ok=True)
    for name, B, N, first, npoint, radius, nsample, m in CASES:
        xyz = synthetic.make_batch(B, N, 0, first_scene=first)[..., :3].contiguous().cuda()
        inds = ext.furthest_point_sampling(xyz, npoint)
        centres = ext.gather_points(xyz.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()
        bq = ext.ball_query(centres, xyz, radius, nsample)
        known = centres[:, :m].contiguous()
        d2, nn_idx = ext.three_nn(centres, known)
        w = 1.0 / (torch.sqrt(d2) + 1e-8)
        w = (w / w.sum(-1, keepdim=True)).contiguous()
        feats = xyz[:, :m].transpose(1, 2).contiguous()
        interp = ext.three_interpolate(feats, nn_idx, w)
        grouped = ext.group_points(xyz.transpose(1, 2).contiguous(), bq)
        torch.cuda.synchronize()
        path = os.path.join(outdir, "ref_ext_%s.npz" % name)
        np.savez_compressed(
            path, B=B, N=N, first_scene=first, npoint=npoint, radius=radius, nsample=nsample, nn_m=m,
            fps_inds=inds.cpu().numpy(), ball_idx=bq.cpu().numpy(), nn_idx=nn_idx.cpu().numpy(),
            nn_dist2=d2.cpu().numpy(), nn_weight=w.cpu().numpy(), interp=interp.cpu().numpy(),
            grouped_xyz_strided=grouped[:, :, ::16].contiguous().cpu().numpy())
        print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
