"""Generate tests/golden/ref_python_layer.npz by running the UNMODIFIED reference Python
layer (/root/reference/lib/pointnet2/{pointnet2_utils,pointnet2_modules,pytorch_utils}.py,
models/{backbone,voting}_module.py) on the CPU.

Only runnable in the build container (needs /root/reference).  The reference ops are
CUDA-only, so `pointnet2._ext` is provided by the C oracle (oracle/cpu_ops.as_ext_module);
what this pins is therefore the *composition* -- QueryAndGroup, SharedMLP, max-pool, the
FP weights, the backbone wiring, the voting module -- i.e. oracle/modules_cpu.py and the
product's module classes, against the reference's own Python.  The ops themselves are
pinned against the reference CUDA extension by make_golden_gpu.py.

    python tests/golden/make_golden_cpu.py
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import cpu_ops  # noqa: E402
from bridgeqa_b200 import synthetic  # noqa: E402  (input + weight generators only)


def sample(a, k=4096):
    flat = np.ascontiguousarray(a).reshape(-1)
    stride = max(1, flat.size // k)
    return flat[::stride][:k].copy()


def main():
    pkg = types.ModuleType("pointnet2")
    pkg._ext = cpu_ops.as_ext_module()
    sys.modules["pointnet2"] = pkg
    sys.modules["pointnet2._ext"] = pkg._ext
    os.chdir(REF)                       # backbone_module.py:8 appends os.getcwd()/lib
    sys.path.insert(0, REF)
    from models.backbone_module import Pointnet2Backbone
    from models.voting_module import VotingModule

    torch.manual_seed(0)
    B, N, C = 2, 4096, 4
    pc = synthetic.make_batch(B, N, C, first_scene=50)
    net = Pointnet2Backbone(input_feature_dim=C)
    synthetic.fill_state_dict(net, seed=7)
    net.eval()
    vote = VotingModule(1, 256)
    synthetic.fill_state_dict(vote, seed=8)
    vote.eval()
    with torch.no_grad():
        dd = net({"point_clouds": pc})
        vxyz, vfeat = vote(dd["fp2_xyz"], dd["fp2_features"])

    out = {"meta_B": B, "meta_N": N, "meta_C": C, "meta_first_scene": 50,
           "meta_backbone_seed": 7, "meta_voting_seed": 8}
    for k in ("sa1_inds", "sa2_inds", "fp2_inds"):
        out[k] = dd[k].numpy().astype(np.int32)
    for k in ("sa1_xyz", "sa2_xyz", "sa3_xyz", "sa4_xyz"):
        out[k] = dd[k].numpy()
    for k in ("sa1_features", "sa2_features", "sa3_features", "sa4_features", "fp2_features"):
        a = dd[k].numpy()
        out[k + "_sample"] = sample(a)
        out[k + "_sum"] = np.float64(a.astype(np.float64).sum())
        out[k + "_abssum"] = np.float64(np.abs(a.astype(np.float64)).sum())
    out["vote_xyz"] = vxyz.numpy()
    out["vote_features_sample"] = sample(vfeat.numpy())
    out["vote_features_abssum"] = np.float64(np.abs(vfeat.numpy().astype(np.float64)).sum())
    # state_dict key list: the checkpoint-compatibility contract
    out["backbone_keys"] = np.array(sorted(net.state_dict().keys()))
    out["voting_keys"] = np.array(sorted(vote.state_dict().keys()))
    path = os.path.join(ROOT, "tests", "golden", "ref_python_layer.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
