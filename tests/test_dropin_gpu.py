"""The literal drop-in, run on the GPU: the UNMODIFIED reference Python layer
(lib/pointnet2/pointnet2_modules.py, pointnet2_utils.py, pytorch_utils.py, models/backbone_module.py,
models/voting_module.py -- staged byte-for-byte into the git-ignored oracle/_ref/ref_tree by
oracle/build_ref.py) executes a forward

  (a) on the reference's own CUDA extension (oracle/_ref/pointnet2_ref/_ext.so), and
  (b) on bridgeqa_b200.ext -- the object compat.install("ext") registers as `pointnet2._ext` --

with the same weights and inputs.  Every op of this repo is bit-exact against the reference's, the
Python layer and torch's convs are the same code, so every tensor of the data_dict must be
BIT-EQUAL.  (c): the same reference model files over compat.install("modules") (this repo's module
mirrors with the fused tcgen05 kernels) stay within the fused tolerance.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from bridgeqa_b200 import ext as b200_ext, synthetic  # noqa: E402
from oracle import build_ref, ref_ext as ref_loader  # noqa: E402


def _need_ref():
    ref = ref_loader.load()
    if ref is None or not build_ref.tree_available():
        pytest.skip("oracle/_ref (reference extension + staged Python layer) not built")
    return ref


def test_unmodified_reference_backbone_and_voting_are_bit_equal_on_both_extensions():
    ref = _need_ref()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    bm, vm, utils = ref_loader.load_reference_modules(ref)
    assert utils._ext is ref and "oracle/_ref/ref_tree" in bm.__file__
    C = 7
    net = synthetic.fill_state_dict(bm.Pointnet2Backbone(input_feature_dim=C), seed=0).cuda().eval()
    vote = synthetic.fill_state_dict(vm.VotingModule(1, 256), seed=8).cuda().eval()
    pc = synthetic.make_batch(4, 40000, C, first_scene=0).cuda()

    def forward():
        with torch.no_grad():
            dd = net({"point_clouds": pc})
            vxyz, vfeat = vote(dd["fp2_xyz"], dd["fp2_features"])
        out = {k: v.clone() for k, v in dd.items() if torch.is_tensor(v)}
        out["vote_xyz"], out["vote_features"] = vxyz.clone(), vfeat.clone()
        # the grouped tensor of SA1 through the reference's QueryAndGroup (pointnet2_utils.py:317-376)
        xyz = pc[..., :3].contiguous()
        feats = pc[..., 3:].transpose(1, 2).contiguous()
        with torch.no_grad():
            grouped, grouped_xyz = net.sa1.grouper(xyz, dd["sa1_xyz"], feats)      # ret_grouped_xyz=True
            out["sa1_grouped"], out["sa1_grouped_xyz"] = grouped.clone(), grouped_xyz.clone()
        return out

    a = forward()                                   # (a) reference extension
    assert a["sa1_inds"].dtype == torch.int32
    utils._ext = b200_ext                           # (b) exactly what compat.install("ext") provides
    try:
        from bridgeqa_b200 import _native
        before = _native.launch_count()
        b = forward()
        assert _native.launch_count() - before >= 20          # this repo's kernels really ran
    finally:
        utils._ext = ref
    assert set(a) == set(b)
    for k in sorted(a):
        assert torch.equal(a[k], b[k]), "reference modules: %s differs between the two extensions" % k


SCRIPT = r'''
import os, sys
import numpy as np, torch
sys.path.insert(0, %(root)r)
from bridgeqa_b200 import compat, synthetic
compat.install(level="modules")
tree = os.path.join(%(root)r, "oracle", "_ref", "ref_tree")
os.chdir(tree); sys.path.insert(0, tree)
from models.backbone_module import Pointnet2Backbone          # the reference's file, unmodified
import lib.pointnet2.pointnet2_modules as m
assert "bridgeqa_b200" in m.__file__, m.__file__
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
net = synthetic.fill_state_dict(Pointnet2Backbone(input_feature_dim=7), seed=0).cuda().eval()
pc = synthetic.make_batch(2, 40000, 7, first_scene=0)
with torch.no_grad():
    dd = net({"point_clouds": pc.cuda()})
sys.path.insert(0, %(root)r)
from oracle import modules_cpu
want = modules_cpu.backbone(pc.numpy(), {k: v.cpu() for k, v in net.state_dict().items()})
for k in ("sa1_inds", "sa2_inds", "fp2_inds", "sa4_xyz"):
    assert np.array_equal(dd[k].cpu().numpy(), want[k]), k
err = np.abs(dd["fp2_features"].cpu().numpy() - want["fp2_features"]).max() / np.abs(want["fp2_features"]).max()
assert err < 2e-3, err
print("DROPIN_MODULES_OK %%.2e" %% err)
'''


def test_unmodified_reference_backbone_over_compat_modules_matches_oracle():
    if not build_ref.tree_available():
        pytest.skip("oracle/_ref/ref_tree not staged")
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "DROPIN_MODULES_OK" in r.stdout, r.stdout[-3000:]
