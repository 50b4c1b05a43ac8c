"""Developer tool: one small forward through every hand-written kernel family, sized so that the
sampling kernels run as multi-CTA clusters (DSMEM + mbarrier path) -- the payload of tools/sanitize.sh."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bridgeqa_b200 import detector, ext, fused, synthetic

B, N, C = 2, 30000, 7
pc = synthetic.make_batch(B, N, C).cuda()
xyz = pc[..., :3].contiguous()
net = synthetic.fill_state_dict(detector.Pointnet2Backbone(input_feature_dim=C), seed=0).cuda().eval()
with torch.no_grad():
    inds = ext.furthest_point_sampling(xyz, 256)              # plain cluster kernel (fps.cu)
    for lean in (False, True):                                 # sorted kernels, latency + throughput variants
        with fused.lean_sampling(lean):
            out = net({"point_clouds": pc})
torch.cuda.synchronize()
print("ok", tuple(out["fp2_features"].shape), int(inds.sum()))
