"""Developer tool: run each hot kernel once at the BASELINE sizes (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bridgeqa_b200 import detector, ext, synthetic

torch.manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
pc = synthetic.make_batch(B, 40000, 7).cuda()
net = synthetic.fill_state_dict(detector.Pointnet2Backbone(input_feature_dim=7), seed=0).cuda().eval()
from bridgeqa_b200 import fused
with torch.no_grad(), fused.lean_sampling(os.environ.get("LEAN", "0") == "1"):   # LEAN=1: throughput variant of the sampling
    for _ in range(2):
        out = net({"point_clouds": pc})
torch.cuda.synchronize()
print("ok", out["fp2_features"].shape)
