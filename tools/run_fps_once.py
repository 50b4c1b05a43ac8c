"""Developer tool: one throughput-variant sampling call at SA1 size (for ncu captures)."""
import sys, torch
sys.path.insert(0, ".")
from bridgeqa_b200 import fused, synthetic
b, n = 16, 40000
xyz = synthetic.make_batch(b, n, 0)[..., :3].contiguous().cuda()
grid = fused.prebuild_ball_query_grid(xyz, 0.2, inline=True)
for _ in range(2):
    fused.furthest_point_sample_grid(xyz, 2048, grid, lean=True)
torch.cuda.synchronize()
