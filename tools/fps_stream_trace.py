"""Phase trace of the one-SM sampling kernel (needs tools/bin/libbqa_stats.so: bash tools/build_stats.sh)."""
import sys, torch
sys.path.insert(0, ".")
import bridgeqa_b200._native as N
N.SO_PATH = "tools/bin/libbqa_stats.so"
from bridgeqa_b200 import fused, synthetic
for b, n in [(16, 40000), (16, 20000)]:
    xyz = synthetic.make_batch(b, n, 0)[..., :3].contiguous().cuda()
    grid = fused.prebuild_ball_query_grid(xyz, 0.2, inline=True)
    for _ in range(2):
        fused.furthest_point_sample_grid(xyz, 2048, grid, lean=True)
        torch.cuda.synchronize()
