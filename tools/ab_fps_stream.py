import os, sys, torch
sys.path.insert(0, ".")
from bridgeqa_b200 import _native as N
if os.environ.get("BQA_SO"):
    N.SO_PATH = os.environ["BQA_SO"]
from bridgeqa_b200 import ext, fused, synthetic
xyz = synthetic.make_batch(16, 40000, 0)[..., :3].contiguous().cuda()
grid = fused.prebuild_ball_query_grid(xyz, 0.2, inline=True)
ref = ext.furthest_point_sampling(xyz, 2048)
for _ in range(3): out = fused.furthest_point_sample_grid(xyz, 2048, grid, lean=True)
torch.cuda.synchronize()
ts = []
for _ in range(15):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = fused.furthest_point_sample_grid(xyz, 2048, grid, lean=True); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print(os.environ.get("BQA_SO", "tree"), "stream fps ms", sorted(ts)[7], "same", torch.equal(ref, out[0]))
