"""FPS at SA1 size: plain register-resident kernel vs sorted/pruned kernel (+ its grid build).
    gpurun -- 'for c in 0 4 6 8; do BQA_FPS_SORTED_CS=$c python tools/time_fps_grid.py; done'
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bridgeqa_b200 import ext, fused, synthetic  # noqa: E402


def timeit(fn, warm=3, it=15):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(it):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


print("BQA_FPS_SORTED_CS =", os.environ.get("BQA_FPS_SORTED_CS"))
for b, n, m in [(16, 40000, 2048), (8, 40000, 2048), (16, 20000, 2048), (16, 100000, 2048), (64, 40000, 2048)]:
    xyz = synthetic.make_batch(b, n, 0)[..., :3].contiguous().cuda()
    R = float(os.environ.get("GRID_R", "0.2")); grid = fused.prebuild_ball_query_grid(xyz, R, inline=True)
    t_plain = timeit(lambda: ext.furthest_point_sampling(xyz, m, return_xyz=True))
    t_build = timeit(lambda: fused.prebuild_ball_query_grid(xyz, 0.2, inline=True))
    t_sorted = timeit(lambda: fused.furthest_point_sample_grid(xyz, m, grid))
    same = torch.equal(ext.furthest_point_sampling(xyz, m), fused.furthest_point_sample_grid(xyz, m, grid)[0])
    print("b=%d n=%d m=%d: plain %.3f ms | grid build %.3f + sorted %.3f ms (%.3f us/iter)  same=%s"
          % (b, n, m, t_plain, t_build, t_sorted, 1e3 * t_sorted / (m - 1), same), flush=True)
