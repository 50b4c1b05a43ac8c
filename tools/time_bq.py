"""ball query latency per SA level, cell grid (build + search) vs index-order scan.
    gpurun -- 'for g in 1 1000000000; do BQA_BQ_GRID_MIN=$g python tools/time_bq.py; done'
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bridgeqa_b200 import ext, synthetic  # noqa: E402


def timeit(fn, warm=5, it=30):
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


cur = synthetic.make_batch(16, 40000, 0)[..., :3].contiguous().cuda()
print("BQA_BQ_GRID_MIN =", os.environ.get("BQA_BQ_GRID_MIN"))
for m, r, ns in [(2048, 0.2, 64), (1024, 0.4, 32), (512, 0.8, 16), (256, 1.2, 16)]:
    _, c = ext.furthest_point_sampling(cur, m, return_xyz=True)
    x = cur
    print("n=%6d m=%5d r=%.1f ns=%3d  %.4f ms" % (x.size(1), m, r, ns, timeit(lambda: ext.ball_query(c, x, r, ns))))
    cur = c
