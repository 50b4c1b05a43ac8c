"""Developer tool: time FPS alone at several shapes and check against the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bridgeqa_b200 import ext, synthetic
from oracle import cpu_ops

def t(fn, it=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(it):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]

for (b, n, m, check) in [(16, 40000, 2048, 2), (8, 40000, 2048, 1), (15, 40000, 512, 1), (2, 20000, 512, 2), (16, 2048, 1024, 0), (64, 40000, 256, 0), (3, 9000, 300, 3)]:
    xyz = synthetic.make_batch(b, n, 0, first_scene=7)[..., :3].contiguous()
    x = xyz.cuda()
    ms = t(lambda: ext.furthest_point_sampling(x, m))
    ok = ""
    if check:
        got = ext.furthest_point_sampling(x, m).cpu().numpy()[:check]
        want = cpu_ops.furthest_point_sampling(xyz.numpy()[:check], m)
        ok = "exact" if np.array_equal(got, want) else "MISMATCH"
        if b % 2 == 1:   # the odd last scene
            got_l = ext.furthest_point_sampling(x, m).cpu().numpy()[-1:]
            want_l = cpu_ops.furthest_point_sampling(xyz.numpy()[-1:], m)
            ok += " last=" + ("exact" if np.array_equal(got_l, want_l) else "MISMATCH")
    print("fps b=%d n=%d m=%d: %.3f ms  %.3f us/iter  %s" % (b, n, m, ms, 1e3 * ms / max(m - 1, 1), ok), flush=True)
