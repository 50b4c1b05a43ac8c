"""BASELINE.json configs[4]: FPS and ball-query latency vs points per scene and batch size
(N in {20k, 40k, 100k} x B in {8, 16, 64}), one GPU.  Prints one JSON object.

    gpurun -- 'python tools/sweep.py > gpurun_out/r1_sweep.json'
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bridgeqa_b200 import ext, fused, synthetic


def med(fn, it=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(it):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


rows = []
base = synthetic.make_batch(8, 100000, 0, first_scene=200)[..., :3].contiguous()
for n in (20000, 40000, 100000):
    for b in (8, 16, 64):
        xyz = base[:, :n].repeat((b + 7) // 8, 1, 1)[:b].contiguous().cuda()
        m = 2048
        fps = med(lambda: ext.furthest_point_sampling(xyz, m))
        def fps_grid():
            g = fused.prebuild_ball_query_grid(xyz, 0.2, inline=True)
            return fused.furthest_point_sample_grid(xyz, m, g)
        fps_g = med(fps_grid) if fused.fps_grid_supported(n, m) else None
        inds, centres = ext.furthest_point_sampling(xyz, m, return_xyz=True)
        bq = med(lambda: ext.ball_query(centres, xyz, 0.2, 64))
        rows.append({"n": n, "b": b, "npoint": m, "fps_ms": round(fps, 3), "fps_us_per_iter": round(1e3 * fps / (m - 1), 3),
                     "fps_scenes_per_s": round(b / fps * 1e3, 1),
                     "fps_sorted_incl_grid_build_ms": round(fps_g, 3) if fps_g else None, "ball_query_ms": round(bq, 3),
                     "ball_query_gpairs_per_s": round(b * n * m / bq / 1e6, 1)})
        print(json.dumps(rows[-1]), file=sys.stderr, flush=True)
        del xyz
print(json.dumps({"gpu": torch.cuda.get_device_name(0), "sweep": rows}))
