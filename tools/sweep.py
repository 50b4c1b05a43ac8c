"""BASELINE.json configs[4]: FPS and ball-query latency vs points per scene and batch size
(N in {20k, 40k, 100k} x B in {8, 16, 64}), on one GPU or with the B scenes sharded over the ranks of a
torchrun launch (scenes are independent: no collective; a row's time is the max over ranks).

    gpurun -- 'python tools/sweep.py > gpurun_out/r2_sweep_1gpu.json'
    gpurun --gpus 8 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
                        --master-port 29533 tools/sweep.py > gpurun_out/r2_sweep_8gpu.json'
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from bridgeqa_b200 import distributed as D, ext, fused, synthetic


def med(fn, it=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(it):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rows = []
    base = synthetic.make_batch(8, 100000, 0, first_scene=200)[..., :3].contiguous()
    m = 2048
    for n in (20000, 40000, 100000):
        for b in (8, 16, 64):
            lo, hi = D.shard_scenes(b, rank, world)
            mine = hi - lo                                   # scenes of this rank (b >= world here)
            xyz = base[:, :n].repeat((b + 7) // 8, 1, 1)[lo:hi].contiguous().to(dev)
            fps = med(lambda: ext.furthest_point_sampling(xyz, m))

            def fps_grid(lean):
                g = fused.prebuild_ball_query_grid(xyz, 0.2, inline=True)
                return fused.furthest_point_sample_grid(xyz, m, g, lean=lean)
            ok = fused.fps_grid_supported(n, m)
            fps_g = med(lambda: fps_grid(False)) if ok else None
            fps_t = med(lambda: fps_grid(True)) if ok else None
            inds, centres = ext.furthest_point_sampling(xyz, m, return_xyz=True)
            bq = med(lambda: ext.ball_query(centres, xyz, 0.2, 64))
            vals = [fps, fps_g or 0.0, fps_t or 0.0, bq]
            vals = [D.max_over_ranks(v, dev) for v in vals]   # a row is as slow as its slowest rank
            fps, fps_g, fps_t, bq = vals
            rows.append({"n": n, "b": b, "scenes_per_gpu": mine, "npoint": m,
                         "fps_ms": round(fps, 3), "fps_us_per_iter": round(1e3 * fps / (m - 1), 3),
                         "fps_scenes_per_s": round(b / fps * 1e3, 1),
                         "fps_sorted_incl_grid_build_ms": round(fps_g, 3) if ok else None,
                         "fps_throughput_variant_incl_grid_build_ms": round(fps_t, 3) if ok else None,
                         "ball_query_ms": round(bq, 3),
                         "ball_query_gpairs_per_s": round(b * n * m / bq / 1e6, 1)})
            if rank == 0:
                print(json.dumps(rows[-1]), file=sys.stderr, flush=True)
            del xyz
    if rank == 0:
        print(json.dumps({"gpu": torch.cuda.get_device_name(local), "n_gpus": world,
                          "sharding": "the B scenes of a row are split over the ranks (no collective); times are "
                                      "the max over ranks", "sweep": rows}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
