"""Sampling at SA1 size: cluster kernels (latency / throughput variants) vs the one-SM kernel (fps_stream.cu).
    gpurun -- 'python tools/time_fps_stream.py; BQA_FPS_STREAM_PPL=1 python tools/time_fps_stream.py'
Each variant is timed alone (kernel latency) and the SMs it holds are reported (SM-time = SMs x ms).
"""
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


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


def main():
    from bridgeqa_b200 import ext, fused, synthetic
    m = 2048
    mode = os.environ.get("BQA_FPS_STREAM", "1")
    for b, n in [(16, 40000), (8, 40000), (16, 20000), (64, 40000), (16, 50000)]:
        xyz = synthetic.make_batch(b, n, 0)[..., :3].contiguous().cuda()
        grid = fused.prebuild_ball_query_grid(xyz, 0.2, inline=True)
        ref = ext.furthest_point_sampling(xyz, m)
        out = {}
        for lean in (False, True):
            t = timeit(lambda: fused.furthest_point_sample_grid(xyz, m, grid, lean=lean))
            same = torch.equal(ref, fused.furthest_point_sample_grid(xyz, m, grid, lean=lean)[0])
            out[lean] = (t, same)
        print("BQA_FPS_STREAM=%s PPL=%s b=%d n=%d: latency variant %.3f ms (%.3f us/iter, same=%s) | throughput variant "
              "%.3f ms (%.3f us/iter, same=%s)" % (mode, os.environ.get("BQA_FPS_STREAM_PPL"), b, n, out[False][0],
                                                   1e3 * out[False][0] / (m - 1), out[False][1], out[True][0],
                                                   1e3 * out[True][0] / (m - 1), out[True][1]), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "all":
        for env in ({"BQA_FPS_STREAM": "0"}, {"BQA_FPS_STREAM": "1"}, {"BQA_FPS_STREAM": "1", "BQA_FPS_STREAM_PPL": "1"}):
            subprocess.run([sys.executable, __file__], env=dict(os.environ, **env))
    else:
        main()
