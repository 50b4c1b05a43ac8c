"""Start/end offsets (ms, relative to the step's first launch) of every C-ABI call of one
backbone forward, across the main and side streams -- a poor man's nsys timeline from the
CUDA events KernelTimer already records.  Measurement tool only.

    gpurun -- 'python tools/timeline.py [detector]'
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bridgeqa_b200 import detector, profiler, synthetic  # noqa: E402


def main():
    det = "detector" in sys.argv
    c = 132 if det else 7
    pcs = [synthetic.make_batch(16, 40000, c, first_scene=16 * i).cuda() for i in range(3)]
    net = detector.VoteNetDetector(c) if det else detector.Pointnet2Backbone(input_feature_dim=c)
    net = synthetic.fill_state_dict(net, seed=0).cuda().eval()
    with torch.no_grad():
        for i in range(4):
            net({"point_clouds": pcs[i % 3]})
        torch.cuda.synchronize()
        with profiler.KernelTimer() as kt:
            base = torch.cuda.Event(enable_timing=True)
            base.record()
            for i in range(3):
                net({"point_clouds": pcs[i % 3]})
            last = torch.cuda.Event(enable_timing=True)
            last.record()
        torch.cuda.synchronize()
    print("3 steps: %.3f ms" % base.elapsed_time(last))
    for name, ints, start, end in kt.records:
        print("%8.3f -> %8.3f  (%6.3f)  %s %s" % (base.elapsed_time(start), base.elapsed_time(end),
                                                 start.elapsed_time(end), name, list(ints)[:5]))


if __name__ == "__main__":
    main()
