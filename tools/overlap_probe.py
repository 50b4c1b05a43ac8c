"""Developer tool: throughput of the graphed backbone forward with 1, 2 or 3 batches in flight.

The 40k-point sampling chain of a batch occupies 96 of the 148 SMs for ~60 % of its forward and
is pure latency; the next batch's chain can run underneath the current batch's SA/FP kernels when
consecutive steps are replayed on different streams.  Prints ms/step for each depth and checks
that the overlapped runs produce bit-identical outputs.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bridgeqa_b200 import detector, synthetic

STEPS = int(os.environ.get("STEPS", "60"))
ROT = 6
dev = torch.device("cuda:0")
net = synthetic.fill_state_dict(detector.Pointnet2Backbone(input_feature_dim=7), seed=0).to(dev).eval()
net.enable_cuda_graph(bind_inputs=True)
host = synthetic.make_batch(16, 40000, 7)
inputs = [torch.roll(host, shifts=997 * i, dims=1).contiguous().to(dev) for i in range(ROT)]
KEYS = ("fp2_features", "fp2_inds", "sa1_inds", "sa4_features")


def step(i):
    with torch.no_grad():
        return net({"point_clouds": inputs[i % ROT]})


ref = []
for i in range(ROT):
    dd = step(i)
    torch.cuda.synchronize()
    ref.append({k: dd[k].clone() for k in KEYS})

res = {}
for depth in (1, 2, 3, 1, 2):
    streams = [torch.cuda.Stream(dev) for _ in range(depth)]
    main = torch.cuda.current_stream(dev)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(2):              # first repetition is the warm-up
        ev0.record(main)
        for s in streams:
            s.wait_stream(main)
        outs = {}
        for i in range(STEPS):
            with torch.cuda.stream(streams[i % depth]):
                outs[i % ROT] = step(i)
        for s in streams:
            main.wait_stream(s)
        ev1.record(main)
        torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / STEPS
    same = all(torch.equal(outs[j][k], ref[j][k]) for j in outs for k in KEYS)
    res.setdefault("depth%d" % depth, []).append({"ms_per_step": round(ms, 4), "scenes_per_s": round(16e3 / ms, 1),
                                                  "bit_identical": bool(same)})
print(json.dumps(res))
