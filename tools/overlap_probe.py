"""Developer tool: throughput of the graphed backbone forward with 1..6 batches in flight
(graphs.InFlight), with the latency and the throughput variant of the sorted sampling kernel.
Prints ms/step per (variant, depth) and checks the outputs against one-at-a-time forwards.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bridgeqa_b200 import detector, synthetic

STEPS = int(os.environ.get("STEPS", "60"))
DEPTHS = [int(x) for x in os.environ.get("DEPTHS", "1,2,3,4,6").split(",")]
ROT = 6
dev = torch.device("cuda:0")
net = synthetic.fill_state_dict(detector.Pointnet2Backbone(input_feature_dim=7), seed=0).to(dev).eval()
host = synthetic.make_batch(16, 40000, 7)
inputs = [torch.roll(host, shifts=997 * i, dims=1).contiguous().to(dev) for i in range(ROT)]
KEYS = ("fp2_features", "fp2_inds", "sa1_inds", "sa4_features")
with torch.no_grad():
    ref = [{k: net({"point_clouds": p})[k].clone() for k in KEYS} for p in inputs]
net.enable_cuda_graph(bind_inputs=True)

res = {}
for lean in (False, True):
    for depth in DEPTHS:
        q = net.in_flight(depth, lean_sampling=lean)
        main = torch.cuda.current_stream(dev)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for rep in range(2):              # first repetition is the warm-up (and the capture)
            torch.cuda.synchronize()
            ev0.record(main)
            tickets = [q.submit({"point_clouds": inputs[i % ROT]}) for i in range(STEPS)]
            q.drain()
            ev1.record(main)
            torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / STEPS
        same = all(torch.equal(tickets[-1 - j].out[k], ref[(STEPS - 1 - j) % ROT][k]) for j in range(ROT) for k in KEYS)
        res["%s_depth%d" % ("lean" if lean else "latency", depth)] = {
            "ms_per_step": round(ms, 4), "scenes_per_s": round(16e3 / ms, 1), "bit_identical": bool(same)}
print(json.dumps(res))
