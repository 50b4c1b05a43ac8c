"""CUPTI (torch profiler) durations of the kernels of five eager backbone forwards, 16 x 40000 points, C = 7.
    python tools/time_sa_fp.py [substring of a kernel name ...]     (default: sa_v2 fp_mlp)
"""
import sys, torch
sys.path.insert(0, ".")
import os
from bridgeqa_b200 import detector, synthetic, _native as N
if os.environ.get("BQA_SO"):
    N.SO_PATH = os.environ["BQA_SO"]            # A/B against another build of the library
C = int(os.environ.get("BQA_C", "7"))
pc = synthetic.make_batch(16, 40000, C).cuda()
net = synthetic.fill_state_dict(detector.Pointnet2Backbone(input_feature_dim=C), seed=0).cuda().eval()
names = ["bqa_sa_mlp_max_forward_v2", "bqa_fp_mlp_forward"]
with torch.no_grad():
    for _ in range(3): net({"point_clouds": pc})
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5): net({"point_clouds": pc})
        torch.cuda.synchronize()
    for e in prof.key_averages():
        if any(k in e.key for k in (sys.argv[1:] or ["sa_v2", "fp_mlp"])):
            print(e.key[:60], e.count, round(e.device_time_total / e.count, 1), "us")
