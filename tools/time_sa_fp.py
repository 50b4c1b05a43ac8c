import sys, torch
sys.path.insert(0, ".")
from bridgeqa_b200 import detector, synthetic, _native as N
pc = synthetic.make_batch(16, 40000, 7).cuda()
net = synthetic.fill_state_dict(detector.Pointnet2Backbone(input_feature_dim=7), seed=0).cuda().eval()
names = ["bqa_sa_mlp_max_forward_v2", "bqa_fp_mlp_forward"]
with torch.no_grad():
    for _ in range(3): net({"point_clouds": pc})
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5): net({"point_clouds": pc})
        torch.cuda.synchronize()
    for e in prof.key_averages():
        if "sa_v2" in e.key or "fp_mlp" in e.key:
            print(e.key[:60], e.count, round(e.device_time_total / e.count, 1), "us")
