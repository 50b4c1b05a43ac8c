"""Per-phase time stamps of the fused SA kernels (needs tools/bin/libbqa_stats.so: bash tools/build_stats.sh)."""
import sys, torch
sys.path.insert(0, ".")
import bridgeqa_b200._native as N
N.SO_PATH = "tools/bin/libbqa_stats.so"
from bridgeqa_b200 import detector, synthetic
cs = [int(a) for a in sys.argv[1:]] or [7]
for c in cs:
    pc = synthetic.make_batch(16, 40000, c).cuda()
    net = synthetic.fill_state_dict(detector.Pointnet2Backbone(input_feature_dim=c), seed=0).cuda().eval()
    with torch.no_grad():
        net({"point_clouds": pc}); torch.cuda.synchronize()
        sys.stderr.write("---- C = %d\n" % c); sys.stderr.flush()
        net({"point_clouds": pc}); torch.cuda.synchronize()
