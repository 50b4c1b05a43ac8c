import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bridgeqa_b200 import ext, synthetic
b, n, m = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
x = synthetic.make_batch(b, n, 0, first_scene=7)[..., :3].contiguous().cuda()
ext.furthest_point_sampling(x, m); torch.cuda.synchronize()
ts = []
for _ in range(5):
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ext.furthest_point_sampling(x, m); e.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(e))
print("b=%d n=%d m=%d  %.3f ms" % (b, n, m, sorted(ts)[2]))
