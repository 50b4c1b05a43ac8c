"""Developer tool: per-step device time and allocator state of the DET training step (C=132, B=16)."""
import gc
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bridgeqa_b200 import detector, synthetic, training

torch.backends.cudnn.allow_tf32 = True
torch.backends.cuda.matmul.allow_tf32 = True
pc = synthetic.make_batch(16, 40000, 132).cuda()
net = synthetic.fill_state_dict(detector.VoteNetDetector(132), seed=0).cuda()
loss_fn = training.ProjectionLoss().cuda()
GC = os.environ.get("GC", "default")
if GC == "off":
    gc.disable()
for _ in range(3):
    training.train_step(net, loss_fn, pc, next_point_clouds=pc)
torch.cuda.synchronize()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
stats = []
evs[0].record()
for i in range(20):
    training.train_step(net, loss_fn, pc, next_point_clouds=pc)
    if GC == "step":
        gc.collect()
    evs[i + 1].record()
    s = torch.cuda.memory_stats()
    stats.append((torch.cuda.memory_allocated() >> 20, torch.cuda.memory_reserved() >> 20,
                  s.get("num_alloc_retries", 0), s.get("num_device_alloc", 0), gc.get_count()))
torch.cuda.synchronize()
for i in range(20):
    print(i, "%.2f ms" % evs[i].elapsed_time(evs[i + 1]), "alloc %d MB reserved %d MB retries %d device_allocs %d gc %s" % stats[i])
print("GC=%s total %.2f ms/step; unreachable found by gc.collect(): %d" % (GC, evs[0].elapsed_time(evs[20]) / 20, gc.collect()))
