"""Developer tool: two DET training steps (C=132, B=16) for ncu captures of the conv / BatchNorm kernels."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bridgeqa_b200 import detector, synthetic, training
torch.backends.cudnn.allow_tf32 = True
pc = synthetic.make_batch(16, 40000, 132).cuda()
net = synthetic.fill_state_dict(detector.VoteNetDetector(132), seed=0).cuda()
loss_fn = training.ProjectionLoss().cuda()
for _ in range(2):
    training.train_step(net, loss_fn, pc)
torch.cuda.synchronize()
print("ok")
