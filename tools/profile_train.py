"""Where does the configs[3] training step spend its time?  torch.profiler kernel table of
one fwd+bwd of the detector (C=132, B=16, train-mode BN).  Measurement tool only.

    gpurun -- 'python tools/profile_train.py > gpurun_out/profile_train.txt'
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bridgeqa_b200 import detector, synthetic, training  # noqa: E402


def main():
    c = int(os.environ.get("C", "132"))
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    pc = synthetic.make_batch(16, 40000, c).cuda()
    net = synthetic.fill_state_dict(detector.VoteNetDetector(input_feature_dim=c), seed=0).cuda()
    loss_fn = training.ProjectionLoss().cuda()
    for _ in range(3):
        training.train_step(net, loss_fn, pc)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        training.train_step(net, loss_fn, pc)
    b.record()
    torch.cuda.synchronize()
    print("train step ms:", a.elapsed_time(b) / 5)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(2):
            training.train_step(net, loss_fn, pc)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90))


if __name__ == "__main__":
    main()
