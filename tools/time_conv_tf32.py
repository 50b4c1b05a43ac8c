"""tcgen05 TF32 1x1 conv kernels (csrc/conv_tf32.cu) vs cuDNN TF32 at the SA-layer shapes of the DET
training step: forward (+stats), dgrad, wgrad -- ms and GB/s of algorithmic bytes.  Measurement tool."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bridgeqa_b200 import train_fused  # noqa: E402

torch.backends.cudnn.allow_tf32 = True
torch.backends.cuda.matmul.allow_tf32 = True


def timeit(fn, warm=3, it=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(it):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


SHAPES = [(16, 135, 64, 2048, 64), (16, 64, 64, 2048, 64), (16, 64, 128, 2048, 64), (16, 131, 128, 1024, 32),
          (16, 128, 128, 1024, 32), (16, 128, 256, 1024, 32), (16, 259, 128, 512, 16), (16, 128, 256, 512, 16),
          (16, 512, 256, 1024, 1)]
for (B, cin, cout, np_, ns) in SHAPES:
    x = torch.randn(B, cin, np_, ns, device="cuda")
    w = torch.randn(cout, cin, device="cuda")
    g = torch.randn(B, cout, np_, ns, device="cuda")
    sums = torch.zeros(2 * cout, dtype=torch.float64, device="cuda")
    gb = 4e-9 * B * np_ * ns * (cin + cout)
    t_f = timeit(lambda: train_fused._conv_forward(x, w, None, sums))
    t_d = timeit(lambda: train_fused._conv_forward(g, w.t()))
    t_w = timeit(lambda: train_fused._conv_wgrad(x, g, cout, cin))
    w4 = w.view(cout, cin, 1, 1)
    t_cf = timeit(lambda: torch.nn.functional.conv2d(x, w4))
    xr, wr = x.clone().requires_grad_(True), w4.clone().requires_grad_(True)

    def cudnn_fb():
        xr.grad = wr.grad = None
        torch.nn.functional.conv2d(xr, wr).backward(g)
    t_cfb = timeit(cudnn_fb)
    print("B=%d %3d->%3d L=%7d (%.2f GB): fwd+stats %.3f ms %4.0f GB/s | dgrad %.3f ms %4.0f | wgrad %.3f ms %4.0f | "
          "cuDNN fwd %.3f ms %4.0f GB/s, fwd+bwd %.3f ms (ours %.3f)"
          % (B, cin, cout, np_ * ns, gb, t_f, gb / t_f * 1e3, t_d, gb / t_d * 1e3, t_w, gb / t_w * 1e3, t_cf,
             gb / t_cf * 1e3, t_cfb, t_f + t_d + t_w), flush=True)
