"""1x1 conv (cuDNN, NCHW) vs the same contraction as a batched matmul (cuBLAS), fwd + bwd, TF32,
at the SA-layer shapes of the DET training step.  Measurement tool."""
import sys, os
import torch
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


for (B, cin, cout, np_, ns) in [(16, 135, 64, 2048, 64), (16, 64, 64, 2048, 64), (16, 64, 128, 2048, 64),
                                (16, 131, 128, 1024, 32), (16, 128, 256, 1024, 32), (16, 259, 128, 512, 16)]:
    x = torch.randn(B, cin, np_, ns, device="cuda", requires_grad=True)
    w = torch.randn(cout, cin, 1, 1, device="cuda", requires_grad=True)
    g = torch.randn(B, cout, np_, ns, device="cuda")

    def conv():
        x.grad = w.grad = None
        y = torch.nn.functional.conv2d(x, w)
        y.backward(g)

    def mm():
        x.grad = w.grad = None
        y = torch.matmul(w.view(cout, cin), x.view(B, cin, np_ * ns)).view(B, cout, np_, ns)
        y.backward(g)

    def conv_f():
        with torch.no_grad():
            torch.nn.functional.conv2d(x, w)

    def mm_f():
        with torch.no_grad():
            torch.matmul(w.view(cout, cin), x.view(B, cin, np_ * ns))

    print("B=%d %d->%d L=%d: conv fwd %.3f fwd+bwd %.3f | matmul fwd %.3f fwd+bwd %.3f ms" % (
        B, cin, cout, np_ * ns, timeit(conv_f), timeit(conv), timeit(mm_f), timeit(mm)), flush=True)
