"""Developer tool: structured inputs through the tcgen05 TF32 conv kernels to see layout errors."""
import sys, torch
sys.path.insert(0, ".")
from bridgeqa_b200 import train_fused
torch.set_printoptions(linewidth=200, precision=3, sci_mode=False)
b, cin, cout, p = 1, 32, 128, 128
x = torch.zeros(b, cin, p)
for c in range(cin):
    x[0, c] = c + torch.arange(p) * 0.001
w = torch.zeros(cout, cin)
for o in range(cout):
    w[o, o % cin] = 1.0
x, w = x.cuda(), w.cuda()
y = train_fused._conv_forward(x, w)
want = torch.matmul(w, x[0])
print("max err", float((y[0] - want).abs().max()))
print("y[0, :8, :8]\n", y[0, :8, :8].cpu())
print("want[:8, :8]\n", want[:8, :8].cpu())
print("y[0, 0, ::8]\n", y[0, 0, ::8].cpu())
print("y[0, ::16, 0]\n", y[0, ::16, 0].cpu())
# which (channel, position) does each output hold?
ch = torch.round(y[0]).long().cpu(); pos = torch.round((y[0].cpu() - ch) * 1000).long()
print("decoded channel of y[o, 0] for o in 0..39:", ch[:40, 0].tolist())
print("decoded position of y[0, p] for p in 0..39:", pos[0, :40].tolist())
print("decoded position of y[1, p] for p in 0..39:", pos[1, :40].tolist())
print("---- wgrad")
b, cin, cout, p = 1, 32, 128, 256
x = torch.randn(b, cin, p).cuda(); dy = torch.randn(b, cout, p).cuda()
dw = train_fused._conv_wgrad(x, dy, cout, cin)
want = torch.einsum("bop,bcp->oc", dy.double(), x.double())
print("wgrad max err", float((dw.double() - want).abs().max()), "max |want|", float(want.abs().max()))
print(dw[:4, :8].cpu()); print(want[:4, :8].cpu())
