"""Per-tensor error of the fused tcgen05 node network (forward, input and parameter gradients) against fp64
torch math.  Usage: python profiles/check_fn.py [M] [Kb] [NO] [p]"""
import sys
import torch
sys.path.insert(0, ".")
from mpgan_b200 import ops

M = int(sys.argv[1]) if len(sys.argv) > 1 else 300
Kb = int(sys.argv[2]) if len(sys.argv) > 2 else 32
NO = int(sys.argv[3]) if len(sys.argv) > 3 else 32
Ka, H = 192, 256
g = torch.Generator().manual_seed(11)
agg = torch.randn(M, Ka, generator=g).cuda().requires_grad_(True)
x = torch.randn(M, Kb, generator=g).cuda().requires_grad_(True)
shapes = [(H, Ka + Kb), (H,), (H, H), (H,), (NO, H), (NO,)]
ws = [(torch.randn(*s, generator=g) / (s[-1] ** 0.5 if len(s) == 2 else 4.0)).cuda().requires_grad_(True) for s in shapes]
out = ops.node_net(agg, x, *ws, 0.2, 0.0)
gout = torch.randn(M, NO, generator=g).cuda()
out.backward(gout)
torch.cuda.synchronize()
a_r, x_r = agg.detach().double().requires_grad_(True), x.detach().double().requires_grad_(True)
w_r = [w.detach().double().requires_grad_(True) for w in ws]
h = torch.cat((a_r, x_r), 1)
y0 = torch.nn.functional.leaky_relu(h @ w_r[0].t() + w_r[1], 0.2)
y1 = torch.nn.functional.leaky_relu(y0 @ w_r[2].t() + w_r[3], 0.2)
o = y1 @ w_r[4].t() + w_r[5]
o.backward(gout.double())
names = ["out", "dagg", "dx", "dw0", "db0", "dw1", "db1", "dw2", "db2"]
got = [out.detach(), agg.grad, x.grad] + [w.grad for w in ws]
ref = [o.detach(), a_r.grad, x_r.grad] + [w.grad for w in w_r]
for n, a, b in zip(names, got, ref):
    a = a.double()
    print(f"{n:5s} relL2 {float((a - b).norm() / b.norm()):.3e}  maxabs {float((a - b).abs().max()):.3e}  |ref| {float(b.abs().max()):.3e} |got| {float(a.abs().max()):.3e}")
if "-v" in sys.argv:
    print(got[3][:4, :6]); print(ref[3][:4, :6])
    print(got[5][:4, :6]); print(ref[5][:4, :6])
