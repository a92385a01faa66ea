"""Debug helper: tcgen05 edge kernels vs the fp32 SIMT kernels, forward then backward, one shape.

    CUDA_LAUNCH_BLOCKING=1 python profiles/check_edge.py B N p_drop
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

import mpgan_b200.ops as O

B, N, p = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
F = int(sys.argv[4]) if len(sys.argv) > 4 else 32
dlike = len(sys.argv) > 5          # discriminator-like data: features zeroed on padded rows, dAgg zero there too
torch.manual_seed(1)
x0 = torch.randn(B, N, F, device="cuda") * 0.5
n = torch.randint(1, N + 1, (B,), device="cuda")
mask = (torch.arange(N, device="cuda")[None, :] < n[:, None]).float().unsqueeze(2)
if dlike:
    x0 = (torch.rand(B, N, F, device="cuda") - 0.5) * mask
ws0 = []
for i, o in ((2 * F, 96), (96, 160), (160, 192)):
    ws0 += [torch.randn(o, i, device="cuda") / i ** 0.5, torch.randn(o, device="cuda") * 0.1]
dagg = torch.randn(B, N, 192, device="cuda")
if dlike:
    dagg = dagg * mask * 1e-3
res = []
for prec in (0, 1):
    O.set_precision(prec)
    O._seed_counter = 4242
    x = x0.clone().requires_grad_(True)
    ws = [w.clone().requires_grad_(True) for w in ws0]
    agg = O.edge_aggregate(x, mask, *ws, p_drop=p)
    torch.cuda.synchronize()
    print(f"prec {prec} forward ok", flush=True)
    agg.backward(dagg)
    torch.cuda.synchronize()
    print(f"prec {prec} backward ok", flush=True)
    res.append([agg.detach(), x.grad] + [w.grad for w in ws])
names = ["agg", "dx", "dW0", "db0", "dW1", "db1", "dW2", "db2"]
for name, r0, r1 in zip(names, res[0], res[1]):
    err = float((r1 - r0).abs().max()) / max(float(r0.abs().max()), 1e-9)
    print(f"{name:4s} rel err {err:.3e}   max |ref| {float(r0.abs().max()):.3e}")
