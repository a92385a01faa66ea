"""Edge kernels at N = 30, every particle real, batch sweep: ncu target for the fixed-cost fit
(duration = fixed + steps_per_cta * period).   python profiles/fixed_cost.py [p_drop]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

from mpgan_b200 import ops

p = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
N, F = 30, 32
torch.manual_seed(0)
ws = []
for i, o in ((2 * F, 96), (96, 160), (160, 192)):
    ws += [(torch.randn(o, i, device="cuda") / i ** 0.5).requires_grad_(True), (torch.randn(o, device="cuda") * 0.1).requires_grad_(True)]
ops.set_precision(1)
for B in (64, 64, 128, 256, 512, 1024):      # the first pass warms everything up
    x = (torch.randn(B, N, F, device="cuda") * 0.5).requires_grad_(True)
    mask = torch.ones(B, N, 1, device="cuda")
    ops.edge_aggregate(x, mask, *ws, p_drop=p).sum().backward()
    torch.cuda.synchronize()
