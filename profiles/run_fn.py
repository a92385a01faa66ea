"""Runs only the fused node network (forward + backward) a few times: target for `ncu -k regex:fn_`.

    python profiles/run_fn.py [M] [p_drop] [iters]
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

from mpgan_b200 import ops

M = int(sys.argv[1]) if len(sys.argv) > 1 else 76800
p = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
Ka, Kb, H, NO = 192, 32, 256, 32
g = torch.Generator().manual_seed(0)
agg = torch.randn(M, Ka, generator=g).cuda().requires_grad_(True)
x = torch.randn(M, Kb, generator=g).cuda().requires_grad_(True)
shapes = [(H, Ka + Kb), (H,), (H, H), (H,), (NO, H), (NO,)]
ws = [(torch.randn(*s, generator=g) / 16).cuda().requires_grad_(True) for s in shapes]
gout = torch.randn(M, NO, generator=g).cuda()
ops.set_precision(1)
for it in range(iters):
    out = ops.node_net(agg, x, *ws, 0.2, p)
    out.backward(gout)
    torch.cuda.synchronize()
print("done")
