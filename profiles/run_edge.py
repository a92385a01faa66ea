"""Runs only the fused edge op (forward + backward) a few times: target for `ncu -k regex:edge_tc`.

    python profiles/run_edge.py [B] [N] [p_drop] [iters]
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

from mpgan_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 150
p = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
F = 32
torch.manual_seed(0)
dev = "cuda"
x = (torch.randn(B, N, F, device=dev) * 0.5).requires_grad_(True)
n = torch.randint(1, N + 1, (B,), device=dev)
mask = (torch.arange(N, device=dev)[None, :] < n[:, None]).float().unsqueeze(2)
ws = []
for i, o in ((2 * F, 96), (96, 160), (160, 192)):
    ws += [(torch.randn(o, i, device=dev) / i ** 0.5).requires_grad_(True), (torch.randn(o, device=dev) * 0.1).requires_grad_(True)]
ops.set_precision(1)
for it in range(iters):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    agg = ops.edge_aggregate(x, mask, *ws, p_drop=p)
    e1.record()
    agg.sum().backward()
    e2.record()
    torch.cuda.synchronize()
    fl = ops.edge_flops(B, N, F, 96, 160, 192)
    print(f"iter {it}: fwd {e0.elapsed_time(e1):.3f} ms ({fl / e0.elapsed_time(e1) / 1e9:.1f} TF/s)  "
          f"bwd {e1.elapsed_time(e2):.3f} ms ({2 * fl / e1.elapsed_time(e2) / 1e9:.1f} TF/s)")
