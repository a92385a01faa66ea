"""Debug helper: run-to-run reproducibility of the tcgen05 edge backward (atomics reorder fp32 sums by ~1e-7)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import mpgan_b200.ops as O

for B, N, F in ((6, 30, 32), (6, 30, 3), (2, 150, 32), (40, 30, 32), (256, 30, 32)):
    torch.manual_seed(1)
    x0 = torch.randn(B, N, F, device="cuda") * 0.5
    n = torch.randint(1, N + 1, (B,), device="cuda")
    mask = (torch.arange(N, device="cuda")[None, :] < n[:, None]).float().unsqueeze(2)
    ws0 = []
    for i, o in ((2 * F, 96), (96, 160), (160, 192)):
        ws0 += [torch.randn(o, i, device="cuda") / i ** 0.5, torch.randn(o, device="cuda") * 0.1]
    dagg = torch.randn(B, N, 192, device="cuda")
    O.set_precision(1)
    runs = []
    for rep in range(4):
        x = x0.clone().requires_grad_(True)
        ws = [w.clone().requires_grad_(True) for w in ws0]
        agg = O.edge_aggregate(x, mask, *ws, p_drop=0.0)
        agg.backward(dagg)
        torch.cuda.synchronize()
        runs.append([agg.detach(), x.grad] + [w.grad for w in ws])
    names = ["agg", "dx", "dW0", "db0", "dW1", "db1", "dW2", "db2"]
    worst = {}
    for rep in range(1, 4):
        for nme, a, b in zip(names, runs[0], runs[rep]):
            worst[nme] = max(worst.get(nme, 0.0), float((a - b).abs().max()) / max(float(a.abs().max()), 1e-9))
    print(f"B={B} N={N} F={F}: " + "  ".join(f"{k}={v:.1e}" for k, v in worst.items()))
