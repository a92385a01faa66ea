"""Event trace of the backward CHAIN kernel (library variant built with
MPG_LIB_VARIANT=trace MPG_NVCC_FLAGS=-DMPG_TRACE python -m mpgan_b200.build; run with MPG_LIB_VARIANT=trace).
Weights are frozen so only CHAIN launches (DW2 shares stamp slots 0-2)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

from mpgan_b200 import _lib, ops

B, N, p = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
F = 32
torch.manual_seed(0)
x = (torch.randn(B, N, F, device="cuda") * 0.5).requires_grad_(True)
mask = torch.ones(B, N, 1, device="cuda")
if len(sys.argv) > 4 and sys.argv[4] == "rand":   # particle counts ~ U{1..N}, sorted as the trainer sorts them
    n = torch.randint(1, N + 1, (B,), device="cuda").sort(descending=True).values
    mask = (torch.arange(N, device="cuda")[None, :] < n[:, None]).float().unsqueeze(2)
ws = []
for i, o in ((2 * F, 96), (96, 160), (160, 192)):
    ws += [torch.randn(o, i, device="cuda") / i ** 0.5, torch.randn(o, device="cuda") * 0.1]
ops.set_precision(1)
L = _lib.lib()
L.mpg_debug_set_trace.argtypes = [ctypes.c_void_p]
dagg = torch.randn(B, N, 192, device="cuda")
ops.edge_aggregate(x, mask, *ws, p_drop=p).backward(dagg)   # warm
agg = ops.edge_aggregate(x, mask, *ws, p_drop=p)
torch.cuda.synchronize()
tr = torch.zeros(256 * 16, dtype=torch.int64, device="cuda")
assert L.mpg_debug_set_trace(tr.data_ptr()) == 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
agg.backward(dagg)
e1.record()
torch.cuda.synchronize()
print(f"backward (all launches) {e0.elapsed_time(e1) * 1e3:.1f} us")
L.mpg_debug_set_trace(None)
t = tr.cpu().view(256, 16)
names = ["top", "D1_seen", "E1_done", "D2_seen", "E2_done", "dH1_seen", "E3_done", "dW1_done_seen", "H0next_built",
         "dH0_seen", "E4_done", "I:H1_seen", "I:M2_issued", "I:G2_seen", "I:G1_seen", "I:G0_seen"]
for it in range(3, 8):
    t0 = int(t[it, 0])
    ev = sorted((int(t[it, k]) - t0, names[k]) for k in range(16) if int(t[it, k]) > 0)
    print(f"step {it}: period {int(t[it + 1, 0]) - t0} clk: " + "  ".join(f"{n}@{d}" for d, n in ev))
per = [int(t[i + 1, 0]) - int(t[i, 0]) for i in range(3, 60) if int(t[i + 1, 0]) > 0]
print("mean period over steps 3..60:", sum(per) / max(len(per), 1), "clk")
ph = [int(v) for v in t[255, :8]]
pre = [int(v) for v in t[255, 8:14]]
print("pre-loop (clk from entry): " + "  ".join(f"{n}@{v - ph[0]}" for n, v in zip(
    ["regs", "setup done", "P+dAgg loaded", "onehot", "Q stage 0 landed", "H0'(0) built"], pre) if v > 0))
nst = max(i for i in range(255) if int(t[i, 0]) > 0) + 1
pn = ["entry", "tmem+barriers", "loop top", "loop end", "last MMA done", "dq flushed", "slab written", "exit sync"]
print(f"CTA 0: {nst} steps; phases (clk from entry): " + "  ".join(f"{n}@{v - ph[0]}" for n, v in zip(pn, ph) if v > 0))
print(f"first step top @{int(t[0, 0]) - ph[0]}, last step top @{int(t[nst - 1, 0]) - ph[0]}")
