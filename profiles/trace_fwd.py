"""Event trace of the forward edge kernel (library built with MPG_NVCC_FLAGS=-DMPG_TRACE)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

from mpgan_b200 import _lib, ops

B, N, p = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
F = 32
torch.manual_seed(0)
x = torch.randn(B, N, F, device="cuda") * 0.5
mask = torch.ones(B, N, 1, device="cuda")
ws = []
for i, o in ((2 * F, 96), (96, 160), (160, 192)):
    ws += [torch.randn(o, i, device="cuda") / i ** 0.5, torch.randn(o, device="cuda") * 0.1]
ops.set_precision(1)
with torch.no_grad():
    ops.edge_aggregate(x, mask, *ws, p_drop=p)
    tr = torch.zeros(256 * 16, dtype=torch.int64, device="cuda")  # [4080:4088] = phase stamps
    L = _lib.lib()
    L.mpg_debug_set_trace.argtypes = [ctypes.c_void_p]
    assert L.mpg_debug_set_trace(tr.data_ptr()) == 0
    ops.edge_aggregate(x, mask, *ws, p_drop=p)
    torch.cuda.synchronize()
t = tr.cpu().view(256, 16)
names = ["d2lo_seen", "f2lo_arr", "d2hi_seen", "d1_seen", "h1_arr", "f2hi_arr", "d1n_seen", "h0_arr",
         "I:h0_seen", "I:m1_done", "I:h1_seen", "I:f2lo_ok", "I:m2lo_iss", "I:f2hi_ok", "I:m2hi_iss"]
for it in range(2, 6):
    t0 = int(t[it, 0])
    ev = sorted((int(t[it, k]) - t0, names[k]) for k in range(15) if int(t[it, k]) > 0)
    print(f"step {it}: period {int(t[it + 1, 0]) - t0} clk: " + "  ".join(f"{n}@{d}" for d, n in ev))

ph = [int(v) for v in tr.cpu()[4080:4088]]
names_p = ["entry", "setup_done", "pre_h0", "h0_0_built", "loop_end", "drain_e2", "flush_end", "final_sync"]
print("phases (clk from entry): " + "  ".join(f"{n}@{v - ph[0]}" for n, v in zip(names_p, ph) if v))
