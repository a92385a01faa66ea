"""In-kernel event trace of the fused node-network kernel (CTA 0): phase durations in ns from globaltimer.
Usage: python profiles/trace_fn.py [M]"""
import ctypes, sys
import torch
sys.path.insert(0, ".")
from mpgan_b200 import ops, _lib

M = int(sys.argv[1]) if len(sys.argv) > 1 else 15360
L = _lib.lib()
g = torch.Generator().manual_seed(0)
Ka, Kb, H, NO = 192, 32, 256, 32
agg = torch.randn(M, Ka, generator=g).cuda().requires_grad_(True)
x = torch.randn(M, Kb, generator=g).cuda().requires_grad_(True)
shapes = [(H, Ka + Kb), (H,), (H, H), (H,), (NO, H), (NO,)]
ws = [(torch.randn(*s, generator=g) / 16).cuda().requires_grad_(True) for s in shapes]
gout = torch.randn(M, NO, generator=g).cuda()
tr = torch.zeros(64, dtype=torch.int64, device="cuda")
names = {15: "entry", 0: "setup done", 1: "A tile built", 2: "D0 ready", 3: "epi0 done", 4: "D1 ready", 5: "epi1 done",
         6: "D2 ready", 9: "A batch0 loads issued", 10: "A batch0 stored", 11: "A loop done", 7: "epi2 done", 8: "exit sync", 16: "mma0 start", 17: "mma0 issued", 18: "mma1 start",
         19: "mma1 issued", 20: "mma2 start", 21: "mma2 issued"}
for mode in ("fwd", "bwd"):
    for _ in range(3):
        out = ops.node_net(agg, x, *ws, 0.2, 0.5)
        out.backward(gout)
    torch.cuda.synchronize()
    out = ops.node_net(agg, x, *ws, 0.2, 0.5)
    torch.cuda.synchronize()
    fn = ctypes.CDLL(_lib.LIB_PATH).mpg_debug_set_fn_trace
    fn.argtypes = [ctypes.c_void_p]
    tr.zero_()
    fn(tr.data_ptr())
    if mode == "fwd":
        out = ops.node_net(agg, x, *ws, 0.2, 0.5)
    else:
        out.backward(gout)
    torch.cuda.synchronize()
    fn(None)
    t = tr.cpu().tolist()
    t0 = t[15]
    print(f"--- {mode} M={M}")
    for k in sorted(names, key=lambda k: t[k]):
        if t[k]:
            print(f"  {names[k]:14s} +{(t[k] - t0) / 1e3:8.2f} us")
