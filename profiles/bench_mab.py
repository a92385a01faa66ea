"""Device time of one GAPT attention block (forward, forward + backward): the fused kernel vs the per-op path.

    python profiles/bench_mab.py [B]
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

from mpgan_b200 import gapt, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ops.set_precision(1)
torch.manual_seed(0)
lin = dict(leaky_relu_alpha=0.2, dropout_p=0.5, batch_norm=False, spectral_norm=False)
m = gapt.MAB(64, 4, ff_layers=[], final_linear=False, dropout_p=0.5, linear_args=lin).cuda().train()


def timeit(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for Nq, Nk in ((30, 30), (10, 30), (30, 10), (1, 30)):
    x = (torch.randn(B, Nq, 64, device="cuda") * 0.5).requires_grad_(True)
    y = x if Nq == Nk else (torch.randn(B, Nk, 64, device="cuda") * 0.5).requires_grad_(True)
    n = torch.randint(1, Nk + 1, (B,), device="cuda")
    mask = (torch.arange(Nk, device="cuda")[None, :] < n[:, None]).float().unsqueeze(2)
    w = torch.randn(B, Nq, 64, device="cuda")
    for fused in (True, False):
        gapt.MAB.fused = fused

        def fwd():
            with torch.no_grad():
                return m(x, y, mask)

        def fwdbwd():
            out = m(x, y, mask)
            out.backward(w)

        # whole-call graphs: device time without host launch gaps
        g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fwd(); fwdbwd()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        with torch.cuda.graph(g1):
            fwd()
        with torch.cuda.graph(g2):
            fwdbwd()
        t1, t2 = timeit(g1.replay), timeit(g2.replay)
        print(f"B={B} Nq={Nq:2d} Nk={Nk:2d} {'fused ' if fused else 'per-op'}: fwd {t1:7.1f} us   fwd+bwd {t2:7.1f} us")
gapt.MAB.fused = True
