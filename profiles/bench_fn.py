"""Device time per call (CUDA events, warm, back to back) of the node-level ops at the row counts of the bench
workloads: fused node network fwd / bwd (input + weight gradients), and for comparison the per-layer path.
Usage: python profiles/bench_fn.py"""
import sys
import torch
sys.path.insert(0, ".")
from mpgan_b200 import ops, model


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


g = torch.Generator().manual_seed(0)
Ka, Kb, H, NO = 192, 32, 256, 32
for M in (7680, 15360, 38400, 76800):
    for p in (0.0, 0.5):
        agg = torch.randn(M, Ka, generator=g).cuda().requires_grad_(True)
        x = torch.randn(M, Kb, generator=g).cuda().requires_grad_(True)
        shapes = [(H, Ka + Kb), (H,), (H, H), (H,), (NO, H), (NO,)]
        ws = [(torch.randn(*s, generator=g) / 16).cuda().requires_grad_(True) for s in shapes]
        gout = torch.randn(M, NO, generator=g).cuda()
        t_f = timeit(lambda: ops.node_net(agg.detach(), x.detach(), *[w.detach() for w in ws], 0.2, p))

        def fb():
            out = ops.node_net(agg, x, *ws, 0.2, p)
            out.backward(gout)
        t_fb = timeit(fb)
        for w in ws:
            w.requires_grad_(False)
        t_fb_dx = timeit(fb)
        for w in ws:
            w.requires_grad_(True)

        def layerwise():
            h = torch.cat((agg, x), 1)
            for i in range(3):
                h = ops.linear(h, ws[2 * i], ws[2 * i + 1], i < 2, 0.2, p, 16 + i, 1)
            return h
        t_lf = timeit(lambda: layerwise())

        def lfb():
            layerwise().backward(gout)
        t_lfb = timeit(lfb)
        print(f"M={M:6d} p={p}: fused fwd {t_f:7.1f} us  fwd+bwd {t_fb:7.1f} us  fwd+bwd(dx only) {t_fb_dx:7.1f} us | "
              f"per-layer fwd {t_lf:7.1f} us  fwd+bwd {t_lfb:7.1f} us")
