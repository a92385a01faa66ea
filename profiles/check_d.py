"""Debug helper: per-tensor relative errors of the D forward/backward golden case at precision 0 / 1."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from mpgan_b200 import ops, presets

gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
cases = torch.load(os.path.join(gold, "disc_fwd_bwd.pt"))
sdD = torch.load(os.path.join(gold, "mp_d_seed4_weights.pt"))
def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    mx = float((a - b).abs().max()) / max(float(b.abs().max()), 1e-6)
    l2 = float((a - b).norm()) / max(float(b.norm()), 1e-12)
    return mx if os.environ.get("ERR", "max") == "max" else l2
for name, N in (("n30", 30), ("n150", 150)):
    c = cases[name]
    print("==", name, tuple(c["x"].shape))
    for prec in (0, 1):
        ops.set_precision(prec)
        D = presets.mp_discriminator(num_hits=N, disc_dropout=0.0).cuda().train()
        D.load_state_dict(sdD, strict=True)
        x = c["x"].cuda().requires_grad_(True)
        out = D(x, c["labels"].cuda())
        ((out - 1) ** 2).mean().backward()
        errs = {"out": rel(out, c["out"]), "dx": rel(x.grad[..., :3], c["dx"][..., :3])}
        for k, g in c["grads"].items():
            errs[k] = rel(dict(D.named_parameters())[k].grad, g)
        print(f"prec {prec}: " + "  ".join(f"{k.replace('mp_layers.', 'L').replace('.net', '')}={v:.1e}" for k, v in errs.items()))

c = cases["g_through_d"]
sdG = torch.load(os.path.join(gold, "mp_g_weights.pt"))
print("== g_through_d", tuple(c["noise"].shape))
for prec in (0, 1):
    ops.set_precision(prec)
    G = presets.mp_generator().cuda().train()
    G.load_state_dict(sdG, strict=True)
    D = presets.mp_discriminator(disc_dropout=0.0).cuda().train()
    D.load_state_dict(sdD, strict=True)
    labels = c["labels"].cuda()
    fake = G(c["noise"].cuda(), labels)
    fake.retain_grad()
    loss = ((D(fake, labels) - 1) ** 2).mean()
    loss.backward()
    errs = {"loss": abs(float(loss) - float(c["loss"])) / abs(float(c["loss"]))}
    for k, g in c["grads"].items():
        errs[k] = rel(dict(G.named_parameters())[k].grad, g)
    print(f"prec {prec}: " + "  ".join(f"{k.replace('mp_layers.', 'L').replace('.net', '')}={v:.1e}" for k, v in errs.items()))
    if prec == 0:
        ref_fake_grad = fake.grad.clone()
    else:
        print("   dL/dfake rel err vs prec0:", rel(fake.grad[..., :3], ref_fake_grad[..., :3]))
