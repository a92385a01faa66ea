"""Per-tensor error of the CUDA path against the golden vectors minted from the unmodified reference, at
precision 0 (fp32-class) and precision 1 (bf16 tcgen05 edge network + TF32 node GEMMs: the benchmarked mode).

    python profiles/error_table.py > profiles/r2_error_table.txt        (on a B200)

Columns: max-abs error / max-abs of the reference tensor, and relative L2 norm of the error.  The tolerances in
tests/test_gpu_parity.py and tests/test_gpu_train_mode.py are set from this table.
"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mpgan_b200 import ops, presets, train  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def load(name):
    return torch.load(os.path.join(GOLD, name), map_location="cpu", weights_only=False)


def err(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return (float((a - b).abs().max()) / max(float(b.abs().max()), 1e-12),
            float((a - b).norm()) / max(float(b.norm()), 1e-20))


def short(k):
    return k.replace("mp_layers.", "L").replace(".net.", ".").replace(".weight", ".w").replace(".bias", ".b")


ROWS = []


def row(case, prec, name, a, b):
    m, l2 = err(a, b)
    ROWS.append((case, prec, name, m, l2))
    print(f"{case:28s} prec {prec}  {short(name):22s} max {m:9.2e}   L2 {l2:9.2e}")


def models(N, sdG, sdD, dropout=0.0, **over):
    G = presets.mp_generator(num_hits=N).cuda().train()
    D = presets.mp_discriminator(num_hits=N, disc_dropout=dropout, **over).cuda().train()
    G.load_state_dict(sdG)
    D.load_state_dict(sdD)
    return G, D


def main():
    sdG, sdD = load("mp_g_weights.pt"), load("mp_d_seed4_weights.pt")
    gen_cases = load("gen_forward.pt")
    d_cases = {**load("disc_fwd_bwd.pt"), **load("disc_fwd_bwd_large.pt")}
    for prec in (0, 1):
        ops.set_precision(prec)
        for name, N in (("survey4", 30), ("b64", 30), ("n100", 100), ("n150", 150)):
            c = gen_cases[name]
            G, _ = models(N, sdG, sdD)
            with torch.no_grad():
                out = G.eval()(c["noise"].cuda(), c["labels"].cuda())
            row("gen_forward/" + name, prec, "out", out, c["out"])
        for name, N in (("n30", 30), ("d_n100", 100), ("n150", 150)):
            c = d_cases[name]
            _, D = models(N, sdG, sdD)
            x = c["x"].cuda().requires_grad_(True)
            out = D(x, c["labels"].cuda())
            ((out - 1) ** 2).mean().backward()
            row("D_fwd_bwd/" + name, prec, "out", out, c["out"])
            row("D_fwd_bwd/" + name, prec, "dx", x.grad[..., :3], c["dx"][..., :3])
            for k, g in c["grads"].items():
                row("D_fwd_bwd/" + name, prec, k, dict(D.named_parameters())[k].grad, g)
        for name, N in (("g_through_d", 30), ("g_through_d_n100", 100), ("g_through_d_n150", 150)):
            c = d_cases[name]
            G, D = models(N, sdG, sdD)
            labels = c["labels"].cuda()
            loss = ((D(G(c["noise"].cuda(), labels), labels) - 1) ** 2).mean()
            loss.backward()
            row(name, prec, "loss", loss, c["loss"])
            for k, g in c["grads"].items():
                row(name, prec, k, dict(G.named_parameters())[k].grad, g)
        for fname, N in (("train_step.pt", 30), ("train_step_n100.pt", 100)):
            c = load(fname)
            G, D = models(N, sdG, sdD)
            tr = train.GANTrainer(G, D, lr_gen=c["lr_g"], lr_disc=c["lr_d"], num_particles=N)
            labels = c["labels"].cuda()
            ld = tr.train_D(c["data"].cuda(), labels, noise=c["noise_d"].cuda())
            gD = tr.named_grads("D")
            lg = tr.train_G(labels, noise=c["noise_g"].cuda())
            gG = tr.named_grads("G")
            row(fname, prec, "loss_d", ld, torch.tensor(c["loss_d"]))
            row(fname, prec, "loss_g", lg, torch.tensor(c["loss_g"]))
            for k, g in c["gradsD"].items():
                row(fname, prec, "D." + k, gD[k], g)
            for k, g in c["gradsG"].items():
                row(fname, prec, "G." + k, gG[k], g)
    ops.set_precision(1)
    print("\n== worst case per (kind, precision) ==")
    for prec in (0, 1):
        fw = [r for r in ROWS if r[1] == prec and r[2] in ("out", "loss", "loss_d", "loss_g")]
        gr = [r for r in ROWS if r[1] == prec and r not in fw]
        print(f"precision {prec}: forward max-abs {max(r[3] for r in fw):.2e}; gradients max-abs {max(r[3] for r in gr):.2e}, "
              f"L2 {max(r[4] for r in gr):.2e}")


if __name__ == "__main__":
    main()
