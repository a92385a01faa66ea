"""Mint golden vectors from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden.py            # needs /root/reference, writes tests/golden/*.pt

The reference cannot travel to the GPU box, so its outputs on seeded inputs are committed as
fixtures; ``tests/test_oracle_golden.py`` pins the oracle restatement (``oracle/*.py``) to them
and the ``-m gpu`` tests pin the CUDA path to both.  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys
from unittest.mock import MagicMock

import torch

REF = os.environ.get("MPGAN_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
sys.path.insert(0, REF)
for m in ("jetnet", "jetnet.datasets", "jetnet.evaluation", "jetnet.datasets.normalisations",
          "jetnet.utils", "matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.cm",
          "matplotlib.lines", "mplhep"):
    sys.modules.setdefault(m, MagicMock())

import gapt  # noqa: E402
import mpgan  # noqa: E402
import setup_training  # noqa: E402
import train as ref_train  # noqa: E402
from gapt.model import GAPT_D, GAPT_G  # noqa: E402
from mpgan.model import LinearNet, MPLayer  # noqa: E402


class Obj:
    def __init__(self, d):
        self.__dict__ = dict(d)


def mp_args(**over):
    a = eval(open(f"{REF}/trained_models/mp_g/args.txt").read())
    a.update(device="cpu", load_model=False, multi_gpu=False)
    a.update(over)
    return Obj(a)


def synthetic_jets(B, N, gen, all_real=False):
    n = torch.full((B,), N) if all_real else torch.randint(1, N + 1, (B,), generator=gen)
    real = (torch.arange(N)[None, :] < n[:, None]).float().unsqueeze(2)
    feats = (torch.rand(B, N, 3, generator=gen) - 0.5) * real
    x = torch.cat((feats, real - 0.5), dim=2)
    labels = (n.float() * torch.tensor(1.0 / N, dtype=torch.float32)).unsqueeze(1)
    return x, labels, n


def grads_of(module):
    return {k: p.grad.clone() for k, p in module.named_parameters() if p.grad is not None}


def save(name, obj):
    os.makedirs(OUT, exist_ok=True)
    torch.save(obj, os.path.join(OUT, name))
    print(name, os.path.getsize(os.path.join(OUT, name)) // 1024, "KiB")


def main():
    torch.set_num_threads(8)
    # ------------------------------------------------------------------ weights
    G = setup_training.setup_mpgan(mp_args(), gen=True)
    sdG = torch.load(f"{REF}/trained_models/mp_g/G_best_epoch.pt", map_location="cpu")
    print(G.load_state_dict(sdG, strict=True))
    torch.manual_seed(4)
    D = setup_training.setup_mpgan(mp_args(disc_dropout=0.0), gen=False)
    sdD = {k: v.clone() for k, v in D.state_dict().items()}
    save("mp_g_weights.pt", {k: v.clone().float() for k, v in sdG.items()})
    save("mp_d_seed4_weights.pt", sdD)
    for jet in ("q", "t"):  # load-unchanged contract: only key/shape manifests for mp_q / mp_t
        sd = torch.load(f"{REF}/trained_models/mp_{jet}/G_best_epoch.pt", map_location="cpu")
        G.load_state_dict(sd, strict=True)
    G.load_state_dict(sdG)
    save("mp_state_manifest.pt", {"G": {k: tuple(v.shape) for k, v in sdG.items()},
                                  "D": {k: tuple(v.shape) for k, v in sdD.items()}})

    # ------------------------------------------------------------------ G forward (SURVEY section 4 vector)
    G.eval()
    g = torch.Generator().manual_seed(1234)
    noise = torch.randn(4, 30, 32, generator=g) * 0.2
    labels = torch.tensor([30, 17, 5, 1.0]).unsqueeze(1) / 30
    with torch.no_grad():
        out = G(noise, labels)
    print("survey vector:", out[0, 0].tolist(), float(out.double().sum()))
    cases = {"survey4": dict(noise=noise, labels=labels, out=out)}
    g = torch.Generator().manual_seed(7)
    n = torch.randint(1, 31, (64,), generator=g)
    noise = torch.randn(64, 30, 32, generator=g) * 0.2
    labels = (n.float() * torch.tensor(1.0 / 30)).unsqueeze(1)
    with torch.no_grad():
        out = G(noise, labels)
    cases["b64"] = dict(noise=noise, labels=labels, out=out)
    # N=150 and N=100 with the same weights (architecture is N-independent)
    for N in (100, 150):
        GN = setup_training.setup_mpgan(mp_args(num_hits=N), gen=True)
        GN.load_state_dict(sdG)
        GN.eval()
        g = torch.Generator().manual_seed(N)
        n = torch.tensor([N, N // 2, 1])
        noise = torch.randn(3, N, 32, generator=g) * 0.2
        labels = (n.float() * torch.tensor(1.0 / N)).unsqueeze(1)
        with torch.no_grad():
            out = GN(noise, labels)
        cases[f"n{N}"] = dict(noise=noise, labels=labels, out=out)
    save("gen_forward.pt", cases)

    # ------------------------------------------------------------------ D forward + backward (eval: no dropout)
    cases = {}
    for N, B in ((30, 6), (150, 2)):
        DN = setup_training.setup_mpgan(mp_args(num_hits=N, disc_dropout=0.0), gen=False)
        DN.load_state_dict(sdD)
        DN.train()
        g = torch.Generator().manual_seed(100 + N)
        x, labels, n = synthetic_jets(B, N, g)
        x.requires_grad_(True)
        out = DN(x, labels)
        loss = ref_train.mse(out, torch.ones(B, 1))
        DN.zero_grad()
        loss.backward()
        cases[f"n{N}"] = dict(x=x.detach().clone(), labels=labels, out=out.detach(), loss=loss.detach(),
                              grads=grads_of(DN), dx=x.grad.clone())
    # G -> D: G-loss gradients for G params (N=30)
    D.load_state_dict(sdD)
    D.train()
    G.train()
    g = torch.Generator().manual_seed(11)
    n = torch.randint(1, 31, (6,), generator=g)
    noise = torch.randn(6, 30, 32, generator=g) * 0.2
    labels = (n.float() * torch.tensor(1.0 / 30)).unsqueeze(1)
    G.zero_grad()
    loss = ref_train.calc_G_loss("ls", D(G(noise, labels), labels))
    loss.backward()
    cases["g_through_d"] = dict(noise=noise, labels=labels, loss=loss.detach(), grads=grads_of(G))
    save("disc_fwd_bwd.pt", cases)

    # ------------------------------------------------------------------ MPLayer variants (small sizes)
    cases = {}
    variants = {
        "plain_sum": dict(),
        "plain_mean": dict(sum=False),
        "posdiff_allef": dict(pos_diffs=True, all_ef=True, delta_r=False),
        "posdiff_deltar": dict(pos_diffs=True, all_ef=False, delta_r=True),
        "posdiff_coords_r": dict(pos_diffs=True, all_ef=False, delta_r=True, delta_coords=True),
        "posdiff_coords": dict(pos_diffs=True, all_ef=False, delta_r=False, delta_coords=True),
    }
    for name, kw in variants.items():
        torch.manual_seed(21)
        layer = MPLayer(5, [16, 24, 32], [40, 40], 6, **kw)
        g = torch.Generator().manual_seed(22)
        x = torch.randn(3, 7, 5, generator=g).requires_grad_(True)
        mask = (torch.rand(3, 7, 1, generator=g) > 0.3).float()
        out = layer(x, True, mask)
        w = torch.randn(out.shape, generator=g)
        (out * w).sum().backward()
        cases[name] = dict(kw=kw, sd={k: v.clone() for k, v in layer.state_dict().items()},
                           x=x.detach().clone(), mask=mask, w=w, out=out.detach(),
                           grads=grads_of(layer), dx=x.grad.clone())
    # unmasked, default sizes, F=3 (D first layer shape) and F=32
    for F_in, F_out in ((3, 32), (32, 3)):
        torch.manual_seed(23)
        layer = MPLayer(F_in, [96, 160, 192], [256, 256], F_out)
        g = torch.Generator().manual_seed(24)
        x = (torch.randn(2, 30, F_in, generator=g) * 0.3).requires_grad_(True)
        out = layer(x)
        w = torch.randn(out.shape, generator=g)
        (out * w).sum().backward()
        cases[f"default_F{F_in}"] = dict(kw={}, sd={k: v.clone() for k, v in layer.state_dict().items()},
                                          x=x.detach().clone(), mask=None, w=w, out=out.detach(),
                                          grads=grads_of(layer), dx=x.grad.clone())
    save("mplayer_variants.pt", cases)

    # ------------------------------------------------------------------ rank-mask truncation sweep (bit exact)
    cases = {}
    for N in (30, 100, 150):
        g = torch.Generator().manual_seed(N + 1)
        x0 = torch.randn(N, N, generator=g) * 0.2  # jet k has n = k+1 particles
        n = torch.arange(1, N + 1)
        GN = setup_training.setup_mpgan(mp_args(num_hits=N), gen=True)
        for conv, lab in (("div", n.float() / N), ("mul", n.float() * torch.tensor(1.0 / N))):
            xin = torch.zeros(N, N, 32)
            xin[:, :, 0] = x0
            _, _, mask, njp = GN._get_mask(xin, lab.unsqueeze(1), **GN.mask_args)
            cases[f"n{N}_{conv}"] = dict(x0=x0, labels=lab.unsqueeze(1), mask=mask.to(torch.uint8),
                                         count=mask.sum((1, 2)).int())
    save("rank_mask.pt", cases)

    # ------------------------------------------------------------------ spectral norm LinearNet
    torch.manual_seed(31)
    net = LinearNet([24, 16], input_size=10, output_size=4, final_linear=True, spectral_norm=True)
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(32)
    x = torch.randn(9, 10, generator=g).requires_grad_(True)
    out = net(x)
    w = torch.randn(out.shape, generator=g)
    (out * w).sum().backward()
    save("spectral_norm.pt", dict(sd0=sd0, sd1={k: v.clone() for k, v in net.state_dict().items()},
                                  x=x.detach().clone(), w=w, out=out.detach(),
                                  grads=grads_of(net), dx=x.grad.clone()))

    # ------------------------------------------------------------------ GAPT (SAB and ISAB), eval-mode dropout 0
    cases = {}
    for isab in (False, True):
        common = dict(num_particles=30, num_heads=4, embed_dim=64, sab_fc_layers=[], use_mask=True,
                      use_isab=isab, num_isab_nodes=10)
        lin = dict(leaky_relu_alpha=0.2, dropout_p=0.0, batch_norm=False, spectral_norm=False)
        torch.manual_seed(41)
        GG = GAPT_G(sab_layers=4, output_feat_size=3, final_fc_layers=[], dropout_p=0.0,
                    layer_norm=False, linear_args=lin, **common)
        GD = GAPT_D(sab_layers=2, input_feat_size=3, final_fc_layers=[], dropout_p=0.0,
                    layer_norm=False, linear_args=lin, **common)
        g = torch.Generator().manual_seed(42)
        n = torch.tensor([30, 17, 5, 1, 29])
        noise = (torch.randn(5, 30, 64, generator=g) * 0.2).requires_grad_(True)
        labels = (n.float() * torch.tensor(1.0 / 30)).unsqueeze(1)
        fake = GG(noise, labels)
        dout = GD(fake, labels)
        loss = ref_train.calc_G_loss("ls", dout)
        loss.backward()
        x, xl, _ = synthetic_jets(5, 30, g)
        x.requires_grad_(True)
        gG = grads_of(GG)
        GD.zero_grad()
        rout = GD(x, xl)
        rl = ref_train.mse(rout, torch.ones(5, 1))
        rl.backward()
        cases["isab" if isab else "sab"] = dict(
            sdG={k: v.clone() for k, v in GG.state_dict().items()},
            sdD={k: v.clone() for k, v in GD.state_dict().items()},
            noise=noise.detach().clone(), labels=labels, fake=fake.detach(), dout=dout.detach(),
            loss=loss.detach(), gradsG=gG, dnoise=noise.grad.clone(),
            x=x.detach().clone(), xlabels=xl, rout=rout.detach(), gradsD=grads_of(GD), dx=x.grad.clone())
    save("gapt.pt", cases)

    # ------------------------------------------------------------------ one full G+D step through train.py
    torch.manual_seed(4)
    args = mp_args(disc_dropout=0.0)
    G2 = setup_training.setup_mpgan(args, gen=True)
    G2.load_state_dict(sdG)
    D2 = setup_training.setup_mpgan(args, gen=False)
    D2.load_state_dict(sdD)
    args.spectral_norm_gen = False
    G_opt, D_opt = setup_training.optimizers(args, G2, D2)
    model_args = {"lfc": False, "lfc_latent_size": 128, "mask_learn_sep": False, "latent_node_size": 32}
    g = torch.Generator().manual_seed(51)
    data, labels, _ = synthetic_jets(8, 30, g)
    noise_d = torch.randn(8, 30, 32, generator=g) * 0.2
    noise_g = torch.randn(8, 30, 32, generator=g) * 0.2
    d_items = ref_train.train_D(model_args, D2, G2, D_opt, G_opt, data, "ls", labels=labels,
                                gen_args={"num_particles": 30, "noise": noise_d})
    gradsD = grads_of(D2)
    g_item = ref_train.train_G(model_args, D2, G2, G_opt, "ls", 8, labels=labels,
                               gen_args={"num_particles": 30, "noise": noise_g})
    gradsG = grads_of(G2)
    save("train_step.pt", dict(data=data, labels=labels, noise_d=noise_d, noise_g=noise_g,
                               loss_d=d_items["D"], loss_g=g_item, gradsD=gradsD, gradsG=gradsG,
                               lr_d=args.lr_disc, lr_g=args.lr_gen,
                               sdD_after={k: v.clone() for k, v in D2.state_dict().items()},
                               sdG_after_sample={k: v.flatten()[:64].clone() for k, v in G2.state_dict().items()}))


if __name__ == "__main__":
    main()
