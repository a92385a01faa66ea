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


def base():
    """Round-1 fixtures."""
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


def _mp_weights():
    sdG = torch.load(os.path.join(OUT, "mp_g_weights.pt"), map_location="cpu")
    sdD = torch.load(os.path.join(OUT, "mp_d_seed4_weights.pt"), map_location="cpu")
    return sdG, sdD


def extra():
    """Round-2 fixtures: N=100/150 D and G-through-D gradients, option variants (lfc, dea=False, mean aggregation +
    mean pooling), an N=100 training step, WGAN-GP (train.gradient_penalty), kNN / conditioning-column MPLayer
    variants and GAPT with LayerNorm."""
    torch.set_num_threads(8)
    sdG, sdD = _mp_weights()

    # ------------------------------------------------------------------ D fwd+bwd at N=100; G-through-D at N=100/150
    cases = {}
    DN = setup_training.setup_mpgan(mp_args(num_hits=100, disc_dropout=0.0), gen=False)
    DN.load_state_dict(sdD)
    DN.train()
    g = torch.Generator().manual_seed(200)
    x, labels, n = synthetic_jets(3, 100, g)
    x.requires_grad_(True)
    out = DN(x, labels)
    loss = ref_train.mse(out, torch.ones(3, 1))
    DN.zero_grad()
    loss.backward()
    cases["d_n100"] = dict(x=x.detach().clone(), labels=labels, out=out.detach(), loss=loss.detach(),
                           grads=grads_of(DN), dx=x.grad.clone())
    for N, B in ((100, 3), (150, 2)):
        GN = setup_training.setup_mpgan(mp_args(num_hits=N), gen=True)
        GN.load_state_dict(sdG)
        DN = setup_training.setup_mpgan(mp_args(num_hits=N, disc_dropout=0.0), gen=False)
        DN.load_state_dict(sdD)
        GN.train()
        DN.train()
        g = torch.Generator().manual_seed(300 + N)
        n = torch.randint(1, N + 1, (B,), generator=g)
        n[0] = N
        noise = torch.randn(B, N, 32, generator=g) * 0.2
        labels = (n.float() * torch.tensor(1.0 / N)).unsqueeze(1)
        GN.zero_grad()
        fake = GN(noise, labels)
        loss = ref_train.calc_G_loss("ls", DN(fake, labels))
        loss.backward()
        cases[f"g_through_d_n{N}"] = dict(noise=noise, labels=labels, fake=fake.detach(), loss=loss.detach(),
                                          grads=grads_of(GN))
    save("disc_fwd_bwd_large.pt", cases)

    # ------------------------------------------------------------------ option variants of the networks (default widths)
    cases = {}
    # lfc generator (model.py:601-606, 740-745): latent vector -> Linear -> [N, 32] node features
    torch.manual_seed(61)
    a = mp_args(lfc=True)
    GL = setup_training.setup_mpgan(a, gen=True)
    GL.train()
    g = torch.Generator().manual_seed(62)
    n = torch.tensor([30, 11, 2, 1])
    noise = torch.randn(4, 128, generator=g) * 0.2
    labels = (n.float() * torch.tensor(1.0 / 30)).unsqueeze(1)
    out = GL(noise, labels)
    w = torch.randn(out.shape, generator=g)
    (out * w).sum().backward()
    cases["lfc"] = dict(sd={k: v.clone() for k, v in GL.state_dict().items()}, noise=noise, labels=labels, w=w,
                        out=out.detach(), grads=grads_of(GL))
    # discriminator variants: dea=False (per-particle score, masked mean) and sum=False (mean aggregation in the
    # layers, masked mean pooling before fnd)
    for name, over in (("dea_false", dict(dea=False)), ("sum_false", dict(sum=False))):
        torch.manual_seed(63)
        DV = setup_training.setup_mpgan(mp_args(disc_dropout=0.0, **over), gen=False)
        DV.train()
        g = torch.Generator().manual_seed(64)
        x, labels, _ = synthetic_jets(5, 30, g)
        x.requires_grad_(True)
        out = DV(x, labels)
        loss = ref_train.mse(out, torch.ones(5, 1))
        loss.backward()
        cases[name] = dict(over=over, sd={k: v.clone() for k, v in DV.state_dict().items()}, x=x.detach().clone(),
                           labels=labels, out=out.detach(), loss=loss.detach(), grads=grads_of(DV), dx=x.grad.clone())
    save("net_variants.pt", cases)

    # ------------------------------------------------------------------ one full G+D step at N=100 (configs[4] shapes)
    torch.manual_seed(4)
    args = mp_args(disc_dropout=0.0, num_hits=100)
    G2 = setup_training.setup_mpgan(args, gen=True)
    G2.load_state_dict(sdG)
    D2 = setup_training.setup_mpgan(args, gen=False)
    D2.load_state_dict(sdD)
    args.spectral_norm_gen = False
    G_opt, D_opt = setup_training.optimizers(args, G2, D2)
    model_args = {"lfc": False, "lfc_latent_size": 128, "mask_learn_sep": False, "latent_node_size": 32}
    g = torch.Generator().manual_seed(71)
    data, labels, _ = synthetic_jets(4, 100, g)
    noise_d = torch.randn(4, 100, 32, generator=g) * 0.2
    noise_g = torch.randn(4, 100, 32, generator=g) * 0.2
    d_items = ref_train.train_D(model_args, D2, G2, D_opt, G_opt, data, "ls", labels=labels,
                                gen_args={"num_particles": 100, "noise": noise_d})
    gradsD = grads_of(D2)
    g_item = ref_train.train_G(model_args, D2, G2, G_opt, "ls", 4, labels=labels,
                               gen_args={"num_particles": 100, "noise": noise_g})
    gradsG = grads_of(G2)
    save("train_step_n100.pt", dict(data=data, labels=labels, noise_d=noise_d, noise_g=noise_g,
                                    loss_d=d_items["D"], loss_g=g_item, gradsD=gradsD, gradsG=gradsG,
                                    lr_d=args.lr_disc, lr_g=args.lr_gen,
                                    sdD_after={k: v.clone() for k, v in D2.state_dict().items()},
                                    sdG_after={k: v.clone() for k, v in G2.state_dict().items()}))

    # ------------------------------------------------------------------ WGAN-GP (train.py:286-324, 380-383)
    cases = {}
    for name, over in (("masked", dict()), ("unmasked", dict(mask_c=False))):
        # loss "w": no final sigmoid (setup_training.py disc_args); dropout 0 (torch's RNG stream cannot be matched)
        torch.manual_seed(81)
        DW = setup_training.setup_mpgan(mp_args(disc_dropout=0.0, loss="w", **over), gen=False)
        DW.train()
        g = torch.Generator().manual_seed(82)
        real, labels, _ = synthetic_jets(6, 30, g)
        fake, _, _ = synthetic_jets(6, 30, g)
        if name == "unmasked":
            real, fake = real[..., :3].contiguous(), fake[..., :3].contiguous()
        torch.manual_seed(83)
        alpha = torch.rand(6, 1, 1)            # the draw gradient_penalty makes first (train.py:289-293)
        torch.manual_seed(83)
        DW.zero_grad()
        gp = ref_train.gradient_penalty(10.0, DW, real, fake, 6, "cpu")
        gp.backward()
        gp_grads = grads_of(DW)
        # the penalised input gradient itself (first order, incl. the mask channel)
        xi = (alpha * real + (1 - alpha) * fake).requires_grad_(True)
        (gi,) = torch.autograd.grad(DW(xi).sum(), xi)
        # whole critic loss with the penalty (calc_D_loss 'w')
        DW.zero_grad()
        torch.manual_seed(83)
        ro, fo = DW(real.clone(), labels), DW(fake, labels)
        dl, items = ref_train.calc_D_loss("w", DW, real, fake, ro, fo, 6, gp_lambda=10.0)
        dl.backward()
        cases[name] = dict(over=over, sd={k: v.clone() for k, v in DW.state_dict().items()}, real=real, fake=fake,
                           labels=labels, alpha=alpha, gp=gp.detach(), gp_grads=gp_grads, dx_interp=gi,
                           d_loss=dl.detach(), d_loss_items=items, d_loss_grads=grads_of(DW))
    save("wgan_gp.pt", cases)

    # ------------------------------------------------------------------ MPLayer: kNN and conditioning columns
    cases = {}
    variants = {
        "knn_plain": dict(fully_connected=False, num_knn=3),
        "knn_posdiff": dict(fully_connected=False, num_knn=4, pos_diffs=True, all_ef=False, delta_r=True),
        "knn_allef_noself_mean": dict(fully_connected=False, num_knn=3, pos_diffs=True, all_ef=True, delta_r=False,
                                      self_loops=False, sum=False),
        "clabels2": dict(clabels=2),
        "mask_fne_np": dict(mask_fne_np=True),
        "clabels1_fne_posdiff": dict(clabels=1, mask_fne_np=True, pos_diffs=True, all_ef=False, delta_r=True),
    }
    for name, kw in variants.items():
        for masked in (True, False):
            torch.manual_seed(91)
            layer = MPLayer(5, [16, 24, 32], [40, 40], 6, **kw)
            g = torch.Generator().manual_seed(92)
            x = torch.randn(3, 9, 5, generator=g).requires_grad_(True)
            mask = (torch.rand(3, 9, 1, generator=g) > 0.3).float() if masked else None
            labels = torch.rand(3, 2, generator=g)
            njp = torch.rand(3, 1, generator=g)
            out = layer(x, masked, mask, labels, njp)
            w = torch.randn(out.shape, generator=g)
            (out * w).sum().backward()
            cases[f"{name}{'_masked' if masked else ''}"] = dict(
                kw=kw, sd={k: v.clone() for k, v in layer.state_dict().items()}, x=x.detach().clone(), mask=mask,
                labels=labels, njp=njp, w=w, out=out.detach(), grads=grads_of(layer), dx=x.grad.clone())
    # default widths, kNN (the reference's own N^2 escape hatch: num_knn 10/20 in its argument files)
    for name, kw, N in (("knn10_default", dict(fully_connected=False, num_knn=10), 30),
                        ("knn20_default_n150", dict(fully_connected=False, num_knn=20), 150)):
        torch.manual_seed(93)
        layer = MPLayer(32, [96, 160, 192], [256, 256], 32, **kw)
        g = torch.Generator().manual_seed(94)
        x = (torch.randn(2, N, 32, generator=g) * 0.3).requires_grad_(True)
        n = torch.tensor([N, N // 3])
        mask = (torch.arange(N)[None, :] < n[:, None]).float().unsqueeze(2)
        out = layer(x, True, mask)
        w = torch.randn(out.shape, generator=g)
        (out * w).sum().backward()
        cases[name] = dict(kw=kw, sd={k: v.clone() for k, v in layer.state_dict().items()}, x=x.detach().clone(),
                           mask=mask, labels=None, njp=None, w=w, out=out.detach(), grads=grads_of(layer),
                           dx=x.grad.clone())
    save("mplayer_variants2.pt", cases)

    # ------------------------------------------------------------------ GAPT with LayerNorm (gapt/model.py:116-118,130-136)
    cases = {}
    for isab in (False, True):
        common = dict(num_particles=30, num_heads=4, embed_dim=64, sab_fc_layers=[], use_mask=True,
                      use_isab=isab, num_isab_nodes=10)
        lin = dict(leaky_relu_alpha=0.2, dropout_p=0.0, batch_norm=False, spectral_norm=False)
        torch.manual_seed(101)
        GG = GAPT_G(sab_layers=2, output_feat_size=3, final_fc_layers=[], dropout_p=0.0,
                    layer_norm=True, linear_args=lin, **common)
        GD = GAPT_D(sab_layers=2, input_feat_size=3, final_fc_layers=[], dropout_p=0.0,
                    layer_norm=True, linear_args=lin, **common)
        # non-trivial LayerNorm affine parameters
        with torch.no_grad():
            for k, p in list(GG.named_parameters()) + list(GD.named_parameters()):
                if ".norm" in k:
                    p.add_(torch.randn(p.shape) * 0.2)
        g = torch.Generator().manual_seed(102)
        n = torch.tensor([30, 17, 5, 1])
        noise = (torch.randn(4, 30, 64, generator=g) * 0.2).requires_grad_(True)
        labels = (n.float() * torch.tensor(1.0 / 30)).unsqueeze(1)
        fake = GG(noise, labels)
        dout = GD(fake, labels)
        loss = ref_train.calc_G_loss("ls", dout)
        loss.backward()
        cases["isab" if isab else "sab"] = dict(
            sdG={k: v.clone() for k, v in GG.state_dict().items()},
            sdD={k: v.clone() for k, v in GD.state_dict().items()},
            noise=noise.detach().clone(), labels=labels, fake=fake.detach(), dout=dout.detach(),
            loss=loss.detach(), gradsG=grads_of(GG), gradsD=grads_of(GD), dnoise=noise.grad.clone())
    save("gapt_layernorm.pt", cases)


if __name__ == "__main__":
    which = sys.argv[1:] or ["base", "extra"]
    if "base" in which:
        base()
    if "extra" in which:
        extra()
