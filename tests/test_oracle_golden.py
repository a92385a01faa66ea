"""Pins the oracle restatement (oracle/*.py) to vectors minted from the unmodified reference."""
import torch

from oracle import gapt_oracle as go
from oracle import mpgan_oracle as mo

G_CFG = mo.NetCfg(num_particles=30, final_activation="tanh")
D_CFG = mo.NetCfg(num_particles=30, final_activation="sigmoid",
                  layers=[mo.EdgeCfg(all_ef=False), mo.EdgeCfg()])


def close(a, b, tol=2e-5):
    scale = max(float(b.abs().max()), 1e-6)
    assert float((a - b).abs().max()) <= tol * scale, (float((a - b).abs().max()), scale)


def leafify(sd):
    return {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}


def test_survey_vector(golden):
    sd = golden("mp_g_weights.pt")
    c = golden("gen_forward.pt")["survey4"]
    out = mo.generator(sd, c["noise"], c["labels"], G_CFG)
    assert out.shape == (4, 30, 4)
    assert torch.equal(out[..., 3], c["out"][..., 3])  # mask channel bit-exact
    assert (out[..., 3] + 0.5).sum(1).tolist() == [30, 17, 5, 1]
    close(out, c["out"], 1e-5)
    assert abs(float(out.double().sum()) - (-49.17503)) < 1e-3


def test_generator_cases(golden):
    sd = golden("mp_g_weights.pt")
    cases = golden("gen_forward.pt")
    for name, N in (("b64", 30), ("n100", 100), ("n150", 150)):
        c = cases[name]
        cfg = mo.NetCfg(num_particles=N, final_activation="tanh")
        out = mo.generator(sd, c["noise"], c["labels"], cfg)
        assert torch.equal(out[..., 3], c["out"][..., 3])
        close(out, c["out"], 1e-5)


def test_discriminator_fwd_bwd(golden):
    cases = golden("disc_fwd_bwd.pt")
    for name, N in (("n30", 30), ("n150", 150)):
        c = cases[name]
        sd = leafify(golden("mp_d_seed4_weights.pt"))
        cfg = mo.NetCfg(num_particles=N, final_activation="sigmoid", layers=D_CFG.layers)
        x = c["x"].clone().requires_grad_(True)
        out = mo.discriminator(sd, x, c["labels"], cfg, training=True)
        close(out, c["out"])
        loss = mo.g_loss_ls(out)
        loss.backward()
        close(x.grad, c["dx"], 1e-4)
        for k, g in c["grads"].items():
            close(sd[k].grad, g, 1e-4)


def test_g_through_d(golden):
    c = golden("disc_fwd_bwd.pt")["g_through_d"]
    sdG = leafify(golden("mp_g_weights.pt"))
    sdD = golden("mp_d_seed4_weights.pt")
    fake = mo.generator(sdG, c["noise"], c["labels"], G_CFG, training=True)
    loss = mo.g_loss_ls(mo.discriminator(sdD, fake, c["labels"], D_CFG, training=True))
    close(loss, c["loss"])
    loss.backward()
    for k, g in c["grads"].items():
        close(sdG[k].grad, g, 1e-4)


def test_mplayer_variants(golden):
    for name, c in golden("mplayer_variants.pt").items():
        sd = leafify({"l." + k: v for k, v in c["sd"].items()})
        ec = mo.EdgeCfg(**c["kw"])
        x = c["x"].clone().requires_grad_(True)
        out = mo.mp_layer(x, sd, "l", ec, c["mask"])
        close(out, c["out"])
        (out * c["w"]).sum().backward()
        close(x.grad, c["dx"], 1e-4)
        for k, g in c["grads"].items():
            close(sd["l." + k].grad, g, 1e-4)


def test_rank_mask_bit_exact(golden):
    for name, c in golden("rank_mask.pt").items():
        N = c["x0"].shape[1]
        m = mo.rank_mask(c["x0"], c["labels"][:, -1], N)
        assert torch.equal(m.to(torch.uint8), c["mask"]), name
    # the truncation off-by-ones the survey probed (n/N convention)
    c = golden("rank_mask.pt")["n150_div"]
    short = [int(i) + 1 for i in torch.nonzero(c["count"] != torch.arange(1, 151)).flatten()]
    assert short == [49, 63, 89, 98, 117, 126]


def test_spectral_norm(golden):
    c = golden("spectral_norm.pt")
    sd = leafify(c["sd0"])
    for k in sd:
        if k.endswith("weight_u") or k.endswith("weight_v"):
            sd[k].requires_grad_(False)
    sn = {}
    x = c["x"].clone().requires_grad_(True)
    out = mo.linear_net(x, sd, "", True, sn_out=sn) if False else None
    # LinearNet state_dict keys are "net.i..." (no prefix): use an empty-prefix shim
    sd2 = {"p." + k: v for k, v in sd.items()}
    out = mo.linear_net(x, sd2, "p", True, sn_out=sn)
    close(out, c["out"])
    (out * c["w"]).sum().backward()
    close(x.grad, c["dx"], 1e-4)
    for k, g in c["grads"].items():
        close(sd2["p." + k].grad, g, 1e-4)
    for k, v in sn.items():
        close(v, c["sd1"][k[2:]])


def test_gapt(golden):
    for name, c in golden("gapt.pt").items():
        cfgG = go.GaptCfg(sab_layers=4, use_isab=(name == "isab"))
        cfgD = go.GaptCfg(sab_layers=2, use_isab=(name == "isab"))
        sdG, sdD = leafify(c["sdG"]), leafify(c["sdD"])
        noise = c["noise"].clone().requires_grad_(True)
        fake = go.gapt_g(sdG, noise, c["labels"], cfgG)
        assert torch.equal(fake[..., 3], c["fake"][..., 3])
        close(fake, c["fake"])
        dout = go.gapt_d(sdD, fake, c["labels"], cfgD)
        close(dout, c["dout"])
        mo.g_loss_ls(dout).backward()
        close(noise.grad, c["dnoise"], 1e-4)
        for k, g in c["gradsG"].items():
            close(sdG[k].grad, g, 1e-4)
        for v in sdD.values():
            v.grad = None
        x = c["x"].clone().requires_grad_(True)
        rout = go.gapt_d(sdD, x, c["xlabels"], cfgD)
        close(rout, c["rout"])
        mo.g_loss_ls(rout).backward()
        close(x.grad, c["dx"], 1e-4)
        for k, g in c["gradsD"].items():
            close(sdD[k].grad, g, 1e-4)


def test_train_step(golden):
    c = golden("train_step.pt")
    sdG, sdD = leafify(golden("mp_g_weights.pt")), leafify(golden("mp_d_seed4_weights.pt"))
    r = mo.gd_step(sdG, sdD, G_CFG, D_CFG, c["data"], c["labels"], c["noise_d"], c["noise_g"],
                   lr_d=c["lr_d"], lr_g=c["lr_g"])
    assert abs(r["loss_d"] - c["loss_d"]) < 1e-5 and abs(r["loss_g"] - c["loss_g"]) < 1e-5
    for k, g in c["gradsD"].items():
        close(r["grads_d"][k], g, 1e-4)
    for k, g in c["gradsG"].items():
        close(r["grads_g"][k], g, 2e-4)
    for k, v in c["sdD_after"].items():
        # RMSprop's first step is ~ lr*10*sign(g): compare where |g| is well away from 0
        g = c["gradsD"][k]
        ok = g.abs() > 1e-3 * g.abs().max()
        assert float((sdD[k].detach() - v)[ok].abs().max()) < 2e-6


# ---- round-2 fixtures --------------------------------------------------------------------------------
def test_large_n_gradients(golden):
    cases = golden("disc_fwd_bwd_large.pt")
    c = cases["d_n100"]
    sd = leafify(golden("mp_d_seed4_weights.pt"))
    cfg = mo.NetCfg(num_particles=100, final_activation="sigmoid", layers=D_CFG.layers)
    x = c["x"].clone().requires_grad_(True)
    out = mo.discriminator(sd, x, c["labels"], cfg, training=True)
    close(out, c["out"])
    mo.g_loss_ls(out).backward()
    close(x.grad, c["dx"], 1e-4)
    for k, g in c["grads"].items():
        close(sd[k].grad, g, 1e-4)
    for N in (100, 150):
        c = cases[f"g_through_d_n{N}"]
        sdG, sdD = leafify(golden("mp_g_weights.pt")), golden("mp_d_seed4_weights.pt")
        cg = mo.NetCfg(num_particles=N, final_activation="tanh")
        cd = mo.NetCfg(num_particles=N, final_activation="sigmoid", layers=D_CFG.layers)
        fake = mo.generator(sdG, c["noise"], c["labels"], cg, training=True)
        close(fake, c["fake"], 1e-5)
        loss = mo.g_loss_ls(mo.discriminator(sdD, fake, c["labels"], cd, training=True))
        close(loss, c["loss"])
        loss.backward()
        for k, g in c["grads"].items():
            close(sdG[k].grad, g, 2e-4)


def test_net_variants(golden):
    cases = golden("net_variants.pt")
    c = cases["lfc"]
    sd = leafify(c["sd"])
    out = mo.generator(sd, c["noise"], c["labels"], mo.NetCfg(final_activation="tanh", lfc=True), training=True)
    assert torch.equal(out[..., 3], c["out"][..., 3])
    close(out, c["out"], 1e-5)
    (out * c["w"]).sum().backward()
    for k, g in c["grads"].items():
        close(sd[k].grad, g, 1e-4)
    for name in ("dea_false", "sum_false"):
        c = cases[name]
        sd = leafify(c["sd"])
        s = c["over"].get("sum", True)
        cfg = mo.NetCfg(final_activation="sigmoid", dea=c["over"].get("dea", True), dea_sum=s,
                        layers=[mo.EdgeCfg(all_ef=False, sum=s), mo.EdgeCfg(sum=s)])
        x = c["x"].clone().requires_grad_(True)
        out = mo.discriminator(sd, x, c["labels"], cfg, training=True)
        close(out, c["out"])
        mo.g_loss_ls(out).backward()
        close(x.grad, c["dx"], 1e-4)
        for k, g in c["grads"].items():
            close(sd[k].grad, g, 1e-4)


def test_train_step_n100(golden):
    c = golden("train_step_n100.pt")
    sdG, sdD = leafify(golden("mp_g_weights.pt")), leafify(golden("mp_d_seed4_weights.pt"))
    cg = mo.NetCfg(num_particles=100, final_activation="tanh")
    cd = mo.NetCfg(num_particles=100, final_activation="sigmoid", layers=D_CFG.layers)
    r = mo.gd_step(sdG, sdD, cg, cd, c["data"], c["labels"], c["noise_d"], c["noise_g"], lr_d=c["lr_d"], lr_g=c["lr_g"])
    assert abs(r["loss_d"] - c["loss_d"]) < 1e-5 and abs(r["loss_g"] - c["loss_g"]) < 1e-5
    for k, g in c["gradsD"].items():
        close(r["grads_d"][k], g, 1e-4)
    for k, g in c["gradsG"].items():
        close(r["grads_g"][k], g, 3e-4)


def test_wgan_gp(golden):
    for name, c in golden("wgan_gp.pt").items():
        sd = leafify(c["sd"])
        cfg = mo.NetCfg(final_activation="", mask_c=c["over"].get("mask_c", True), layers=D_CFG.layers)
        gp = mo.gradient_penalty(sd, cfg, c["real"], c["fake"], c["alpha"], 10.0)
        close(gp, c["gp"], 1e-4)
        gp.backward()
        for k, g in c["gp_grads"].items():
            close(sd[k].grad, g, 2e-4)
        # first-order input gradient, mask channel included
        xi = (c["alpha"] * c["real"] + (1 - c["alpha"]) * c["fake"]).requires_grad_(True)
        (gi,) = torch.autograd.grad(mo.discriminator(sd, xi, None, cfg, training=True).sum(), xi)
        close(gi, c["dx_interp"], 1e-4)
        for v in sd.values():
            v.grad = None
        dl = mo.d_loss_w(mo.discriminator(sd, c["real"].clone(), c["labels"], cfg, training=True),
                         mo.discriminator(sd, c["fake"], c["labels"], cfg, training=True)) + \
            mo.gradient_penalty(sd, cfg, c["real"], c["fake"], c["alpha"], 10.0)
        close(dl, c["d_loss"], 1e-4)
        dl.backward()
        for k, g in c["d_loss_grads"].items():
            close(sd[k].grad, g, 2e-4)


def test_mplayer_variants2(golden):
    """kNN message passing and the conditioning columns (clabels / mask_fne_np)."""
    for name, c in golden("mplayer_variants2.pt").items():
        sd = leafify({"l." + k: v for k, v in c["sd"].items()})
        ec = mo.EdgeCfg(**c["kw"])
        x = c["x"].clone().requires_grad_(True)
        out = mo.mp_layer(x, sd, "l", ec, c["mask"], c["labels"], c["njp"])
        close(out, c["out"], 5e-5)
        (out * c["w"]).sum().backward()
        close(x.grad, c["dx"], 2e-4)
        for k, g in c["grads"].items():
            close(sd["l." + k].grad, g, 2e-4)


def test_gapt_layernorm(golden):
    for name, c in golden("gapt_layernorm.pt").items():
        cfg = go.GaptCfg(sab_layers=2, use_isab=(name == "isab"), layer_norm=True)
        sdG, sdD = leafify(c["sdG"]), leafify(c["sdD"])
        noise = c["noise"].clone().requires_grad_(True)
        fake = go.gapt_g(sdG, noise, c["labels"], cfg)
        close(fake, c["fake"])
        dout = go.gapt_d(sdD, fake, c["labels"], cfg)
        close(dout, c["dout"])
        mo.g_loss_ls(dout).backward()
        close(noise.grad, c["dnoise"], 1e-4)
        for k, g in c["gradsG"].items():
            close(sdG[k].grad, g, 1e-4)
        for k, g in c["gradsD"].items():
            close(sdD[k].grad, g, 1e-4)
