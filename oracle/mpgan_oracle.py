"""fp32 CPU restatement of the MPGAN message-passing path (TEST INFRASTRUCTURE ONLY).

Functional over a reference-layout ``state_dict``; every function cites the reference lines it
restates (paths relative to the reference repo root).  Autograd through these functions is the
gradient oracle.  Pinned by ``tests/test_oracle_golden.py`` against vectors minted from the
unmodified reference by ``oracle/make_golden.py``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------------
@dataclass
class EdgeCfg:
    """Options of one MPLayer (mpgan/model.py:129-148)."""

    pos_diffs: bool = False
    all_ef: bool = True
    coords: str = "polarrel"
    delta_coords: bool = False
    delta_r: bool = True
    clabels: int = 0
    mask_fne_np: bool = False
    sum: bool = True
    fully_connected: bool = True
    num_knn: int = 20
    self_loops: bool = True


@dataclass
class NetCfg:
    """Options of an MPNet / MPGenerator / MPDiscriminator (mpgan/model.py:420-436,593,785-792)."""

    num_particles: int = 30
    mp_iters: int = 2
    alpha: float = 0.2
    dropout_p: float = 0.0
    spectral_norm: bool = False
    final_activation: str = ""
    mask_c: bool = True
    dea: bool = True
    dea_sum: bool = True
    lfc: bool = False
    layers: List[EdgeCfg] = field(default_factory=list)

    def layer(self, i: int) -> EdgeCfg:
        return self.layers[i] if self.layers else EdgeCfg()


# --------------------------------------------------------------------------------------------
# spectral norm  (mpgan/spectral_normalization.py:8-33)
# --------------------------------------------------------------------------------------------
def l2normalize(v: Tensor, eps: float = 1e-12) -> Tensor:
    return v / (v.norm() + eps)  # :8-9


def spectral_weight(w_bar: Tensor, u: Tensor, v: Tensor):
    """One power iteration; returns (W, u', v').  u', v' carry no grad; sigma does (:21-33)."""
    with torch.no_grad():
        v_new = l2normalize(torch.mv(w_bar.t(), u))  # :28
        u_new = l2normalize(torch.mv(w_bar, v_new))  # :29
    sigma = u_new.dot(w_bar.mv(v_new))  # :32
    return w_bar / (sigma + 1e-12), u_new, v_new  # :33


# --------------------------------------------------------------------------------------------
# LinearNet  (mpgan/model.py:70-85)
# --------------------------------------------------------------------------------------------
def _n_layers(sd: Dict[str, Tensor], prefix: str) -> int:
    n = 0
    while f"{prefix}.net.{n}.weight" in sd or f"{prefix}.net.{n}.module.weight_bar" in sd:
        n += 1
    return n


def linear_net(
    x: Tensor,
    sd: Dict[str, Tensor],
    prefix: str,
    final_linear: bool,
    alpha: float = 0.2,
    dropout_p: float = 0.0,
    training: bool = False,
    sn_out: Optional[dict] = None,
) -> Tensor:
    """Linear -> leaky_relu (skipped on the last layer iff final_linear) -> Dropout ALWAYS (:77-83)."""
    n = _n_layers(sd, prefix)
    for i in range(n):
        if f"{prefix}.net.{i}.weight" in sd:
            w, b = sd[f"{prefix}.net.{i}.weight"], sd[f"{prefix}.net.{i}.bias"]
        else:  # SpectralNorm wrapper (mpgan/model.py:65-68, spectral_normalization.py:44-60)
            p = f"{prefix}.net.{i}.module"
            w, u2, v2 = spectral_weight(sd[p + ".weight_bar"], sd[p + ".weight_u"], sd[p + ".weight_v"])
            if sn_out is not None:
                sn_out[p + ".weight_u"], sn_out[p + ".weight_v"] = u2, v2
            b = sd[p + ".bias"]
        x = F.linear(x, w, b)
        if i != n - 1 or not final_linear:
            x = F.leaky_relu(x, negative_slope=alpha)
        x = F.dropout(x, p=dropout_p, training=training)
    return x


# --------------------------------------------------------------------------------------------
# MPLayer, fully connected  (mpgan/model.py:206-317)
# --------------------------------------------------------------------------------------------
def num_edge_features(ec: EdgeCfg) -> int:
    """mpgan/model.py:173-181."""
    n = 0
    if ec.pos_diffs:
        if ec.delta_coords:
            n += 3 if ec.coords == "cartesian" else 2
        if ec.delta_r or ec.all_ef:
            n += 1
    return n


def pair_tensor(x: Tensor, ec: EdgeCfg) -> Tensor:
    """A[b, i*N+j] = (x_i | x_j | [diffs] | [dist]) : receiver i, sender j (:284-317)."""
    B, N, Fdim = x.shape
    x1 = x.unsqueeze(2).expand(B, N, N, Fdim)  # receiver (:294)
    x2 = x.unsqueeze(1).expand(B, N, N, Fdim)  # sender   (:295)
    parts = [x1, x2]
    if ec.pos_diffs:
        nc = 3 if ec.coords == "cartesian" else 2
        diffs = (x2 - x1) if ec.all_ef else (x2[..., :nc] - x1[..., :nc])  # :299-302
        dists = torch.norm(diffs + 1e-12, dim=3, keepdim=True)  # eps per component (:304)
        if ec.delta_r and ec.delta_coords:
            parts += [diffs, dists]
        elif ec.delta_r or ec.all_ef:
            parts += [dists]
        elif ec.delta_coords:
            parts += [diffs]
    return torch.cat(parts, dim=3).reshape(B * N * N, -1)


def knn_pair_tensor(x: Tensor, ec: EdgeCfg, mask: Optional[Tensor]):
    """k-nearest-neighbour edge inputs (:319-381): A[b, i*k+m] = (x_i | x_nbr(i,m) | [dist]), plus the neighbours'
    masks.  Distances are taken to the senders scaled by 1e4 where masked (so padded particles are never picked
    before real ones), over all features unless pos_diffs without all_ef (then the coordinates only); the distance
    fed to the edge network is the (differentiable) sorted value itself."""
    B, N, Fdim = x.shape
    k = ec.num_knn
    x1 = x.unsqueeze(2).expand(B, N, N, Fdim)
    x2s = x if mask is None else ((1 - 1e4) * mask + 1e4) * x       # :336-340
    x2 = x2s.unsqueeze(1).expand(B, N, N, Fdim)
    nc = 3 if ec.coords == "cartesian" else 2
    diffs = (x2 - x1) if (ec.all_ef or not ec.pos_diffs) else (x2[..., :nc] - x1[..., :nc])   # :343-346
    dists = torch.norm(diffs + 1e-12, dim=3)                         # [B, N, N]  (:348)
    srt = torch.sort(dists, dim=2)                                   # :351
    s0 = int(ec.self_loops is False)                                 # :354
    d_k = srt[0][:, :, s0:k + s0].reshape(B, N * k, 1)               # :358-360
    idx = srt[1][:, :, s0:k + s0].reshape(B, N * k, 1)               # :361-363
    x1_knn = x.unsqueeze(2).expand(B, N, k, Fdim).reshape(B, N * k, Fdim)   # :366
    A_mask = None
    if mask is not None:                                             # :369-374
        g = torch.gather(torch.cat((x, mask), dim=2), 1, idx.expand(-1, -1, Fdim + 1))
        A_mask, x2_knn = g[:, :, -1:], g[:, :, :-1]
    else:
        x2_knn = torch.gather(x, 1, idx.expand(-1, -1, Fdim))        # :376
    A = torch.cat((x1_knn, x2_knn, d_k), dim=2) if ec.pos_diffs else torch.cat((x1_knn, x2_knn), dim=2)   # :380-383
    return A.reshape(B * N * k, -1), A_mask


def mp_layer(
    x: Tensor,
    sd: Dict[str, Tensor],
    prefix: str,
    ec: EdgeCfg,
    mask: Optional[Tensor] = None,
    labels: Optional[Tensor] = None,
    num_jet_particles: Optional[Tensor] = None,
    alpha: float = 0.2,
    dropout_p: float = 0.0,
    training: bool = False,
    sn_out: Optional[dict] = None,
) -> Tensor:
    B, N, _ = x.shape
    if ec.fully_connected:  # :241-246
        A, A_mask, K = pair_tensor(x, ec), None, N
    else:
        A, A_mask = knn_pair_tensor(x, ec, mask)
        K = ec.num_knn
    if ec.clabels:  # :247-249  (row r of jet b gets labels[r % B] -- reference quirk of .repeat)
        A = torch.cat((A, labels[:, : ec.clabels].repeat(N * K, 1)), dim=1)
    if ec.mask_fne_np:  # :251-253
        A = torch.cat((A, num_jet_particles.repeat(N * K, 1)), dim=1)
    A = linear_net(A, sd, prefix + ".fe", False, alpha, dropout_p, training, sn_out)  # :256
    A = A.view(B, N, K, -1)
    if mask is not None:
        if ec.fully_connected:
            A = A * mask.unsqueeze(1)  # sender axis (:262)
        else:
            A = A * A_mask.view(B, N, K, 1)  # the gathered neighbours' masks (:264)
    A = A.sum(2) if ec.sum else A.mean(2)  # mean divides by N (:267)
    h = torch.cat((A, x), 2).reshape(B * N, -1)  # :268
    if ec.clabels:
        h = torch.cat((h, labels[:, : ec.clabels].repeat(N, 1)), dim=1)  # :270-272
    if ec.mask_fne_np:
        h = torch.cat((h, num_jet_particles.repeat(N, 1)), dim=1)  # :274-276
    h = linear_net(h, sd, prefix + ".fn", True, alpha, dropout_p, training, sn_out)  # :279
    return h.view(B, N, -1)


# --------------------------------------------------------------------------------------------
# generator / discriminator  (mpgan/model.py:498-523, 689-699, 723-752, 810-831, 881-884)
# --------------------------------------------------------------------------------------------
def rank_mask(x0: Tensor, labels_last: Tensor, num_particles: int) -> Tensor:
    """mask_c: n = int(fp32(label)*N) - 1 (truncation), mask = rank(x[:,:,0]) <= n  (:692-699)."""
    n = (labels_last * num_particles).int() - 1
    return (x0.argsort(1).argsort(1) <= n.unsqueeze(1)).unsqueeze(2).float()


def _run_layers(x, sd, cfg, mask, labels, training, sn_out, num_jet_particles=None):
    for i in range(cfg.mp_iters):
        x = mp_layer(
            x, sd, f"mp_layers.{i}", cfg.layer(i), mask, labels, num_jet_particles,
            cfg.alpha, cfg.dropout_p, training, sn_out,
        )
    return x


def generator(sd, noise: Tensor, labels: Optional[Tensor], cfg: NetCfg, training=False, sn_out=None):
    x = noise
    if cfg.lfc:  # :601-606
        x = F.linear(x, sd["lfc_layer.weight"], sd["lfc_layer.bias"]).reshape(
            x.shape[0], cfg.num_particles, -1
        )
    mask = rank_mask(x[:, :, 0], labels[:, -1], cfg.num_particles) if cfg.mask_c else None
    x = _run_layers(x, sd, cfg, mask, labels, training, sn_out)
    if cfg.final_activation == "tanh":  # :535-536
        x = torch.tanh(x)
    elif cfg.final_activation == "sigmoid":
        x = torch.sigmoid(x)
    return torch.cat((x, mask - 0.5), dim=2) if mask is not None else x  # :752


def discriminator(sd, x: Tensor, labels: Optional[Tensor], cfg: NetCfg, training=False, sn_out=None):
    mask = None
    if cfg.mask_c:
        mask = x[:, :, -1:] + 0.5  # real-valued multiplier (:881)
        x = x[:, :, :-1]  # :884
    njp = mask.mean(1) if (mask is not None and any(l.mask_fne_np for l in cfg.layers)) else None  # :886-887
    x = _run_layers(x, sd, cfg, mask, labels, training, sn_out, njp)
    do_mean = not (cfg.dea and cfg.dea_sum)  # :811-813
    if mask is not None:
        x = (x * mask).sum(1)  # :816-817
        if do_mean:
            x = x / (mask.sum(1) + 1e-12)  # :820
    else:
        x = x.mean(1) if do_mean else x.sum(1)  # :822
    if cfg.dea:  # :825-829 (fnd LinearNet, final_linear=True, dropout still applied)
        x = linear_net(x, sd, "fnd_layer", True, cfg.alpha, cfg.dropout_p, training, sn_out)
    if cfg.final_activation == "sigmoid":  # :537-538
        x = torch.sigmoid(x)
    elif cfg.final_activation == "tanh":
        x = torch.tanh(x)
    return x


# --------------------------------------------------------------------------------------------
# losses + one G+D step  (train.py:331-395, 398-523; optimizer setup_training.py:1511-1513)
# --------------------------------------------------------------------------------------------
def d_loss_ls(real_out: Tensor, fake_out: Tensor) -> Tensor:
    return F.mse_loss(real_out, torch.ones_like(real_out)) + F.mse_loss(
        fake_out, torch.zeros_like(fake_out)
    )  # train.py:357-358,369-370,378


def g_loss_ls(fake_out: Tensor) -> Tensor:
    return F.mse_loss(fake_out, torch.ones_like(fake_out))  # train.py:467,472


def gradient_penalty(sdD, cfgD: NetCfg, real: Tensor, fake: Tensor, alpha: Tensor, gp_lambda: float,
                     training: bool = True) -> Tensor:
    """WGAN-GP term (train.py:286-324): D at x = alpha*real + (1-alpha)*fake, the gradient of sum(D(x)) w.r.t. x with
    the graph kept, gp = lambda * mean((sqrt(sum_jet grad^2 + 1e-12) - 1)^2).  ``alpha`` [B,1,1] is the uniform draw
    the reference makes first (:289-293); D is called WITHOUT labels (:303)."""
    x = (alpha * real + (1 - alpha) * fake).detach().requires_grad_(True)
    out = discriminator(sdD, x, None, cfgD, training=training)
    (g,) = torch.autograd.grad(out, x, grad_outputs=torch.ones_like(out), create_graph=True, retain_graph=True)
    gn = torch.sqrt(torch.sum(g.reshape(x.shape[0], -1) ** 2, dim=1) + 1e-12)
    return gp_lambda * ((gn - 1) ** 2).mean()


def d_loss_w(real_out: Tensor, fake_out: Tensor) -> Tensor:
    return -real_out.mean() + fake_out.mean()  # train.py:371-373


def rmsprop_step(p: Tensor, g: Tensor, sq: Tensor, lr: float, alpha=0.99, eps=1e-8):
    """torch.optim.RMSprop defaults (no momentum, not centered)."""
    sq.mul_(alpha).addcmul_(g, g, value=1 - alpha)
    p.addcdiv_(g, sq.sqrt().add_(eps), value=-lr)


def synthetic_jets(B: int, N: int, gen: torch.Generator, all_real: bool = False):
    """features U(-.5,.5) zeroed on padded rows, 4th channel mask-0.5; labels n*(1/N)."""
    n = torch.full((B,), N) if all_real else torch.randint(1, N + 1, (B,), generator=gen)
    real = (torch.arange(N)[None, :] < n[:, None]).float().unsqueeze(2)
    feats = (torch.rand(B, N, 3, generator=gen) - 0.5) * real
    x = torch.cat((feats, real - 0.5), dim=2)
    labels = (n.float() * torch.tensor(1.0 / N, dtype=torch.float32)).unsqueeze(1)
    return x, labels, n


def gd_step(sdG, sdD, cfgG: NetCfg, cfgD: NetCfg, data, labels, noise_d, noise_g,
            lr_d=3e-5, lr_g=1e-5, stateD=None, stateG=None, d_training=True):
    """One train_D then train_G exactly as train.py:398-462, 479-523 (loss 'ls', gp 0, RMSprop).

    ``sdG``/``sdD`` are dicts of leaf tensors (requires_grad) updated in place.  Returns the
    loss values and the gradients each optimizer consumed.  train_D back-propagates into G as
    well (no detach, train.py:428-437) -- those G grads are discarded by G's zero_grad (:495).
    """
    stateD = {} if stateD is None else stateD
    stateG = {} if stateG is None else stateG
    pD = [k for k, v in sdD.items() if v.requires_grad]
    pG = [k for k, v in sdG.items() if v.requires_grad]
    # ---- train_D: D.train(), G.eval() (:419-421)
    real_out = discriminator(sdD, data.clone(), labels, cfgD, training=d_training)  # :425
    fake = generator(sdG, noise_d, labels, cfgG, training=False)  # :428-437
    fake_out = discriminator(sdD, fake, labels, cfgD, training=d_training)  # :446
    loss_d = d_loss_ls(real_out, fake_out)
    gD = torch.autograd.grad(loss_d, [sdD[k] for k in pD])
    with torch.no_grad():
        for k, g in zip(pD, gD):
            rmsprop_step(sdD[k], g, stateD.setdefault(k, torch.zeros_like(g)), lr_d)
    # ---- train_G: G.train(); D stays in train mode (:494)
    fake = generator(sdG, noise_g, labels, cfgG, training=True)  # :499-507
    fake_out = discriminator(sdD, fake, labels, cfgD, training=d_training)  # :513
    loss_g = g_loss_ls(fake_out)
    gG = torch.autograd.grad(loss_g, [sdG[k] for k in pG])
    with torch.no_grad():
        for k, g in zip(pG, gG):
            rmsprop_step(sdG[k], g, stateG.setdefault(k, torch.zeros_like(g)), lr_g)
    return {
        "loss_d": float(loss_d), "loss_g": float(loss_g),
        "grads_d": dict(zip(pD, gD)), "grads_g": dict(zip(pG, gG)),
    }
