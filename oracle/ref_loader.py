"""Imports the UNMODIFIED reference copied under ``oracle/_ref`` (see build_ref.py) and wires its own step functions
for the baseline arms of bench.py.  TEST / BASELINE INFRASTRUCTURE ONLY: nothing under ``mpgan_b200/`` imports this.

The reference's top-level module names (``mpgan``, ``gapt``, ``train``, ``setup_training``) are imported from
``oracle/_ref``; ``jetnet`` / ``matplotlib`` / ``mplhep`` (data set, metrics and plotting packages the training step
never calls) are replaced by inert stubs, exactly as ``oracle/make_golden.py`` does.
"""
from __future__ import annotations

import os
import sys
from unittest.mock import MagicMock

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_STUBS = ("jetnet", "jetnet.datasets", "jetnet.evaluation", "jetnet.datasets.normalisations", "jetnet.utils",
          "matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.cm", "matplotlib.lines", "mplhep")

_mods = None


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "train.py")) and os.path.isdir(os.path.join(REF_DIR, "mpgan"))


class Obj:
    def __init__(self, d):
        self.__dict__ = dict(d)


def load():
    """Returns a namespace with the reference modules: .train, .setup_training, .mpgan, .gapt."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError(f"{REF_DIR} not found: run `python oracle/build_ref.py` where /root/reference exists")
    for m in _STUBS:
        sys.modules.setdefault(m, MagicMock())
    if "mpgan" in sys.modules and not getattr(sys.modules["mpgan"], "__file__", "").startswith(REF_DIR):
        raise RuntimeError("a different `mpgan` module is already imported")
    sys.path.insert(0, REF_DIR)
    try:
        import gapt
        import mpgan
        import setup_training
        import train
    finally:
        sys.path.remove(REF_DIR)
    _mods = Obj(dict(train=train, setup_training=setup_training, mpgan=mpgan, gapt=gapt))
    return _mods


def mp_args(**over):
    """Argument namespace of the published mp_g model (trained_models/mp_g/args.txt) with overrides."""
    a = eval(open(os.path.join(REF_DIR, "trained_models", "mp_g", "args.txt")).read())
    a.update(device="cpu", load_model=False, multi_gpu=False)
    a.update(over)
    return Obj(a)


def gapt_models(N, isab, device, disc_dropout=0.5):
    """GAPT_G / GAPT_D wired as setup_training.setup_gapt (:1296-1347) does from its argparse defaults (:551-617)."""
    m = load()
    common = dict(num_particles=N, num_heads=4, embed_dim=64, sab_fc_layers=[], use_mask=True, use_isab=isab,
                  num_isab_nodes=10)
    lin = lambda p: dict(leaky_relu_alpha=0.2, dropout_p=p, batch_norm=False, spectral_norm=False)
    G = m.gapt.GAPT_G(sab_layers=4, output_feat_size=3, final_fc_layers=[], dropout_p=0.0, layer_norm=False,
                      linear_args=lin(0.0), **common).to(device)
    D = m.gapt.GAPT_D(sab_layers=2, input_feat_size=3, final_fc_layers=[], dropout_p=disc_dropout, layer_norm=False,
                      linear_args=lin(disc_dropout), **common).to(device)
    return G, D


class RefStep:
    """One G+D training step through the reference's OWN train.train_D / train.train_G (train.py:398-523) with its
    own modules and torch.optim.RMSprop (setup_training.optimizers, :1511-1513)."""

    def __init__(self, N, device="cpu", model="mpgan", isab=False, weights=None):
        import torch
        m = load()
        self.m, self.N, self.device, self.model = m, N, device, model
        if model == "gapt":
            self.G, self.D = gapt_models(N, isab, device)
            args = mp_args(num_hits=N, lr_gen=1.5e-4, lr_disc=0.5e-4)
            self.latent = 64
        else:
            args = mp_args(num_hits=N)
            self.G = m.setup_training.setup_mpgan(args, gen=True).to(device)
            self.D = m.setup_training.setup_mpgan(args, gen=False).to(device)
            if weights is not None:
                self.G.load_state_dict(weights[0])
                self.D.load_state_dict(weights[1])
            self.latent = 32
        args.spectral_norm_gen = False
        self.G_opt, self.D_opt = m.setup_training.optimizers(args, self.G, self.D)
        self.model_args = {"lfc": False, "lfc_latent_size": 128, "mask_learn_sep": False, "latent_node_size": self.latent}
        self.torch = torch

    def step(self, data, labels):
        t, m, B = self.torch, self.m, data.shape[0]
        nd = t.randn(B, self.N, self.latent, device=self.device) * 0.2
        ng = t.randn(B, self.N, self.latent, device=self.device) * 0.2
        d = m.train.train_D(self.model_args, self.D, self.G, self.D_opt, self.G_opt, data, "ls", labels=labels,
                            model=self.model, gen_args={"num_particles": self.N, "noise": nd})
        g = m.train.train_G(self.model_args, self.D, self.G, self.G_opt, "ls", B, labels=labels,
                            model=self.model, gen_args={"num_particles": self.N, "noise": ng})
        return d, g

    def generate(self, labels):
        t = self.torch
        self.G.eval()
        with t.no_grad():
            noise = t.randn(labels.shape[0], self.N, self.latent, device=self.device) * 0.2
            return self.G(noise, labels)
