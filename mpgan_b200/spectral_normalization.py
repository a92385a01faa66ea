"""Spectral-norm wrapper with the reference's parameter layout
(``mpgan/spectral_normalization.py:12-64``): ``module.weight`` is replaced by ``weight_bar``,
``weight_u`` and ``weight_v`` (u, v without grad).  One power iteration per forward, in train and
eval alike, executed by one warp-shuffle kernel (``mpg_sn_fwd``).
"""
import torch
from torch import nn
from torch.nn import Parameter

from . import ops


def l2normalize(v, eps=1e-12):
    return v / (v.norm() + eps)


class SpectralNorm(nn.Module):
    def __init__(self, module, name="weight", power_iterations=1):
        super().__init__()
        self.module = module
        self.name = name
        if power_iterations != 1:
            raise NotImplementedError("the fused spectral-norm kernel runs exactly one power iteration")
        self.power_iterations = power_iterations
        if not self._made_params():
            self._make_params()

    def compute_weight(self):
        """Runs the power iteration (updates u, v in place) and returns W_bar / (sigma + 1e-12)."""
        u = getattr(self.module, self.name + "_u")
        v = getattr(self.module, self.name + "_v")
        w = getattr(self.module, self.name + "_bar")
        return ops.spectral_normalize(w, u.data, v.data)

    def _made_params(self):
        return all(hasattr(self.module, self.name + s) for s in ("_u", "_v", "_bar"))

    def _make_params(self):
        w = getattr(self.module, self.name)
        height = w.data.shape[0]
        width = w.view(height, -1).data.shape[1]
        u = Parameter(w.data.new(height).normal_(0, 1), requires_grad=False)
        v = Parameter(w.data.new(width).normal_(0, 1), requires_grad=False)
        u.data = l2normalize(u.data)
        v.data = l2normalize(v.data)
        w_bar = Parameter(w.data)
        del self.module._parameters[self.name]
        self.module.register_parameter(self.name + "_u", u)
        self.module.register_parameter(self.name + "_v", v)
        self.module.register_parameter(self.name + "_bar", w_bar)

    def forward(self, x):
        return ops.linear(x, self.compute_weight(), self.module.bias, False, 0.0, 0.0)
