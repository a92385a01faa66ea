"""Drop-in GAPT modules (reference ``gapt/model.py``: ``MAB`` :93-139, ``SAB`` :143-154, ``PMA``
:158-174, ``ISAB`` :178-191, ``GAPT_G`` :205-274, ``GAPT_D`` :277-344) on the sm_100a kernels.

Parameter names match the reference (``attention.in_proj_weight``, ``attention.out_proj.weight``,
``ff.net.i.*``, ``I``, ``S``...) so reference checkpoints load with ``strict=True``.  The attention
mask is never expanded to ``[B*heads, Nq, Nk]``: the kernel takes the per-key mask directly, which
is all SAB / ISAB / PMA ever build (:127,152,172,189).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
from torch import Tensor

from . import ops
from .model import LinearNet


class MAB(nn.Module):
    fused = True    # class-level switch (tests compare the fused kernel with the per-op path)

    def __init__(self, embed_dim: int, num_heads: int, ff_layers: list = [], layer_norm: bool = False,
                 dropout_p: float = 0.0, final_linear: bool = True, linear_args={}):
        super().__init__()
        self.num_heads = num_heads
        self.embed_dim = embed_dim
        # parameter container only (in_proj_weight [3E,E], in_proj_bias, out_proj.{weight,bias});
        # its forward is never called
        self.attention = nn.MultiheadAttention(embed_dim, num_heads, batch_first=True)
        self.ff = LinearNet(ff_layers, input_size=embed_dim, output_size=embed_dim, final_linear=final_linear,
                            **linear_args)
        self.layer_norm = layer_norm
        if layer_norm:   # reference :116-118 (same module / parameter names: norm1.weight, norm1.bias, ...)
            self.norm1 = nn.LayerNorm(embed_dim)
            self.norm2 = nn.LayerNorm(embed_dim)
        self.dropout_p = float(dropout_p)
        self.dropout = nn.Dropout(p=dropout_p)

    def forward(self, x: Tensor, y: Tensor, y_mask: Tensor = None):
        """``y_mask``: per-key mask [B, Nk] / [B, Nk, 1] in JetNet convention (1 = real); keys whose
        mask is not exactly 1 are ignored.  (The reference takes the expanded boolean ignore-mask.)"""
        E = self.embed_dim
        att = self.attention
        w, b = att.in_proj_weight, att.in_proj_bias
        if self.fused and not self.layer_norm and len(self.ff.net) == 1 and not self.ff.final_linear and x.dim() == 3 \
                and ops.mab_supported(E, self.num_heads, x.shape[1], y.shape[1]):
            # the whole block as ONE kernel per direction (ops.MabFn)
            wf, bf = self.ff.layer_params(0)
            tr = self.training
            return ops.mab(x, None if x is y else y, y_mask, w, b, att.out_proj.weight, att.out_proj.bias, wf, bf,
                           self.num_heads, self.ff.leaky_relu_alpha, self.dropout_p if tr else 0.0,
                           self.ff.dropout_p if tr else 0.0)
        if x is y:
            qkv = ops.linear(x, w, b, False, 0.0, 0.0)
            q, k, v = qkv[..., :E], qkv[..., E:2 * E], qkv[..., 2 * E:]
        else:
            q = ops.linear(x, w[:E], b[:E], False, 0.0, 0.0)
            kv = ops.linear(y, w[E:], b[E:], False, 0.0, 0.0)
            k, v = kv[..., :E], kv[..., E:]
        o = ops.attention(q, k, v, y_mask, self.num_heads)
        a = ops.linear(o, att.out_proj.weight, att.out_proj.bias, False, 0.0, 0.0)
        p = self.dropout_p if self.training else 0.0
        if self.layer_norm:   # x + attn -> LayerNorm -> Dropout -> + ff -> LayerNorm -> Dropout (:129-137)
            h = ops.layer_norm(ops.residual_dropout(x, a, 0.0), self.norm1.weight, self.norm1.bias, self.norm1.eps)
            h = ops.residual_dropout(h, None, p, rng_stream=48) if p > 0 else h
            f = self.ff(h)
            o = ops.layer_norm(ops.residual_dropout(h, f, 0.0), self.norm2.weight, self.norm2.bias, self.norm2.eps)
            return ops.residual_dropout(o, None, p, rng_stream=49) if p > 0 else o
        h = ops.residual_dropout(x, a, p, rng_stream=48)
        f = self.ff(h)
        return ops.residual_dropout(h, f, p, rng_stream=49)


class SAB(nn.Module):
    def __init__(self, **mab_args):
        super().__init__()
        self.mab = MAB(**mab_args)

    def forward(self, x: Tensor, mask: Tensor = None):
        return self.mab(x, x, mask)


class PMA(nn.Module):
    def __init__(self, embed_dim: int, num_seeds: int, **mab_args):
        super().__init__()
        self.S = nn.Parameter(torch.Tensor(1, num_seeds, embed_dim))
        nn.init.xavier_uniform_(self.S)
        self.mab = MAB(embed_dim, **mab_args)

    def forward(self, x: Tensor, mask: Tensor = None):
        return self.mab(self.S.expand(x.size(0), -1, -1).contiguous(), x, mask)


class ISAB(nn.Module):
    def __init__(self, num_inds, embed_dim, **mab_args):
        super().__init__()
        self.I = nn.Parameter(torch.Tensor(1, num_inds, embed_dim))
        self.num_inds = num_inds
        nn.init.xavier_uniform_(self.I)
        self.mab0 = MAB(embed_dim=embed_dim, **mab_args)
        self.mab1 = MAB(embed_dim=embed_dim, **mab_args)

    def forward(self, X, mask: Tensor = None):
        H = self.mab0(self.I.expand(X.size(0), -1, -1).contiguous(), X, mask)
        return self.mab1(X, H)          # second MAB attends to every inducing point (reference :191)


def _attn_mask(mask: Tensor) -> Optional[Tensor]:
    """Kept for API parity (reference :194-202).  The kernels consume the JetNet mask itself and
    apply the ``(1 - mask).bool()`` rule per key, so this is the identity."""
    return mask


class GAPT_G(nn.Module):
    def __init__(self, num_particles: int, output_feat_size: int, sab_layers: int = 2, num_heads: int = 4,
                 embed_dim: int = 32, sab_fc_layers: list = [], layer_norm: bool = False, dropout_p: float = 0.0,
                 final_fc_layers: list = [], use_mask: bool = True, use_isab: bool = False,
                 num_isab_nodes: int = 10, linear_args: dict = {}):
        super().__init__()
        self.num_particles = num_particles
        self.output_feat_size = output_feat_size
        self.use_mask = use_mask
        self.sabs = nn.ModuleList()
        sab_args = {"embed_dim": embed_dim, "ff_layers": sab_fc_layers, "final_linear": False,
                    "num_heads": num_heads, "layer_norm": layer_norm, "dropout_p": dropout_p,
                    "linear_args": linear_args}
        for _ in range(sab_layers):
            self.sabs.append(SAB(**sab_args) if not use_isab else ISAB(num_isab_nodes, **sab_args))
        self.final_fc = LinearNet(final_fc_layers, input_size=embed_dim, output_size=output_feat_size,
                                  final_linear=True, **linear_args)

    def forward(self, x: Tensor, labels: Tensor = None):
        if not x.is_cuda:
            raise RuntimeError("mpgan_b200 modules run on CUDA only (no CPU fallback)")
        mask = ops.rank_mask(x, labels, self.num_particles) if self.use_mask else None
        for sab in self.sabs:
            x = sab(x, _attn_mask(mask))
        x = self.final_fc(x)
        if mask is None:
            return ops.activation(x, "tanh")
        return ops.gen_tail(x, mask, "tanh")


class GAPT_D(nn.Module):
    def __init__(self, num_particles: int, input_feat_size: int, sab_layers: int = 2, num_heads: int = 4,
                 embed_dim: int = 32, sab_fc_layers: list = [], layer_norm: bool = False, dropout_p: float = 0.0,
                 final_fc_layers: list = [], use_mask: bool = True, use_isab: bool = False,
                 num_isab_nodes: int = 10, linear_args: dict = {}):
        super().__init__()
        self.num_particles = num_particles
        self.input_feat_size = input_feat_size
        self.use_mask = use_mask
        self.sabs = nn.ModuleList()
        sab_args = {"embed_dim": embed_dim, "ff_layers": sab_fc_layers, "final_linear": False,
                    "num_heads": num_heads, "layer_norm": layer_norm, "dropout_p": dropout_p,
                    "linear_args": linear_args}
        self.input_embedding = LinearNet([], input_size=input_feat_size, output_size=embed_dim, **linear_args)
        for _ in range(sab_layers):
            self.sabs.append(SAB(**sab_args) if not use_isab else ISAB(num_isab_nodes, **sab_args))
        self.pma = PMA(num_seeds=1, **sab_args)
        self.final_fc = LinearNet(final_fc_layers, input_size=embed_dim, output_size=1, final_linear=True,
                                  **linear_args)

    def forward(self, x: Tensor, labels: Tensor = None):
        if not x.is_cuda:
            raise RuntimeError("mpgan_b200 modules run on CUDA only (no CPU fallback)")
        if self.use_mask:
            mask = ops.split_mask(x)
            x = x[..., :-1]
        else:
            mask = None
        x = self.input_embedding(x)
        for sab in self.sabs:
            x = sab(x, _attn_mask(mask))
        x = self.pma(x, _attn_mask(mask)).squeeze(1)
        return ops.activation(self.final_fc(x), "sigmoid")
