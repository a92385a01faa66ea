"""fp32 CPU restatement of GAPT's masked set attention (TEST INFRASTRUCTURE ONLY).

Multi-head attention is written out explicitly (packed in_proj rows 0:E=Q, E:2E=K, 2E:3E=V;
head h = channels h*d:(h+1)*d; scale 1/sqrt(d); -inf on ignored keys) instead of calling
``nn.MultiheadAttention``, so that the oracle pins what the CUDA kernel must compute.
Pinned against the unmodified reference by ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from .mpgan_oracle import linear_net, rank_mask

Tensor = torch.Tensor


@dataclass
class GaptCfg:
    """gapt/model.py:206-221, 278-293."""

    num_particles: int = 30
    num_heads: int = 4
    embed_dim: int = 64
    sab_layers: int = 2
    use_mask: bool = True
    use_isab: bool = False
    layer_norm: bool = False
    alpha: float = 0.2
    dropout_p: float = 0.0  # MAB dropout
    linear_dropout_p: float = 0.0  # LinearNet dropout (linear_args)


def mha(x: Tensor, y: Tensor, sd, prefix: str, heads: int, ignore: Optional[Tensor]) -> Tensor:
    """nn.MultiheadAttention(E, heads, batch_first=True)(x, y, y, attn_mask) (gapt/model.py:107,129).

    ``ignore``: bool [B, Nk], True = key is masked out (same for every query and head, which is
    what SAB/ISAB/PMA build at gapt/model.py:127,152,172,189).
    """
    B, Nq, E = x.shape
    Nk = y.shape[1]
    d = E // heads
    w, b = sd[prefix + ".in_proj_weight"], sd[prefix + ".in_proj_bias"]
    q = F.linear(x, w[:E], b[:E]).view(B, Nq, heads, d).transpose(1, 2)
    k = F.linear(y, w[E : 2 * E], b[E : 2 * E]).view(B, Nk, heads, d).transpose(1, 2)
    v = F.linear(y, w[2 * E :], b[2 * E :]).view(B, Nk, heads, d).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(d)
    if ignore is not None:
        s = s.masked_fill(ignore[:, None, None, :], float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, Nq, E)
    return F.linear(o, sd[prefix + ".out_proj.weight"], sd[prefix + ".out_proj.bias"])


def mab(x, y, sd, prefix, cfg: GaptCfg, ignore=None, training=False):
    """gapt/model.py:124-139: residual attn -> [LN] -> Dropout -> residual ff -> [LN] -> Dropout."""
    x = x + mha(x, y, sd, prefix + ".attention", cfg.num_heads, ignore)
    if cfg.layer_norm:
        x = F.layer_norm(x, (x.shape[-1],), sd[prefix + ".norm1.weight"], sd[prefix + ".norm1.bias"])
    x = F.dropout(x, cfg.dropout_p, training)
    # ff = LinearNet(ff_layers, E->E, final_linear=False) -> lrelu + dropout apply (:108-114,229-237)
    x = x + linear_net(x, sd, prefix + ".ff", False, cfg.alpha, cfg.linear_dropout_p, training)
    if cfg.layer_norm:
        x = F.layer_norm(x, (x.shape[-1],), sd[prefix + ".norm2.weight"], sd[prefix + ".norm2.bias"])
    return F.dropout(x, cfg.dropout_p, training)


def sab_or_isab(x, sd, prefix, cfg: GaptCfg, ignore, training=False):
    if not cfg.use_isab:  # SAB (:148-154)
        return mab(x, x, sd, prefix + ".mab", cfg, ignore, training)
    I = sd[prefix + ".I"].expand(x.shape[0], -1, -1)  # ISAB (:187-191)
    H = mab(I, x, sd, prefix + ".mab0", cfg, ignore, training)
    return mab(x, H, sd, prefix + ".mab1", cfg, None, training)  # second MAB unmasked


def _ignore(mask: Optional[Tensor]) -> Optional[Tensor]:
    """_attn_mask: (1 - mask).bool(), True = ignore (:194-202) -> any mask != 1.0 is ignored."""
    return None if mask is None else (1 - mask).bool().squeeze(-1)


def gapt_g(sd, x: Tensor, labels: Optional[Tensor], cfg: GaptCfg, training=False):
    mask = rank_mask(x[:, :, 0], labels[:, -1], cfg.num_particles) if cfg.use_mask else None  # :255-262
    for i in range(cfg.sab_layers):
        x = sab_or_isab(x, sd, f"sabs.{i}", cfg, _ignore(mask), training)  # :269-270
    x = torch.tanh(linear_net(x, sd, "final_fc", True, cfg.alpha, cfg.linear_dropout_p, training))  # :272
    return torch.cat((x, mask - 0.5), dim=2) if mask is not None else x  # :274


def gapt_d(sd, x: Tensor, labels: Optional[Tensor], cfg: GaptCfg, training=False):
    mask = None
    if cfg.use_mask:  # :333-335
        mask = x[..., -1:] + 0.5
        x = x[..., :-1]
    x = linear_net(x, sd, "input_embedding", False, cfg.alpha, cfg.linear_dropout_p, training)  # :339
    for i in range(cfg.sab_layers):
        x = sab_or_isab(x, sd, f"sabs.{i}", cfg, _ignore(mask), training)  # :341-342
    S = sd["pma.S"].expand(x.shape[0], -1, -1)  # PMA (:170-174), one seed
    p = mab(S, x, sd, "pma.mab", cfg, _ignore(mask), training)
    out = linear_net(p.squeeze(1), sd, "final_fc", True, cfg.alpha, cfg.linear_dropout_p, training)
    return torch.sigmoid(out)  # :344
