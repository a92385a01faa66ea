"""Drop-in MPGAN modules backed by the sm_100a kernels.

Same constructor signatures, attribute names and ``state_dict`` layout as the reference
(``mpgan/model.py``: ``LinearNet`` :11-88, ``MPLayer`` :91-384, ``MPNet`` :387-569, ``MPGenerator``
:572-757, ``MPDiscriminator`` :760-894), so ``trained_models/mp_*`` load with ``strict=True``.
Everything below ``forward`` runs in ``libmpgan_b200.so``; there is no PyTorch fallback, and
options the kernels do not cover raise ``NotImplementedError`` at construction or call time.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch import Tensor

from . import ops
from .spectral_normalization import SpectralNorm


class LinearNet(nn.Module):
    """Fully connected network with leaky-relu activations (reference ``LinearNet``, model.py:11-88).

    Each layer is Linear -> leaky_relu (skipped on the last layer iff ``final_linear``) -> Dropout
    (always), executed as one fused GEMM-epilogue kernel per layer.
    """

    def __init__(
        self,
        layers: list,
        input_size: int = 0,
        output_size: int = 0,
        final_linear: bool = False,
        leaky_relu_alpha: float = 0.2,
        dropout_p: float = 0,
        batch_norm: bool = False,
        spectral_norm: bool = False,
    ):
        super().__init__()
        if batch_norm:
            raise NotImplementedError(
                "batch_norm couples the jets of a batch (BatchNorm1d over B*N^2 rows) and is not "
                "supported by the fused kernels; the reference default is batch_norm=False")
        self.final_linear = final_linear
        self.leaky_relu_alpha = leaky_relu_alpha
        self.batch_norm = batch_norm
        self.dropout_p = float(dropout_p)
        self.dropout = nn.Dropout(p=dropout_p)  # kept for repr / attribute parity; math is in-kernel

        layers = layers.copy()
        if input_size:
            layers.insert(0, input_size)
        if output_size:
            layers.append(output_size)

        self.net = nn.ModuleList()
        for i in range(len(layers) - 1):
            self.net.append(nn.Linear(layers[i], layers[i + 1]))

        if spectral_norm:
            for i in range(len(self.net)):
                if i != len(self.net) - 1 or not final_linear:
                    self.net[i] = SpectralNorm(self.net[i])

    def layer_params(self, i: int):
        """(weight, bias) of layer ``i``; runs the spectral-norm power iteration if wrapped."""
        layer = self.net[i]
        if isinstance(layer, SpectralNorm):
            return layer.compute_weight(), layer.module.bias
        return layer.weight, layer.bias

    def has_act(self, i: int) -> bool:
        return i != len(self.net) - 1 or not self.final_linear

    def forward(self, x: Tensor):
        p = self.dropout_p if self.training else 0.0
        seed = ops.next_seed() if p > 0 else 0   # one seed per call; layer i draws from RNG stream 16 + i
        for i in range(len(self.net)):
            w, b = self.layer_params(i)
            x = ops.linear(x, w, b, self.has_act(i), self.leaky_relu_alpha, p, rng_stream=16 + i, seed=seed)
        return x

    def __repr__(self):
        return f"{self.__class__.__name__}(net = {self.net})"


class MPLayer(nn.Module):
    """Fully-connected message-passing layer (reference ``MPLayer``, model.py:91-384).

    ``fe`` over every (receiver i, sender j) pair, mask on the sender axis, sum/mean over senders
    and the node network ``fn``.  The first ``fe`` layer is factorised per node and the
    ``[B*N*N, hidden]`` tensors never reach HBM (``ops.edge_aggregate``).
    """

    def __init__(
        self,
        input_node_size: int,
        fe_layers: list,
        fn_layers: list,
        output_node_size: int,
        pos_diffs: bool = False,
        all_ef: bool = True,
        coords: str = "polarrel",
        delta_coords: bool = False,
        delta_r: bool = True,
        int_diffs: bool = False,
        clabels: int = 0,
        mask_fne_np: bool = False,
        fully_connected: bool = True,
        num_knn: int = 20,
        self_loops: bool = True,
        sum: bool = True,
        **linear_args,
    ):
        super().__init__()
        if int_diffs:
            raise NotImplementedError("int_diffs is not implemented in the reference either")
        if len(fe_layers) != 3:
            raise NotImplementedError("the fused edge kernel is built for a 3-layer edge network")

        self.input_node_size = input_node_size
        self.output_node_size = output_node_size
        self.fe_layers = fe_layers
        self.fn_layers = fn_layers
        self.pos_diffs = pos_diffs
        self.all_ef = all_ef
        self.coords = coords
        self.delta_coords = delta_coords
        self.delta_r = delta_r
        self.int_diffs = int_diffs
        self.clabels = clabels
        self.mask_fne_np = mask_fne_np
        self.fully_connected = fully_connected
        self.num_knn = num_knn
        self.self_loops = self_loops
        self.sum = sum

        num_ef = 0
        if pos_diffs:
            if delta_coords:
                num_ef += 3 if coords == "cartesian" else 2
            if delta_r or all_ef:
                num_ef += 1
        self.num_ef = num_ef

        # kernel-side description of the pair features (reference _getA_fully_connected :284-317)
        self._ef_mode, self._nd = 0, 0
        if pos_diffs:
            ncoord = 3 if coords == "cartesian" else 2
            self._nd = input_node_size if all_ef else ncoord
            if delta_r and delta_coords:
                self._ef_mode = 3
            elif delta_r or all_ef:
                self._ef_mode = 1
            elif delta_coords:
                self._ef_mode = 2
            cols = (self._nd if self._ef_mode & 2 else 0) + (self._ef_mode & 1)
            if fully_connected and cols != num_ef:
                raise ValueError(
                    f"pair-feature options give {cols} columns but the edge network expects {num_ef} "
                    "(the reference fails with a shape error for this combination)")
        if not fully_connected:
            # kNN (reference _getA_knn :319-381): neighbours by distance over all features unless pos_diffs without
            # all_ef (then the coordinates); with pos_diffs the ONLY pair feature is that distance (:380-383)
            ncoord = 3 if coords == "cartesian" else 2
            self._nd = input_node_size if (all_ef or not pos_diffs) else ncoord
            self._ef_mode = 1 if pos_diffs else 0
            if pos_diffs and num_ef != 1:
                raise ValueError("kNN message passing feeds one distance column; these pair-feature options make the "
                                 f"edge network expect {num_ef} (the reference fails with a shape error)")

        fe_in_size = 2 * input_node_size + num_ef + clabels + mask_fne_np
        self.fe = LinearNet(self.fe_layers, input_size=fe_in_size, final_linear=False, **linear_args)
        fe_out_size = self.fe_layers[-1]
        fn_in_size = fe_out_size + input_node_size + clabels + mask_fne_np
        self.fn = LinearNet(self.fn_layers, input_size=fn_in_size, output_size=output_node_size,
                            final_linear=True, **linear_args)

    def forward(self, x: Tensor, use_mask: bool = False, mask: Tensor = None, labels: Tensor = None,
                num_jet_particles: Tensor = None):
        batch_size, num_nodes = x.size(0), x.size(1)
        assert not (use_mask and mask is None), "need ``mask`` tensor if using ``use_mask`` option"
        fe = self.fe
        w0, b0 = fe.layer_params(0)
        w1, b1 = fe.layer_params(1)
        w2, b2 = fe.layer_params(2)
        p = fe.dropout_p if self.training else 0.0
        ncond = self.clabels + int(self.mask_fne_np)
        cond = lc = None
        if ncond:
            # conditioning columns (reference :247-253, 270-276): labels[:, :clabels] and / or the particle count, fed to
            # BOTH networks.  The reference builds them with .repeat, which pairs row r of the edge / node list with jet
            # r % B -- reproduced as is.  In the edge network they only shift the first layer: Lc = cond W0c^T.
            assert not (self.clabels and labels is None), "need ``labels`` tensor if using ``clabels`` option"
            assert not (self.mask_fne_np and num_jet_particles is None), \
                "need ``num_jet_particles`` tensor if using ``mask_fne_np`` option"
            parts = ([labels[:, :self.clabels]] if self.clabels else []) + \
                ([num_jet_particles.reshape(batch_size, 1)] if self.mask_fne_np else [])
            cond = torch.cat(parts, 1).detach().float().contiguous()
            nmain = w0.shape[1] - ncond
            lc = ops.linear(cond, w0[:, nmain:], None, False, 0.0, 0.0)
            w0 = w0[:, :nmain]
        if self.fully_connected and lc is None:
            agg = ops.edge_aggregate(x, mask if use_mask else None, w0, b0, w1, b1, w2, b2, ef_mode=self._ef_mode,
                                     nd=self._nd, mean=not self.sum, alpha=fe.leaky_relu_alpha, p_drop=p)
        else:
            m = mask if use_mask else None
            nbr = None if self.fully_connected else ops.knn_select(x, m, self.num_knn, self._nd, self.self_loops)
            agg = ops.edge_aggregate_knn(x, m, nbr, w0, b0, w1, b1, w2, b2, ef_mode=self._ef_mode, nd=self._nd,
                                         mean=not self.sum, alpha=fe.leaky_relu_alpha, p_drop=p, lc=lc)
        if cond is not None:
            x = ops.cond_columns(x, cond)      # (x | cond) for the node network
        fn = self.fn
        pn = fn.dropout_p if self.training else 0.0
        if len(fn.net) == 3 and fn.final_linear and ops.node_net_supported(
                agg.shape[2], x.shape[2], self._fn_out(0), self._fn_out(1),
                self._fn_out(2), pn):
            # cat(agg, x) -> fn as ONE kernel per direction (the cat is never materialised)
            (f0, c0), (f1, c1), (f2, c2) = fn.layer_params(0), fn.layer_params(1), fn.layer_params(2)
            return ops.node_net(agg, x, f0, c0, f1, c1, f2, c2, fn.leaky_relu_alpha, pn)
        h = torch.cat((agg, x), 2).view(batch_size * num_nodes, -1)
        h = self.fn(h)
        return h.view(batch_size, num_nodes, self.output_node_size)

    def _fn_out(self, i: int) -> int:
        layer = self.fn.net[i]
        return (layer.module if isinstance(layer, SpectralNorm) else layer).out_features

    def __repr__(self):
        return f"MPLayer(fe = {self.fe}, fn = {self.fn})"


class MPNet(nn.Module):
    """Base message-passing network (reference ``MPNet``, model.py:387-569)."""

    def __init__(
        self,
        num_particles: int,
        input_node_size: int,
        mp_iters: int = 2,
        fe_layers: list = [96, 160, 192],
        fn_layers: list = [256, 256],
        fe1_layers: list = None,
        fn1_layers: list = None,
        hidden_node_size: int = 32,
        output_node_size: int = 0,
        final_activation: str = "",
        linear_args: dict = {},
        mp_args: dict = {},
        mp_args_first_layer: dict = {},
        mask_args: dict = {},
    ):
        super().__init__()
        self.num_particles = num_particles
        self.input_node_size = input_node_size
        self.output_node_size = output_node_size if output_node_size > 0 else hidden_node_size
        self.mp_iters = mp_iters
        fe1_layers = fe_layers if fe1_layers is None else fe1_layers
        fn1_layers = fn_layers if fn1_layers is None else fn1_layers
        self.hidden_node_size = hidden_node_size
        self.final_activation = final_activation
        self.linear_args = linear_args

        mp_args_first_layer = dict(mp_args_first_layer)
        for key in mp_args:
            if key not in mp_args_first_layer:
                mp_args_first_layer[key] = mp_args[key]

        self.mask_args = mask_args
        self._init_mask(**mask_args)

        self.mp_layers = nn.ModuleList()
        self.mp_layers.append(MPLayer(input_node_size, fe1_layers, fn1_layers, hidden_node_size,
                                      **mp_args_first_layer, **linear_args))
        for _ in range(mp_iters - 2):
            self.mp_layers.append(MPLayer(hidden_node_size, fe_layers, fn_layers, hidden_node_size,
                                          **mp_args, **linear_args))
        self.mp_layers.append(MPLayer(hidden_node_size, fe_layers, fn_layers, self.output_node_size,
                                      **mp_args, **linear_args))
        # The reference's conditioning columns pair row r of its flattened lists with jet r % B: the result then depends
        # on the order of jets and particles, so the layout permutations (real particles first; jets by count in the
        # trainer) must stay off for such networks.
        self.order_dependent = any(l.clabels or l.mask_fne_np for l in self.mp_layers)

    def forward(self, x: Tensor, labels: Tensor = None) -> Tensor:
        if not x.is_cuda:
            raise RuntimeError("mpgan_b200 modules run on CUDA only (no CPU fallback)")
        x = self._pre_mp(x, labels)
        x, use_mask, mask, num_jet_particles = self._get_mask(x, labels, **self.mask_args)
        idx = None
        if use_mask and self.sort_particles and x.shape[1] > 1 and not mask.requires_grad and not self.order_dependent:
            # (a mask that carries a gradient -- D differentiated w.r.t. its input's mask channel -- keeps its layout)
            # Real particles first inside every jet.  The layers are permutation-equivariant over particles
            # (fully connected, sum / mean aggregation), so this only changes the layout: a generated jet's real
            # particles are wherever its noise ranks put them, and a sender index that is padded in every jet of a
            # 128-particle tile is a (tile, sender) step the edge kernels drop.
            idx, mask_sorted = ops.particle_order(mask)       # idx[b, i] = new position of particle i
            x = ops.permute_rows(x, idx, 0)
            mask_orig, mask = mask, mask_sorted
        cmap = None
        if (use_mask and self.compact_receivers and ops.get_precision() == 1 and not mask.requires_grad
                and not self.order_dependent and all(l.fully_connected for l in self.mp_layers)):
            # Padded particles are dropped downstream (masked as senders, multiplied by the mask at the pooling):
            # the edge kernels need not compute what they RECEIVE either.  One map per call, shared by the layers.
            cmap = ops.compact_map(mask)
        with ops.receiver_compaction(cmap):
            for i in range(self.mp_iters):
                x = self.mp_layers[i](x, use_mask, mask, labels, num_jet_particles)
        if idx is not None and not self._pool_is_order_free():
            x = ops.permute_rows(x, idx, 1)                    # back to the caller's particle order
            mask = mask_orig
        x = self._post_mp(x, labels, use_mask, mask, num_jet_particles)
        return self._tail(x, mask)

    sort_particles = True     # class-level switch (tests compare both layouts)
    compact_receivers = False  # only networks that drop their padded particles downstream may switch this on

    def _pool_is_order_free(self) -> bool:
        """True if ``_post_mp`` reduces over particles (then the permutation need not be undone)."""
        return False

    def _tail(self, x, mask):
        x = self._final_activation(x)
        return self._final_mask(x, mask, **self.mask_args)

    def _pre_mp(self, x, labels):
        return x

    def _post_mp(self, x, labels, use_mask, mask, num_jet_particles):
        return x

    def _final_activation(self, x):
        return ops.activation(x, self.final_activation)

    def _init_mask(self, **mask_args):
        return

    def _get_mask(self, x: Tensor, labels: Tensor, **mask_args):
        return x, False, None, None

    def _final_mask(self, x: Tensor, mask: Tensor, **mask_args):
        return x

    def __repr__(self):
        return f"MPLayers = {self.mp_layers})"


class MPGenerator(MPNet):
    """Message-passing generator (reference ``MPGenerator``, model.py:572-757)."""

    def __init__(self, lfc: bool = False, lfc_latent_size: int = 128, **mpnet_args):
        super().__init__(**mpnet_args)
        self.lfc = lfc
        if lfc:
            self.lfc_layer = nn.Linear(lfc_latent_size, self.num_particles * self.input_node_size)

    def _pre_mp(self, x, labels):
        if self.lfc:
            x = ops.linear(x, self.lfc_layer.weight, self.lfc_layer.bias, False, 0.0, 0.0)
            x = x.reshape(x.shape[0], self.num_particles, self.input_node_size)
        return x

    def _init_mask(self, mask_learn: bool = False, mask_learn_sep: bool = False, fmg: list = [64], **mask_args):
        if mask_learn or mask_learn_sep:
            # the reference constructor fails here too (model.py:626 reads an undefined attribute)
            raise NotImplementedError("mask_learn / mask_learn_sep are broken in the reference and unsupported")

    def _get_mask(self, x: Tensor, labels: Tensor = None, mask_learn: bool = False, mask_learn_bin: bool = True,
                  mask_learn_sep: bool = False, mask_c: bool = True, mask_fne_np: bool = False, **mask_args):
        use_mask = mask_learn or mask_c or mask_learn_sep
        if not use_mask:
            return x, use_mask, None, None
        if not mask_c:
            raise NotImplementedError("only the mask_c masking strategy is supported")
        assert labels is not None, "mask_c needs ``labels`` (last column = normalised particle count)"
        mask = ops.rank_mask(x, labels, self.num_particles)
        num_jet_particles = None
        return x, use_mask, mask, num_jet_particles

    def _final_mask(self, x: Tensor, mask: Tensor, mask_feat_bin: bool = False, **mask_args):
        if mask_feat_bin:
            raise NotImplementedError("mask_feat_bin fails with a shape error in the reference; unsupported")
        return ops.gen_tail(x, mask, "") if mask is not None else x

    def _tail(self, x, mask):
        if self.mask_args.get("mask_feat_bin", False):
            raise NotImplementedError("mask_feat_bin fails with a shape error in the reference; unsupported")
        if mask is None:
            return self._final_activation(x)
        return ops.gen_tail(x, mask, self.final_activation)  # act + cat(mask - 0.5) in one kernel

    def __repr__(self):
        lfc_str = f"LFC = {self.lfc_layer},\n" if self.lfc else ""
        return f"{self.__class__.__name__}({lfc_str}MPLayers = {self.mp_layers})"


class MPDiscriminator(MPNet):
    """Message-passing discriminator (reference ``MPDiscriminator``, model.py:760-894)."""

    def __init__(self, dea: bool = True, dea_sum: bool = True, fnd: list = [], mask_fnd_np: bool = False,
                 **mpnet_args):
        super().__init__(output_node_size=1 if not dea else 0, **mpnet_args)
        self.dea = dea
        self.dea_sum = dea_sum
        self.mask_fnd_np = mask_fnd_np
        if mask_fnd_np:
            raise NotImplementedError("mask_fnd_np is not supported")
        if dea:
            self.fnd_layer = LinearNet(fnd, input_size=self.hidden_node_size + int(mask_fnd_np), output_size=1,
                                       final_linear=True, **self.linear_args)

    compact_receivers = True   # reference :810-822,881-884: x * mask before the pooling, mask on the sender axis

    def _pool_is_order_free(self) -> bool:
        return True   # masked sum / mean over particles (with or without the fnd head)

    def _post_mp(self, x, labels, use_mask, mask, num_jet_particles):
        do_mean = not (self.dea and self.dea_sum)
        x = ops.masked_pool(x, mask if use_mask else None, do_mean)
        if self.dea:
            x = self.fnd_layer(x)
        return x

    def _get_mask(self, x: Tensor, labels: Tensor, mask_manual: bool = False, mask_learn: bool = False,
                  mask_learn_sep: bool = False, mask_c: bool = True, mask_fne_np: bool = False,
                  mask_fnd_np: bool = False, **mask_args):
        mask = None
        use_mask = mask_manual or mask_learn or mask_c or mask_learn_sep
        if use_mask or mask_fnd_np:
            mask = ops.split_mask(x)          # x[:, :, -1:] + 0.5, a real-valued multiplier (differentiable)
        if use_mask:
            x = x[:, :, :-1]                  # strided view: the kernels take a row stride, no copy
        njp = None
        if mask_fne_np:                       # reference :886-887: mean of the mask over particles, a conditioning value
            njp = ops.masked_pool(mask.detach(), None, True)
        return x, use_mask, mask, njp

    def __repr__(self):
        dea_str = f",\nFND = {self.fnd_layer}" if self.dea else ""
        return f"{self.__class__.__name__}(MPLayers = {self.mp_layers}{dea_str})"
