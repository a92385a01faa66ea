"""Autograd bindings of the C-ABI kernels (include/mpgan_b200.h).

Each op is a ``torch.autograd.Function`` whose forward/backward call straight into
``libmpgan_b200.so`` on the current CUDA stream.  PyTorch is used for device memory and autograd
bookkeeping only.

Double backward (``create_graph=True``: the WGAN-GP term, train.py:286-324) is supported for the ops a
discriminator is made of -- ``linear``, ``edge_aggregate``, ``node_net``, ``split_mask``, ``masked_pool`` -- through
second functions (``*BwdFn``) whose forward is the first-order backward and whose backward holds the second-order
products (the networks are piecewise linear, so these are tangent passes along the slopes of the primal pass, see
``mpg_edge_bwd2``).  Everything else is ``once_differentiable`` and raises under ``create_graph`` instead of silently
producing wrong gradients.
"""
from __future__ import annotations

import torch
from torch.autograd.function import once_differentiable

from . import _lib

# --------------------------------------------------------------------------------------------------
# global knobs
# --------------------------------------------------------------------------------------------------
_PRECISION = 1          # 0: fp32-class (3xTF32 + fp32 SIMT edge), 1: fast (TF32 + bf16 tcgen05 edge)
_seed_counter = 0
_device_seed = None     # optional int64 CUDA tensor added to every dropout seed (CUDA-graph replay)


def set_precision(p: int):
    """0 = fp32-class reference accuracy, 1 = fast tensor-core path (default)."""
    global _PRECISION
    if p not in (0, 1):
        raise ValueError("precision must be 0 or 1")
    _PRECISION = p


def get_precision() -> int:
    return _PRECISION


# When on, weight/bias gradients of leaf parameters that already own a ``.grad`` buffer (train.FlatParams) are
# accumulated by the kernels straight into that buffer and the autograd functions return ``None`` for them:
# no zero-filled temporaries, no AccumulateGrad add kernel per parameter and backward pass.
_DIRECT_GRAD = False
# Keep the forward's edge workspace (P/Q, weight images, work list: ~0.8 KB per particle) alive for the backward
# instead of rebuilding it there.
_SAVE_EDGE_WS = True


def set_direct_grad(on: bool):
    global _DIRECT_GRAD
    _DIRECT_GRAD = bool(on)


class direct_grad:
    """Context manager scoping the direct-gradient mode (train.GANTrainer wraps its own backward passes in it, so
    modules outside the trainer keep ordinary autograd semantics: returned gradients, hooks, autograd.grad)."""

    def __init__(self, on: bool = True):
        self.on = bool(on)

    def __enter__(self):
        global _DIRECT_GRAD
        self.prev, _DIRECT_GRAD = _DIRECT_GRAD, self.on
        return self

    def __exit__(self, *exc):
        global _DIRECT_GRAD
        _DIRECT_GRAD = self.prev
        return False


# Inside ``input_grad_only()`` the backward passes skip weight / bias gradients altogether (they return None for
# them): ``torch.autograd.grad(D(x), x, create_graph=True)`` of the gradient penalty asks for dD/dx only, but
# ``needs_input_grad`` is True for every parameter that requires grad.
_INPUT_GRAD_ONLY = False


class input_grad_only:
    def __enter__(self):
        global _INPUT_GRAD_ONLY
        self.prev, _INPUT_GRAD_ONLY = _INPUT_GRAD_ONLY, True
        return self

    def __exit__(self, *exc):
        global _INPUT_GRAD_ONLY
        _INPUT_GRAD_ONLY = self.prev
        return False


def _want_wgrad(ctx, lo, hi):
    return (not _INPUT_GRAD_ONLY) and any(ctx.needs_input_grad[lo:hi])


def _grad_sink(p):
    """``p.grad`` if the kernels may accumulate into it directly, else ``None``."""
    if _DIRECT_GRAD and p.is_leaf and p.requires_grad and p.grad is not None and p.grad.is_contiguous() \
            and p.grad.dtype == torch.float32:
        return p.grad
    return None


def set_device_seed(t):
    """Register a 1-element int64 CUDA tensor whose value is added to every dropout seed."""
    global _device_seed
    if t is not None and not (t.is_cuda and t.dtype == torch.int64 and t.numel() == 1):
        raise ValueError("device seed must be a 1-element int64 CUDA tensor")
    _device_seed = t


def next_seed() -> int:
    """Fresh 64-bit dropout seed derived from torch's seed and a call counter."""
    global _seed_counter
    _seed_counter += 1
    return (torch.initial_seed() * 0x9E3779B97F4A7C15 + _seed_counter * 0xD1342543DE82EF95) % (1 << 64)


def _seed_ptr():
    return None if _device_seed is None else _device_seed.data_ptr()


# optional per-call device timing of the edge-network launches (bench.py's roofline leg): a list of
# (name, start_event, end_event, algorithmic_flops) filled while profiling is on
_profile = None


def profile_start():
    global _profile
    _profile = []
    return _profile


def profile_stop():
    global _profile
    p, _profile = _profile, None
    return p


class _Timed:
    """Brackets one C-ABI edge call with events; additionally arms the library's one-shot probes so
    the tcgen05 kernels inside the call are timed on their own (kernel ids: include/mpgan_b200.h)."""

    def __init__(self, name, flops, probes=(), frac=1.0):
        # frac: share of the (tile, sender) steps the kernels execute (fully masked steps are skipped)
        self.name, self.flops, self.probes, self.frac = name, flops, probes, frac

    def __enter__(self):
        if _profile is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.pe = []
            L = _lib.lib()
            for kid, kname, kflops in self.probes:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()  # materialise the cudaEvent_t handles
                b.record()
                L.mpg_probe(kid, a.cuda_event, b.cuda_event)
                self.pe.append((kname, a, b, kflops, kflops * self.frac))
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if _profile is not None:
            self.e1.record()
            _profile.append((self.name, self.e0, self.e1, self.flops, self.flops * self.frac))
            _profile.extend(self.pe)
        return False


def edge_flops(B, N, F, H0, H1, H2):
    """Algorithmic forward FLOPs of one fused edge call (first layer factorised)."""
    return 4.0 * B * N * F * H0 + 2.0 * B * N * N * (H0 * H1 + H1 * H2)


def _pair_flops(B, N, Ha, Hb):
    return 2.0 * B * N * N * Ha * Hb


def edge_active_fraction(mask, B, N, cmap=None):
    """Share of the (128-receiver tile, sender) steps of the padded problem the tcgen05 kernels execute: a step is
    dropped when the sender is masked in every jet the tile touches (csrc/edge_tc_common.cuh: step_list_block); with a
    receiver compaction map the tiles are the map's.  Reporting helper (synchronises); not on the compute path."""
    if mask is None:
        return 1.0
    m = (mask.reshape(B, N) != 0).to(torch.int32)
    BN = B * N
    tiles = (BN + 127) // 128
    if cmap is not None:
        tmax = (int(cmap.numel()) - 2) // 130
        nt = int(cmap[0].item())
        j0 = cmap[2:2 + nt].long()
        j1 = j0 + cmap[2 + tmax:2 + tmax + nt].long() - 1
    else:
        r0 = torch.arange(tiles, device=m.device) * 128
        j0 = r0 // N
        j1 = torch.clamp(r0 + 127, max=BN - 1) // N
    cs = torch.cat((torch.zeros(1, N, dtype=torch.int32, device=m.device), m.cumsum(0).to(torch.int32)), 0)
    act = (cs[j1 + 1] - cs[j0]) > 0
    return float(act.sum().item()) / float(tiles * N)


def _rows(t):
    """View a [..., K] tensor as rows with a constant row stride (no copy when possible)."""
    if t.dim() == 2 and t.stride(1) == 1:
        return t, t.stride(0)
    if t.stride(-1) == 1 and t.dim() == 3 and t.stride(0) == t.shape[1] * t.stride(1):
        return t, t.stride(1)
    t = t.contiguous()
    return t, t.shape[-1]


# --------------------------------------------------------------------------------------------------
# LinearNet layer
# --------------------------------------------------------------------------------------------------
class LinearFn(torch.autograd.Function):
    """y = dropout(act(x W^T + b)) on rows; mpgan/model.py:77-83."""

    @staticmethod
    def forward(ctx, x, w, b, act, alpha, p_drop, rng_stream, seed=None):
        L = _lib.lib()
        x2, ldx = _rows(x)
        lead = x2.shape[:-1]
        M = int(x2.numel() // x2.shape[-1]) if x2.shape[-1] else 0
        K, N = x2.shape[-1], w.shape[0]
        w = w.contiguous()
        b = None if b is None else b.contiguous()
        y = torch.empty(*lead, N, device=x.device, dtype=torch.float32)
        if seed is None:
            seed = next_seed() if p_drop > 0 else 0
        _lib.check(L.mpg_linear_fwd(_lib.ptr(x2), ldx, _lib.ptr(w), _lib.ptr(b), _lib.ptr(y), M, K, N, int(act),
                                    float(alpha), float(p_drop), seed, _seed_ptr(), int(rng_stream), _PRECISION,
                                    _lib.stream()), "mpg_linear_fwd")
        ctx.save_for_backward(x2, w, y)
        ctx.params = (w, b)
        ctx.cfg = (ldx, M, K, N, int(act), float(alpha), float(p_drop), seed, int(rng_stream), _PRECISION,
                   _seed_ptr())
        return y

    @staticmethod
    def backward(ctx, dy):
        L = _lib.lib()
        x2, w, y = ctx.saved_tensors
        ldx, M, K, N, act, alpha, p, seed, rstream, prec, sptr = ctx.cfg
        if torch.is_grad_enabled():   # create_graph: differentiable backward
            wp, bp = ctx.params
            need_w = _want_wgrad(ctx, 1, 3)
            dx, dw, db = LinearBwdFn.apply(dy, y, x2, wp, ctx.cfg, ctx.needs_input_grad[0], need_w)
            return (dx if ctx.needs_input_grad[0] else None, dw if need_w else None, db if need_w else None, None, None,
                    None, None, None)
        dy = dy.contiguous()
        dz = torch.empty_like(dy) if (act or p > 0) else None
        dx = torch.empty(*x2.shape, device=dy.device, dtype=torch.float32) if ctx.needs_input_grad[0] else None
        wp, bp = ctx.params
        need_w = ctx.needs_input_grad[1] and not _INPUT_GRAD_ONLY
        need_b = ctx.needs_input_grad[2] and not _INPUT_GRAD_ONLY
        dw_sink = _grad_sink(wp) if need_w else None
        db_sink = _grad_sink(bp) if need_b else None
        dw = dw_sink if dw_sink is not None else (torch.zeros_like(w) if need_w else None)
        db = db_sink if db_sink is not None else (
            torch.zeros(N, device=dy.device, dtype=torch.float32) if need_b else None)
        _lib.check(L.mpg_linear_bwd(_lib.ptr(dy), _lib.ptr(y), _lib.ptr(x2), ldx, _lib.ptr(w), _lib.ptr(dz),
                                    _lib.ptr(dx), K, 0, _lib.ptr(dw), _lib.ptr(db), M, K, N, act, alpha, p, seed,
                                    sptr, rstream, prec, _lib.stream()), "mpg_linear_bwd")
        return (dx, None if dw_sink is not None else dw, None if db_sink is not None else db, None, None, None, None,
                None)


class LinearBwdFn(torch.autograd.Function):
    """First-order backward of ``LinearFn`` as a differentiable function of (dy, w): dz = dy * S with S = d(act,
    dropout)/dz read off the layer output y (piecewise constant), dx = dz W, dw = dz^T x, db = colsum(dz).  Its
    backward (the double backward) for the cotangent v of dx:  d<v,dx>/d(dy) = S * (v W^T),  d<v,dx>/dW = dz^T v;
    x and y get none (piecewise linear).  Cotangents of dw / db are not supported (nothing on this path forms them)."""

    @staticmethod
    def forward(ctx, dy, y, x2, w, cfg, need_dx, need_w):
        L = _lib.lib()
        ldx, M, K, N, act, alpha, p, seed, rstream, prec, sptr = cfg
        dy = dy.contiguous()
        wc = w.contiguous()
        dz = torch.empty_like(dy) if (act or p > 0) else None
        dx = torch.empty(*x2.shape, device=dy.device, dtype=torch.float32) if need_dx else None
        dw = torch.zeros_like(wc) if need_w else None
        db = torch.zeros(N, device=dy.device, dtype=torch.float32) if need_w else None
        _lib.check(L.mpg_linear_bwd(_lib.ptr(dy), _lib.ptr(y), _lib.ptr(x2), ldx, _lib.ptr(wc), _lib.ptr(dz),
                                    _lib.ptr(dx), K, 0, _lib.ptr(dw), _lib.ptr(db), M, K, N, act, alpha, p, seed,
                                    sptr, rstream, prec, _lib.stream()), "mpg_linear_bwd")
        ctx.save_for_backward(dy, y, wc)
        ctx.cfg = cfg
        ctx.xshape = x2.shape
        if dx is None:
            dx = torch.zeros(0, device=dy.device)
        if dw is None:
            dw, db = torch.zeros(0, device=dy.device), torch.zeros(0, device=dy.device)
        ctx.mark_non_differentiable(dw, db)
        return dx, dw, db

    @staticmethod
    @once_differentiable
    def backward(ctx, v, v_w, v_b):
        L = _lib.lib()
        dy, y, w = ctx.saved_tensors
        ldx, M, K, N, act, alpha, p, seed, rstream, prec, sptr = ctx.cfg
        g_dy = g_w = None
        if v is not None and v.numel():
            v = v.contiguous().view(M, K)
            if ctx.needs_input_grad[0]:
                t = torch.empty(M, N, device=v.device, dtype=torch.float32)      # v W^T
                _lib.check(L.mpg_linear_fwd(_lib.ptr(v), K, _lib.ptr(w), None, _lib.ptr(t), M, K, N, 0, 0.0, 0.0, 0, None,
                                            0, prec, _lib.stream()), "mpg_linear_fwd")
                if act or p > 0:                                                 # * S (dz-only call)
                    g_dy = torch.empty_like(t)
                    _lib.check(L.mpg_linear_bwd(_lib.ptr(t), _lib.ptr(y), None, 0, _lib.ptr(w), _lib.ptr(g_dy), None, K, 0,
                                                None, None, M, K, N, act, alpha, p, seed, sptr, rstream, prec,
                                                _lib.stream()), "mpg_linear_bwd")
                else:
                    g_dy = t
                g_dy = g_dy.view(dy.shape)
            if ctx.needs_input_grad[3]:
                g_w = torch.zeros_like(w)                                        # dz^T v
                dz = torch.empty(M, N, device=v.device, dtype=torch.float32) if (act or p > 0) else None
                _lib.check(L.mpg_linear_bwd(_lib.ptr(dy), _lib.ptr(y), _lib.ptr(v), K, _lib.ptr(w), _lib.ptr(dz), None, K, 0,
                                            _lib.ptr(g_w), None, M, K, N, act, alpha, p, seed, sptr, rstream, prec,
                                            _lib.stream()), "mpg_linear_bwd")
        return g_dy, None, None, g_w, None, None, None


def linear(x, w, b, act: bool, alpha: float, p_drop: float, rng_stream: int = 16, seed=None):
    """``seed``: dropout seed (None draws a fresh one); layers of one LinearNet call share a seed and differ
    in ``rng_stream``, which is what lets the fused node-network kernel reproduce their masks."""
    return LinearFn.apply(x, w, b, act, alpha, p_drop, rng_stream, seed)


# --------------------------------------------------------------------------------------------------
# fused edge network + aggregation
# --------------------------------------------------------------------------------------------------
_CMAP = None            # receiver compaction map of the edge calls inside ``receiver_compaction`` (int32 tensor or None)


def compact_map(mask):
    """Map that packs the particles with mask != 0 into the 128-row tiles of the tcgen05 edge kernels
    (``mpg_compact_map``): padded particles then cost nothing as receivers either.  ``mask`` [B, N(, 1)]."""
    L = _lib.lib()
    B, N = int(mask.shape[0]), int(mask.shape[1])
    m = mask.detach().reshape(B, N).contiguous().float()
    cmap = torch.empty(int(L.mpg_compact_map_ints(B, N)), device=m.device, dtype=torch.int32)
    scratch = torch.empty(2 * B + 8, device=m.device, dtype=torch.int32)
    _lib.check(L.mpg_compact_map(_lib.ptr(m), B, N, _lib.ptr(cmap), _lib.ptr(scratch), _lib.stream()), "mpg_compact_map")
    return cmap


class receiver_compaction:
    """``with receiver_compaction(cmap):`` the fused edge ops inside build their tiles from ``cmap`` (None: no-op).
    Exact only where padded particles are dropped downstream (the discriminator); see include/mpgan_b200.h."""

    def __init__(self, cmap):
        self.cmap = cmap

    def __enter__(self):
        global _CMAP
        self.prev, _CMAP = _CMAP, self.cmap

    def __exit__(self, *exc):
        global _CMAP
        _CMAP = self.prev


class _compaction_for_call:
    """Arms / disarms the library's thread-local map around one C call."""

    def __init__(self, cmap):
        self.cmap = cmap

    def __enter__(self):
        if self.cmap is not None:
            _lib.lib().mpg_edge_set_compaction(_lib.ptr(self.cmap))

    def __exit__(self, *exc):
        if self.cmap is not None:
            _lib.lib().mpg_edge_set_compaction(None)


class EdgeAggFn(torch.autograd.Function):
    """agg[b,i] = scale * sum_j mask[b,j] fe(x_i | x_j | ef_ij); mpgan/model.py:256-267,284-317."""

    @staticmethod
    def forward(ctx, x, mask, w0, b0, w1, b1, w2, b2, ef_mode, nd, mean, alpha, p_drop):
        L = _lib.lib()
        x3, ldx = _rows(x)
        B, N, F = x3.shape
        H0, H1, H2 = w0.shape[0], w1.shape[0], w2.shape[0]
        # forward-only part of the workspace (P/Q, weight images, work list); kept for the backward when one follows
        ws_bytes = L.mpg_edge_fwd_workspace_bytes(B, N, F, H0, H1, H2)
        ws = torch.empty(ws_bytes, device=x.device, dtype=torch.uint8)
        agg = torch.empty(B, N, H2, device=x.device, dtype=torch.float32)
        m = None if mask is None else mask.reshape(B, N).contiguous()
        ws_ = [t.contiguous() for t in (w0, b0, w1, b1, w2, b2)]
        seed = next_seed() if p_drop > 0 else 0
        cmap = _CMAP if (m is not None and _PRECISION == 1 and ef_mode == 0) else None
        ctx.cmap = cmap
        frac = edge_active_fraction(m, B, N, cmap) if _profile is not None else 1.0
        with _Timed("edge_fwd", edge_flops(B, N, F, H0, H1, H2),
                    [(1, "edge_tc_fwd_kernel", _pair_flops(B, N, H0, H1) + _pair_flops(B, N, H1, H2))], frac), \
                _compaction_for_call(cmap):
            _lib.check(L.mpg_edge_fwd(_lib.ptr(x3), ldx, _lib.ptr(m), *[_lib.ptr(t) for t in ws_], B, N, F, H0, H1,
                                      H2, int(ef_mode), int(nd), int(mean), float(alpha), float(p_drop), seed,
                                      _seed_ptr(), _PRECISION, ws.data_ptr(), ws_bytes, _lib.ptr(agg),
                                      _lib.stream()), "mpg_edge_fwd")
        # (x, mask themselves are saved too: the double backward returns gradients w.r.t. the ORIGINAL inputs)
        ctx.save_for_backward(x3, m, *ws_, x, mask if mask is not None else x3.new_zeros(0))
        ctx.params = (w0, b0, w1, b1, w2, b2)
        ctx.mask_shape = None if mask is None else tuple(mask.shape)
        ctx.fwd_ws = ws if _SAVE_EDGE_WS else None
        ctx.cfg = (ldx, B, N, F, H0, H1, H2, int(ef_mode), int(nd), int(mean), float(alpha), float(p_drop), seed,
                   _PRECISION, _seed_ptr())
        return agg

    @staticmethod
    def backward(ctx, dagg):
        L = _lib.lib()
        x3, m, w0, b0, w1, b1, w2, b2, x_in, mask_in = ctx.saved_tensors
        ldx, B, N, F, H0, H1, H2, ef_mode, nd, mean, alpha, p, seed, prec, sptr = ctx.cfg
        need_mask = m is not None and ctx.needs_input_grad[1]
        if torch.is_grad_enabled():   # create_graph: differentiable backward (second-order products in EdgeAggBwdFn)
            need_w = _want_wgrad(ctx, 2, 8)
            outs = EdgeAggBwdFn.apply(dagg, x_in, mask_in if m is not None else None, *ctx.params, ctx.cfg, need_mask,
                                      need_w)
            dx, dmask, grads = outs[0], outs[1], outs[2:]
            return (dx, dmask.view(ctx.mask_shape) if need_mask else None, *(grads if need_w else [None] * 6), None, None,
                    None, None, None)
        dagg = dagg.contiguous()
        ws_bytes = L.mpg_edge_workspace_bytes(B, N, F, H0, H1, H2)
        ws = torch.empty(ws_bytes, device=dagg.device, dtype=torch.uint8)
        dx = torch.empty(B, N, F, device=dagg.device, dtype=torch.float32)
        want_w = _want_wgrad(ctx, 2, 8)
        sinks = [_grad_sink(p) for p in ctx.params] if (want_w and all(ctx.needs_input_grad[2:8])) else [None] * 6
        direct = all(g is not None for g in sinks)
        if direct:
            grads = sinks
        elif want_w:
            # one zero-filled buffer, six views (one fill kernel instead of six)
            ws_ = (w0, b0, w1, b1, w2, b2)
            flat = torch.zeros(sum(t.numel() for t in ws_), device=dagg.device, dtype=torch.float32)
            grads, off = [], 0
            for t in ws_:
                grads.append(flat[off:off + t.numel()].view_as(t))
                off += t.numel()
        else:   # frozen weights (train_G back-propagating through D): input gradient only
            grads = [None] * 6
        dmask = None
        if need_mask:
            # the mask multiplier itself needs a gradient (D differentiated w.r.t. its input's mask channel): the fp32
            # kernels compute it next to dx
            dmask = torch.empty(B, N, device=dagg.device, dtype=torch.float32)
            _lib.check(L.mpg_edge_nbr_bwd(None, 0, 0, None, _lib.ptr(x3), ldx, _lib.ptr(m), *[_lib.ptr(t) for t in (w0, b0, w1, b1, w2, b2)],
                                          B, N, F, H0, H1, H2, ef_mode, nd, mean, alpha, p, seed, sptr, ws.data_ptr(), ws_bytes,
                                          _lib.ptr(dagg), _lib.ptr(dx), F, _lib.ptr(dmask), None, *[_lib.ptr(g) for g in grads],
                                          _lib.stream()), "mpg_edge_nbr_bwd")
            dmask = dmask.view(ctx.mask_shape)
        else:
            frac = edge_active_fraction(m, B, N, getattr(ctx, "cmap", None)) if _profile is not None else 1.0
            with _Timed("edge_bwd", 2.0 * edge_flops(B, N, F, H0, H1, H2),
                        [(2, "edge_tc_bwd_chain_kernel", 2 * _pair_flops(B, N, H0, H1) + _pair_flops(B, N, H1, H2)),
                         (3, "edge_tc_bwd_dw2_kernel", _pair_flops(B, N, H1, H2))], frac), \
                    _compaction_for_call(getattr(ctx, "cmap", None)):
                args = (_lib.ptr(x3), ldx, _lib.ptr(m), *[_lib.ptr(t) for t in (w0, b0, w1, b1, w2, b2)], B, N, F, H0, H1,
                        H2, ef_mode, nd, mean, alpha, p, seed, sptr, prec, ws.data_ptr(), ws_bytes, _lib.ptr(dagg),
                        _lib.ptr(dx), F, *[_lib.ptr(g) for g in grads], _lib.stream())
                fws = ctx.fwd_ws
                if fws is not None:   # P/Q, weight images and work list as the forward left them
                    _lib.check(L.mpg_edge_bwd_saved(fws.data_ptr(), fws.numel(), *args), "mpg_edge_bwd_saved")
                else:
                    _lib.check(L.mpg_edge_bwd(*args), "mpg_edge_bwd")
        if direct:
            grads = [None] * 6
        return (dx, dmask, *grads, None, None, None, None, None)


class EdgeAggBwdFn(torch.autograd.Function):
    """First-order backward of ``EdgeAggFn`` as a differentiable function of (dagg, x, mask, weights); outputs
    (dx, dmask, dw0, db0, dw1, db1, dw2, db2).  Its backward -- the double backward the WGAN-GP term needs -- takes the
    cotangents u of dx and vm of dmask (none of the weight gradients: nothing on this path forms them) and returns

      d/d(dagg) = tagg(u) + agg(x; mask := vm)          tangent pass + an ordinary forward with vm as the mask
      d/d(mask) = gmask(u)
      d/d(x)    = dx(x, dagg; mask := vm)               ordinary backward with vm as the mask
      d/dW      = dW2nd(u) + dW(x, dagg; mask := vm)

    because fe is piecewise linear: <u, dx> = sum_ij mask_j <J_ij u, dagg_i> and <vm, dmask> = sum_ij vm_j <fe_ij,
    dagg_i> (mpg_edge_bwd2; fp32 kernels; same dropout seed as the forward)."""

    @staticmethod
    def forward(ctx, dagg, x, mask, w0, b0, w1, b1, w2, b2, cfg, need_mask, need_w):
        L = _lib.lib()
        _, B, N, F, H0, H1, H2, ef_mode, nd, mean, alpha, p, seed, prec, sptr = cfg
        x3, ldx = _rows(x)
        m = None if mask is None else mask.reshape(B, N).contiguous()
        cfg = (ldx,) + tuple(cfg[1:])
        dev = dagg.device
        dagg = dagg.contiguous()
        ws_ = [t.contiguous() for t in (w0, b0, w1, b1, w2, b2)]
        ws_bytes = L.mpg_edge_workspace_bytes(B, N, F, H0, H1, H2)
        ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
        dx = torch.empty(B, N, F, device=dev, dtype=torch.float32)
        grads = [torch.zeros_like(t) for t in ws_] if need_w else [None] * 6
        dmask = torch.empty(B, N, device=dev, dtype=torch.float32) if need_mask else None
        if need_mask:
            _lib.check(L.mpg_edge_nbr_bwd(None, 0, 0, None, _lib.ptr(x3), ldx, _lib.ptr(m), *[_lib.ptr(t) for t in ws_], B, N, F,
                                          H0, H1, H2, ef_mode, nd, mean, alpha, p, seed, sptr, ws.data_ptr(), ws_bytes,
                                          _lib.ptr(dagg), _lib.ptr(dx), F, _lib.ptr(dmask), None, *[_lib.ptr(g) for g in grads],
                                          _lib.stream()), "mpg_edge_nbr_bwd")
        else:
            _lib.check(L.mpg_edge_bwd(_lib.ptr(x3), ldx, _lib.ptr(m), *[_lib.ptr(t) for t in ws_], B, N, F, H0, H1, H2,
                                      ef_mode, nd, mean, alpha, p, seed, sptr, prec, ws.data_ptr(), ws_bytes,
                                      _lib.ptr(dagg), _lib.ptr(dx), F, *[_lib.ptr(g) for g in grads], _lib.stream()),
                       "mpg_edge_bwd")
        ctx.save_for_backward(dagg, x3, m, *ws_)
        ctx.cfg = cfg
        ctx.mask_in_shape = None if mask is None else tuple(mask.shape)
        empty = torch.zeros(0, device=dev)
        outs = [dx, dmask if need_mask else empty] + [g if need_w else empty for g in grads]
        ctx.mark_non_differentiable(*outs[2:])
        return tuple(outs)

    @staticmethod
    @once_differentiable
    def backward(ctx, u, vm, *v_w):
        L = _lib.lib()
        dagg, x3, m, w0, b0, w1, b1, w2, b2 = ctx.saved_tensors
        ldx, B, N, F, H0, H1, H2, ef_mode, nd, mean, alpha, p, seed, prec, sptr = ctx.cfg
        if ef_mode != 0:
            raise NotImplementedError("double backward through pair features (pos_diffs) is not implemented")
        dev = dagg.device
        ws_ = (w0, b0, w1, b1, w2, b2)
        need_w = any(ctx.needs_input_grad[3:9])
        g_dagg = torch.zeros(B, N, H2, device=dev, dtype=torch.float32)
        g_x = g_mask = None
        g_ws = [torch.zeros_like(t) for t in ws_] if need_w else [None] * 6
        have_u = u is not None and u.numel() > 0
        have_vm = vm is not None and vm.numel() > 0 and m is not None
        if have_u:
            u = u.contiguous()
            ws_bytes = L.mpg_edge_bwd2_workspace_bytes(B, N, F, H0, H1, H2)
            ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
            g_mask = torch.empty(B, N, device=dev, dtype=torch.float32) if (m is not None and ctx.needs_input_grad[2]) else None
            _lib.check(L.mpg_edge_bwd2(_lib.ptr(x3), ldx, _lib.ptr(u), F, _lib.ptr(m), *[_lib.ptr(t) for t in ws_], B, N, F,
                                       H0, H1, H2, mean, alpha, p, seed, sptr, ws.data_ptr(), ws_bytes, _lib.ptr(dagg),
                                       _lib.ptr(g_dagg), _lib.ptr(g_mask), _lib.ptr(g_ws[0]), _lib.ptr(g_ws[2]),
                                       _lib.ptr(g_ws[4]), _lib.stream()), "mpg_edge_bwd2")
        if have_vm:
            vmc = vm.contiguous().view(B, N)
            ws_bytes = L.mpg_edge_workspace_bytes(B, N, F, H0, H1, H2)
            ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
            agg_vm = torch.empty(B, N, H2, device=dev, dtype=torch.float32)
            _lib.check(L.mpg_edge_fwd(_lib.ptr(x3), ldx, _lib.ptr(vmc), *[_lib.ptr(t) for t in ws_], B, N, F, H0, H1, H2,
                                      ef_mode, nd, mean, alpha, p, seed, sptr, 0, ws.data_ptr(), ws_bytes,
                                      _lib.ptr(agg_vm), _lib.stream()), "mpg_edge_fwd")
            g_dagg += agg_vm
            g_x = torch.empty(B, N, F, device=dev, dtype=torch.float32)
            _lib.check(L.mpg_edge_bwd(_lib.ptr(x3), ldx, _lib.ptr(vmc), *[_lib.ptr(t) for t in ws_], B, N, F, H0, H1, H2,
                                      ef_mode, nd, mean, alpha, p, seed, sptr, 0, ws.data_ptr(), ws_bytes, _lib.ptr(dagg),
                                      _lib.ptr(g_x), F, *[_lib.ptr(g) for g in g_ws], _lib.stream()), "mpg_edge_bwd")
        # gradients are w.r.t. this function's inputs: x [B, N, F] (dense), mask in its original shape
        if g_mask is not None:
            g_mask = g_mask.view(ctx.mask_in_shape)
        return (g_dagg, g_x, g_mask, *g_ws, None, None, None)


def edge_aggregate(x, mask, w0, b0, w1, b1, w2, b2, ef_mode=0, nd=0, mean=False, alpha=0.2, p_drop=0.0):
    return EdgeAggFn.apply(x, mask, w0, b0, w1, b1, w2, b2, ef_mode, nd, mean, alpha, p_drop)


# --------------------------------------------------------------------------------------------------
# fused node network fn: cat(agg, x) -> H1 -> H2 -> out in one tcgen05 kernel per direction
# --------------------------------------------------------------------------------------------------
class NodeNetFn(torch.autograd.Function):
    """out = LinearNet([H1, H2] -> NO, final_linear)(cat(agg, x)); mpgan/model.py:268-279 with :70-85."""

    @staticmethod
    def forward(ctx, agg, x, w0, b0, w1, b1, w2, b2, alpha, p_drop, seed):
        L = _lib.lib()
        a2, lda = _rows(agg)
        x2, ldx = _rows(x)
        Ka, Kb = a2.shape[-1], x2.shape[-1]
        M = int(a2.numel() // Ka)
        H1, H2, NO = w0.shape[0], w1.shape[0], w2.shape[0]
        ws_ = [t.contiguous() for t in (w0, b0, w1, b1, w2, b2)]
        ws_bytes = L.mpg_fn_workspace_bytes(Ka, Kb, H1, H2, NO)
        ws = torch.empty(ws_bytes, device=agg.device, dtype=torch.uint8)
        y0 = torch.empty(M, H1, device=agg.device, dtype=torch.float32)
        y1 = torch.empty(M, H2, device=agg.device, dtype=torch.float32)
        out = torch.empty(*a2.shape[:-1], NO, device=agg.device, dtype=torch.float32)
        if seed is None:
            seed = next_seed() if p_drop > 0 else 0
        _lib.check(L.mpg_fn_fwd(_lib.ptr(a2), lda, Ka, _lib.ptr(x2), ldx, Kb, M, *[_lib.ptr(t) for t in ws_], H1, H2,
                                NO, float(alpha), float(p_drop), seed, _seed_ptr(), ws.data_ptr(), ws_bytes,
                                _lib.ptr(y0), _lib.ptr(y1), _lib.ptr(out), _lib.stream()), "mpg_fn_fwd")
        ctx.save_for_backward(a2, x2, y0, y1, ws_[0], ws_[2], ws_[4])
        ctx.params = (w0, b0, w1, b1, w2, b2)
        ctx.cfg = (lda, ldx, Ka, Kb, M, H1, H2, NO, float(alpha), float(p_drop), seed, _seed_ptr())
        return out

    @staticmethod
    def backward(ctx, dout):
        L = _lib.lib()
        a2, x2, y0, y1, w0, w1, w2 = ctx.saved_tensors
        lda, ldx, Ka, Kb, M, H1, H2, NO, alpha, p, seed, sptr = ctx.cfg
        if torch.is_grad_enabled():
            # create_graph: the chain layer by layer through the differentiable per-layer backward (same saved layer
            # outputs, same dropout seed / streams 16, 17, 18 as the fused kernel)
            need_w = _want_wgrad(ctx, 2, 8)
            pw0, pb0, pw1, pb1, pw2, pb2 = ctx.params
            lead = dout.shape[:-1]
            cat = torch.cat((a2.reshape(M, Ka), x2.reshape(M, Kb)), 1)
            cfg = lambda K, N, act, st: (K, M, K, N, act, alpha, p, seed, st, 1, sptr)
            d1, gw2, gb2 = LinearBwdFn.apply(dout.reshape(M, NO), dout.reshape(M, NO), y1, pw2, cfg(H2, NO, 0, 18), True, need_w)
            d0, gw1, gb1 = LinearBwdFn.apply(d1, y1, y0, pw1, cfg(H1, H2, 1, 17), True, need_w)
            dc, gw0, gb0 = LinearBwdFn.apply(d0, y0, cat, pw0, cfg(Ka + Kb, H1, 1, 16), True, need_w)
            da = dc[:, :Ka].reshape(*lead, Ka) if ctx.needs_input_grad[0] else None
            db = dc[:, Ka:].reshape(*lead, Kb) if ctx.needs_input_grad[1] else None
            gs = (gw0, gb0, gw1, gb1, gw2, gb2) if need_w else (None,) * 6
            return (da, db, *gs, None, None, None)
        dev = dout.device
        dout = dout.contiguous()
        ws_bytes = L.mpg_fn_workspace_bytes(Ka, Kb, H1, H2, NO)
        ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
        dz0 = torch.empty(M, H1, device=dev, dtype=torch.float32)
        dz1 = torch.empty(M, H2, device=dev, dtype=torch.float32)
        dz2 = torch.empty(M, NO, device=dev, dtype=torch.float32) if p > 0 else None
        da = torch.empty(*a2.shape, device=dev, dtype=torch.float32)
        db = torch.empty(*x2.shape, device=dev, dtype=torch.float32)
        need_w = _want_wgrad(ctx, 2, 8)
        sinks = [_grad_sink(q) for q in ctx.params] if (need_w and all(ctx.needs_input_grad[2:8])) else [None] * 6
        direct = all(g is not None for g in sinks)
        if direct:
            grads = sinks
        elif need_w:
            flat = torch.zeros(sum(q.numel() for q in ctx.params), device=dev, dtype=torch.float32)
            grads, off = [], 0
            for q in ctx.params:
                grads.append(flat[off:off + q.numel()].view(q.shape))
                off += q.numel()
        else:
            grads = [None] * 6
        _lib.check(L.mpg_fn_bwd(_lib.ptr(dout), _lib.ptr(y0), _lib.ptr(y1), _lib.ptr(a2), lda, Ka, _lib.ptr(x2), ldx,
                                Kb, M, _lib.ptr(w0), _lib.ptr(w1), _lib.ptr(w2), H1, H2, NO, alpha, p, seed, sptr,
                                ws.data_ptr(), ws_bytes, _lib.ptr(dz0), _lib.ptr(dz1), _lib.ptr(dz2), _lib.ptr(da),
                                _lib.ptr(db), *[_lib.ptr(g) for g in grads], _lib.stream()), "mpg_fn_bwd")
        if direct:
            grads = [None] * 6
        return (da if ctx.needs_input_grad[0] else None, db if ctx.needs_input_grad[1] else None, *grads,
                None, None, None)


def node_net_supported(Ka, Kb, H1, H2, NO, p_drop) -> bool:
    """True iff the fused tcgen05 node network covers this shape in the current precision mode."""
    return _PRECISION == 1 and bool(_lib.lib().mpg_fn_supported(int(Ka), int(Kb), int(H1), int(H2), int(NO),
                                                                float(p_drop)))


def node_net(agg, x, w0, b0, w1, b1, w2, b2, alpha: float, p_drop: float, seed=None):
    return NodeNetFn.apply(agg, x, w0, b0, w1, b1, w2, b2, alpha, p_drop, seed)


# --------------------------------------------------------------------------------------------------
# masks, tails, pooling
# --------------------------------------------------------------------------------------------------
def rank_mask(x, labels, num_particles: int):
    """mask = rank(x[:, :, 0]) <= int(labels[:, -1] * N) - 1 as fp32 [B, N, 1] (bit-exact)."""
    L = _lib.lib()
    B, N = x.shape[0], x.shape[1]
    x3, ldx = _rows(x.detach())
    lab = labels.detach()[:, -1].contiguous().float()
    mask = torch.empty(B, N, 1, device=x.device, dtype=torch.float32)
    if N != num_particles:
        raise RuntimeError(f"rank_mask: x has {N} particles, model was built for {num_particles}")
    _lib.check(L.mpg_rank_mask(_lib.ptr(x3), ldx, _lib.ptr(lab), 1, B, N, _lib.ptr(mask), _lib.stream()),
               "mpg_rank_mask")
    return mask


def particle_order(mask):
    """(pos int32 [B, N], mask in the new order [B, N, 1]): real particles (mask != 0) first, stable."""
    L = _lib.lib()
    B, N = mask.shape[0], mask.shape[1]
    m = mask.detach().reshape(B, N).contiguous()
    pos = torch.empty(B, N, device=mask.device, dtype=torch.int32)
    ms = torch.empty(B, N, 1, device=mask.device, dtype=torch.float32)
    _lib.check(L.mpg_particle_order(_lib.ptr(m), B, N, _lib.ptr(pos), _lib.ptr(ms), _lib.stream()),
               "mpg_particle_order")
    return pos, ms


class PermuteRowsFn(torch.autograd.Function):
    """mode 0: out[b, pos[b,i]] = x[b, i];  mode 1: out[b, i] = x[b, pos[b,i]]  (adjoint of each other)."""

    @staticmethod
    def forward(ctx, x, pos, mode):
        L = _lib.lib()
        x3, ldx = _rows(x)
        B, N, F = x3.shape
        out = torch.empty(B, N, F, device=x.device, dtype=torch.float32)
        _lib.check(L.mpg_permute_rows(_lib.ptr(x3), ldx, _lib.ptr(out), F, _lib.ptr(pos), B, N, F, int(mode),
                                      _lib.stream()), "mpg_permute_rows")
        ctx.save_for_backward(pos)
        ctx.mode = int(mode)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        L = _lib.lib()
        (pos,) = ctx.saved_tensors
        dout = dout.contiguous()
        B, N, F = dout.shape
        dx = torch.empty_like(dout)
        _lib.check(L.mpg_permute_rows(_lib.ptr(dout), F, _lib.ptr(dx), F, _lib.ptr(pos), B, N, F, 1 - ctx.mode,
                                      _lib.stream()), "mpg_permute_rows")
        return dx, None, None


def permute_rows(x, pos, mode: int):
    return PermuteRowsFn.apply(x, pos, mode)


class LSLossFn(torch.autograd.Function):
    """mean((d[:n0] - t0)^2) + mean((d[n0:] - t1)^2): the least-squares GAN losses of train.py:357-378,467-472."""

    @staticmethod
    def forward(ctx, d, n0, t0, t1):
        L = _lib.lib()
        d1 = d.reshape(-1).contiguous()
        loss = torch.empty((), device=d.device, dtype=torch.float32)
        _lib.check(L.mpg_ls_loss_fwd(_lib.ptr(d1), d1.numel(), int(n0), float(t0), float(t1), _lib.ptr(loss),
                                     _lib.stream()), "mpg_ls_loss_fwd")
        ctx.save_for_backward(d1)
        ctx.cfg = (int(n0), float(t0), float(t1), d.shape)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        L = _lib.lib()
        (d1,) = ctx.saved_tensors
        n0, t0, t1, shape = ctx.cfg
        dd = torch.empty_like(d1)
        g = gout.contiguous().float()
        _lib.check(L.mpg_ls_loss_bwd(_lib.ptr(d1), _lib.ptr(g), d1.numel(), n0, t0, t1, _lib.ptr(dd), _lib.stream()),
                   "mpg_ls_loss_bwd")
        return dd.view(shape), None, None, None


def ls_loss(d, n_first: int, target_first: float, target_rest: float = 0.0):
    return LSLossFn.apply(d, n_first, target_first, target_rest)


def batch_order(labels):
    """pos int32 [B]: index of jet b when the batch is ordered by descending labels[:, -1] (particle count)."""
    L = _lib.lib()
    lab = labels.detach()
    if lab.dim() == 1:
        lab = lab.unsqueeze(1)
    lab = lab.float()
    if lab.stride(1) != 1:
        lab = lab.contiguous()
    B = lab.shape[0]
    key = lab[:, -1]
    pos = torch.empty(B, device=lab.device, dtype=torch.int32)
    _lib.check(L.mpg_batch_order(key.data_ptr(), lab.stride(0), B, _lib.ptr(pos), _lib.stream()), "mpg_batch_order")
    return pos


def permute_batch(x, pos, mode: int):
    """Rows of a batch-first tensor moved by ``pos`` (mode 0: out[pos[b]] = x[b]; mode 1: out[b] = x[pos[b]])."""
    B = x.shape[0]
    return permute_rows(x.reshape(1, B, -1), pos.view(1, B), mode).view(x.shape)


def _split_mask_raw(x):
    L = _lib.lib()
    x3, ldx = _rows(x)
    if ldx != x3.shape[-1]:
        x3 = x3.contiguous()
        ldx = x3.shape[-1]
    B, N = x3.shape[0], x3.shape[1]
    mask = torch.empty(B, N, 1, device=x.device, dtype=torch.float32)
    _lib.check(L.mpg_split_mask(_lib.ptr(x3), ldx, B * N, _lib.ptr(mask), _lib.stream()), "mpg_split_mask")
    return mask


class SplitMaskFn(torch.autograd.Function):
    """mask = x[..., -1:] + 0.5, differentiable (mpgan/model.py:881, gapt/model.py:334): the gradient lands in the last
    column of dx.  Linear, so the double backward is the slice of the cotangent."""

    @staticmethod
    def forward(ctx, x):
        ctx.xshape = tuple(x.shape)
        return _split_mask_raw(x)

    @staticmethod
    def backward(ctx, dmask):
        if torch.is_grad_enabled():
            return SplitMaskBwdFn.apply(dmask, ctx.xshape)
        return _split_mask_bwd_raw(dmask, ctx.xshape)


def _split_mask_bwd_raw(dmask, xshape):
    L = _lib.lib()
    dm = dmask.contiguous()
    dx = torch.empty(xshape, device=dm.device, dtype=torch.float32)
    _lib.check(L.mpg_split_mask_bwd(_lib.ptr(dm), _lib.ptr(dx), xshape[-1], dm.numel(), _lib.stream()),
               "mpg_split_mask_bwd")
    return dx


class SplitMaskBwdFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dmask, xshape):
        ctx.mshape = tuple(dmask.shape)
        return _split_mask_bwd_raw(dmask, xshape)

    @staticmethod
    @once_differentiable
    def backward(ctx, v):
        return (_split_mask_raw(v) - 0.5).view(ctx.mshape), None


def split_mask(x):
    """mask = x[..., -1:] + 0.5 (fp32 multiplier) for the discriminator input.  Differentiable when ``x`` requires
    grad, unless ``x`` is marked as having a constant mask channel (``GenTailFn`` output: a generated jet's mask is a
    function of the noise ranks only, so its gradient would be discarded upstream anyway)."""
    if x.requires_grad and torch.is_grad_enabled() and not getattr(x, "_mpg_const_mask", False):
        return SplitMaskFn.apply(x)
    return _split_mask_raw(x.detach())


_ACT = {"": 0, "tanh": 1, "sigmoid": 2}


class GenTailFn(torch.autograd.Function):
    """cat(act(h), mask - 0.5); mpgan/model.py:535-536,752."""

    @staticmethod
    def forward(ctx, h, mask, act):
        L = _lib.lib()
        h = h.contiguous()
        B, N, Fo = h.shape
        ldo = Fo + (mask is not None)
        out = torch.empty(B, N, ldo, device=h.device, dtype=torch.float32)
        m = None if mask is None else mask.reshape(B, N).contiguous()
        _lib.check(L.mpg_gen_tail_fwd(_lib.ptr(h), _lib.ptr(m), _lib.ptr(out), B * N, Fo, act, _lib.stream()),
                   "mpg_gen_tail_fwd")
        ctx.save_for_backward(out)
        ctx.cfg = (B, N, Fo, ldo, act)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        L = _lib.lib()
        (out,) = ctx.saved_tensors
        B, N, Fo, ldo, act = ctx.cfg
        dout = dout.contiguous()
        dh = torch.empty(B, N, Fo, device=dout.device, dtype=torch.float32)
        _lib.check(L.mpg_gen_tail_bwd(_lib.ptr(dout), _lib.ptr(out), _lib.ptr(dh), B * N, Fo, ldo, act,
                                      _lib.stream()), "mpg_gen_tail_bwd")
        return dh, None, None


def gen_tail(h, mask, activation: str):
    out = GenTailFn.apply(h, mask, _ACT[activation])
    if mask is not None:
        out._mpg_const_mask = True   # see split_mask: nothing upstream consumes d/d(mask channel)
    return out


class PoolFn(torch.autograd.Function):
    """sum_i h[b,i]*mask[b,i] (optionally / sum mask); mpgan/model.py:810-822."""

    @staticmethod
    def forward(ctx, h, mask, mean):
        L = _lib.lib()
        h = h.contiguous()
        B, N, Cc = h.shape
        m = None if mask is None else mask.reshape(B, N).contiguous()
        out = torch.empty(B, Cc, device=h.device, dtype=torch.float32)
        _lib.check(L.mpg_pool_fwd(_lib.ptr(h), _lib.ptr(m), _lib.ptr(out), B, N, Cc, int(mean), _lib.stream()),
                   "mpg_pool_fwd")
        ctx.save_for_backward(m) if m is not None else None
        ctx.has_mask = m is not None
        ctx.cfg = (B, N, Cc, int(mean))
        return out

    @staticmethod
    def backward(ctx, dout):
        m = ctx.saved_tensors[0] if ctx.has_mask else None
        if torch.is_grad_enabled():   # linear in dout: its own adjoint is the forward pool
            return PoolBwdFn.apply(dout, m, ctx.cfg), None, None
        return _pool_bwd(dout, m, ctx.cfg), None, None


def _pool_bwd(dout, m, cfg):
    L = _lib.lib()
    B, N, Cc, mean = cfg
    dout = dout.contiguous()
    dh = torch.empty(B, N, Cc, device=dout.device, dtype=torch.float32)
    _lib.check(L.mpg_pool_bwd(_lib.ptr(dout), _lib.ptr(m), _lib.ptr(dh), B, N, Cc, mean, _lib.stream()), "mpg_pool_bwd")
    return dh


class PoolBwdFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dout, m, cfg):
        ctx.m, ctx.cfg = m, cfg
        return _pool_bwd(dout, m, cfg)

    @staticmethod
    @once_differentiable
    def backward(ctx, v):
        L = _lib.lib()
        B, N, Cc, mean = ctx.cfg
        v = v.contiguous()
        g = torch.empty(B, Cc, device=v.device, dtype=torch.float32)
        _lib.check(L.mpg_pool_fwd(_lib.ptr(v), _lib.ptr(ctx.m), _lib.ptr(g), B, N, Cc, mean, _lib.stream()), "mpg_pool_fwd")
        return g, None, None


class PoolSumFn(torch.autograd.Function):
    """sum_i h[b,i] * mask[b,i] with gradients w.r.t. BOTH h and the mask (bilinear); double-differentiable."""

    @staticmethod
    def forward(ctx, h, mask):
        L = _lib.lib()
        hc = h.contiguous()
        B, N, Cc = hc.shape
        m = mask.reshape(B, N).contiguous()
        out = torch.empty(B, Cc, device=h.device, dtype=torch.float32)
        _lib.check(L.mpg_pool_fwd(_lib.ptr(hc), _lib.ptr(m), _lib.ptr(out), B, N, Cc, 0, _lib.stream()), "mpg_pool_fwd")
        ctx.save_for_backward(h, mask)
        return out

    @staticmethod
    def backward(ctx, dout):
        h, mask = ctx.saved_tensors
        if torch.is_grad_enabled():
            return PoolSumBwdFn.apply(dout, h, mask)
        return _pool_sum_bwd(dout, h, mask)


def _pool_sum_bwd(dout, h, mask):
    L = _lib.lib()
    hc, dout = h.contiguous(), dout.contiguous()
    B, N, Cc = hc.shape
    m = mask.reshape(B, N).contiguous()
    dh = torch.empty(B, N, Cc, device=h.device, dtype=torch.float32)
    dm = torch.empty(B, N, device=h.device, dtype=torch.float32)
    _lib.check(L.mpg_pool_bwd(_lib.ptr(dout), _lib.ptr(m), _lib.ptr(dh), B, N, Cc, 0, _lib.stream()), "mpg_pool_bwd")
    _lib.check(L.mpg_pool_dmask(_lib.ptr(hc), _lib.ptr(dout), _lib.ptr(dm), B, N, Cc, 1.0, _lib.stream()), "mpg_pool_dmask")
    return dh, dm.view(mask.shape)


class PoolSumBwdFn(torch.autograd.Function):
    """(dh, dmask) = (dout * mask, <h, dout>); backward for cotangents (vh, vm):
    d/d(dout) = pool(vh, mask) + pool(h, vm),  d/dh = dout * vm,  d/d(mask) = <vh, dout>."""

    @staticmethod
    def forward(ctx, dout, h, mask):
        ctx.save_for_backward(dout, h, mask)
        return _pool_sum_bwd(dout, h, mask)

    @staticmethod
    @once_differentiable
    def backward(ctx, vh, vm):
        L = _lib.lib()
        dout, h, mask = ctx.saved_tensors
        hc, dout = h.contiguous(), dout.contiguous()
        B, N, Cc = hc.shape
        m = mask.reshape(B, N).contiguous()
        vh = vh.contiguous()
        vmc = vm.reshape(B, N).contiguous()
        g_dout = torch.empty(B, Cc, device=h.device, dtype=torch.float32)
        t = torch.empty_like(g_dout)
        _lib.check(L.mpg_pool_fwd(_lib.ptr(vh), _lib.ptr(m), _lib.ptr(g_dout), B, N, Cc, 0, _lib.stream()), "mpg_pool_fwd")
        _lib.check(L.mpg_pool_fwd(_lib.ptr(hc), _lib.ptr(vmc), _lib.ptr(t), B, N, Cc, 0, _lib.stream()), "mpg_pool_fwd")
        g_dout += t
        g_h = torch.empty(B, N, Cc, device=h.device, dtype=torch.float32)
        _lib.check(L.mpg_pool_bwd(_lib.ptr(dout), _lib.ptr(vmc), _lib.ptr(g_h), B, N, Cc, 0, _lib.stream()), "mpg_pool_bwd")
        g_m = torch.empty(B, N, device=h.device, dtype=torch.float32)
        _lib.check(L.mpg_pool_dmask(_lib.ptr(vh), _lib.ptr(dout), _lib.ptr(g_m), B, N, Cc, 1.0, _lib.stream()), "mpg_pool_dmask")
        return g_dout, g_h, g_m.view(mask.shape)


def masked_pool(h, mask, mean: bool):
    """sum_i h[b,i] * mask[b,i], divided by (sum_i mask + 1e-12) if ``mean`` (mpgan/model.py:810-822).  When the mask
    itself carries a gradient (D differentiated w.r.t. its input's mask channel, e.g. the gradient penalty) the
    bilinear sum is its own doubly differentiable op and the division is left to autograd on the [B, C] result."""
    if mask is not None and mask.requires_grad and torch.is_grad_enabled():
        s = PoolSumFn.apply(h, mask)
        return s / (mask.sum(1) + 1e-12) if mean else s
    return PoolFn.apply(h, mask, mean)


class UnaryFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, act):
        L = _lib.lib()
        x = x.contiguous()
        y = torch.empty_like(x)
        _lib.check(L.mpg_unary_fwd(_lib.ptr(x), _lib.ptr(y), x.numel(), act, _lib.stream()), "mpg_unary_fwd")
        ctx.save_for_backward(y)
        ctx.act = act
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        L = _lib.lib()
        (y,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(dy)
        _lib.check(L.mpg_unary_bwd(_lib.ptr(dy), _lib.ptr(y), _lib.ptr(dx), dy.numel(), ctx.act, _lib.stream()),
                   "mpg_unary_bwd")
        return dx, None


def activation(x, name: str):
    return x if name == "" else UnaryFn.apply(x, _ACT[name])


# --------------------------------------------------------------------------------------------------
# spectral norm
# --------------------------------------------------------------------------------------------------
class SpectralNormFn(torch.autograd.Function):
    """One power iteration (u, v updated in place, no grad) and W = W_bar / (sigma + 1e-12)."""

    @staticmethod
    def forward(ctx, w_bar, u, v):
        L = _lib.lib()
        wb = w_bar.contiguous()
        H = wb.shape[0]
        Wd = wb.numel() // H
        w = torch.empty_like(wb)
        sigma = torch.empty(1, device=wb.device, dtype=torch.float32)
        _lib.check(L.mpg_sn_fwd(_lib.ptr(wb), _lib.ptr(u), _lib.ptr(v), _lib.ptr(w), _lib.ptr(sigma), H, Wd,
                                _lib.stream()), "mpg_sn_fwd")
        ctx.save_for_backward(wb, u.clone(), v.clone(), sigma)
        return w

    @staticmethod
    @once_differentiable
    def backward(ctx, dw):
        L = _lib.lib()
        wb, u, v, sigma = ctx.saved_tensors
        H = wb.shape[0]
        Wd = wb.numel() // H
        dw = dw.contiguous()
        dwb = torch.zeros_like(wb)
        _lib.check(L.mpg_sn_bwd(_lib.ptr(dw), _lib.ptr(wb), _lib.ptr(u), _lib.ptr(v), _lib.ptr(sigma),
                                _lib.ptr(dwb), H, Wd, _lib.stream()), "mpg_sn_bwd")
        return dwb, None, None


def spectral_normalize(w_bar, u, v):
    return SpectralNormFn.apply(w_bar, u, v)


# --------------------------------------------------------------------------------------------------
# optimizer
# --------------------------------------------------------------------------------------------------
def rmsprop_(p_flat, g_flat, sq_flat, lr, alpha=0.99, eps=1e-8, gscale=1.0):
    L = _lib.lib()
    _lib.check(L.mpg_rmsprop(_lib.ptr(p_flat), _lib.ptr(g_flat), _lib.ptr(sq_flat), p_flat.numel(), float(lr),
                             float(alpha), float(eps), float(gscale), _lib.stream()), "mpg_rmsprop")


def allreduce_rmsprop_(p_flat, sq_flat, grad_ptrs, flag_ptrs, rank, world, ctas, lr, alpha=0.99, eps=1e-8, multicast=None):
    """Fused data-parallel update (mpg_allreduce_rmsprop): ``grad_ptrs`` / ``flag_ptrs`` are ctypes arrays of the
    ranks' peer-mapped gradient buffers and flag blocks; ``multicast``: the buffers' multicast address (int) or None."""
    L = _lib.lib()
    _lib.check(L.mpg_allreduce_rmsprop(_lib.ptr(p_flat), _lib.ptr(sq_flat), grad_ptrs, flag_ptrs, multicast or None,
                                       p_flat.numel(), int(rank),
                                       int(world), int(ctas), float(lr), float(alpha), float(eps), _lib.stream()),
               "mpg_allreduce_rmsprop")


# --------------------------------------------------------------------------------------------------
# GAPT set attention
# --------------------------------------------------------------------------------------------------
class AttnFn(torch.autograd.Function):
    """Masked multi-head softmax(q k^T / sqrt(d)) v on projected rows; gapt/model.py:129."""

    @staticmethod
    def forward(ctx, q, k, v, key_mask, heads):
        L = _lib.lib()
        B, Nq, E = q.shape
        Nk = k.shape[1]
        q2, ldq = _rows(q)
        k2, ldk = _rows(k)
        v2, ldv = _rows(v)
        km = None if key_mask is None else key_mask.reshape(B, Nk).contiguous().float()
        o = torch.empty(B, Nq, E, device=q.device, dtype=torch.float32)
        P = torch.empty(B, heads, Nq, Nk, device=q.device, dtype=torch.float32)
        _lib.check(L.mpg_attn_fwd(_lib.ptr(q2), ldq, _lib.ptr(k2), ldk, _lib.ptr(v2), ldv, _lib.ptr(km), B, Nq, Nk,
                                  E, heads, _lib.ptr(o), _lib.ptr(P), _lib.stream()), "mpg_attn_fwd")
        ctx.save_for_backward(q2, k2, v2, km, P)
        ctx.cfg = (ldq, ldk, ldv, B, Nq, Nk, E, heads)
        return o

    @staticmethod
    @once_differentiable
    def backward(ctx, do):
        L = _lib.lib()
        q2, k2, v2, km, P = ctx.saved_tensors
        ldq, ldk, ldv, B, Nq, Nk, E, heads = ctx.cfg
        do = do.contiguous()
        dq = torch.empty(B, Nq, E, device=do.device, dtype=torch.float32)
        dk = torch.empty(B, Nk, E, device=do.device, dtype=torch.float32)
        dv = torch.empty(B, Nk, E, device=do.device, dtype=torch.float32)
        _lib.check(L.mpg_attn_bwd(_lib.ptr(q2), ldq, _lib.ptr(k2), ldk, _lib.ptr(v2), ldv, _lib.ptr(km), B, Nq, Nk,
                                  E, heads, _lib.ptr(P), _lib.ptr(do), _lib.ptr(dq), _lib.ptr(dk), _lib.ptr(dv),
                                  _lib.stream()), "mpg_attn_bwd")
        return dq, dk, dv, None, None


def attention(q, k, v, key_mask, heads: int):
    return AttnFn.apply(q, k, v, key_mask, heads)


class ResidualDropoutFn(torch.autograd.Function):
    """out = dropout(x + r); the residual + nn.Dropout steps of MAB.forward (gapt/model.py:129-137)."""

    @staticmethod
    def forward(ctx, x, r, p_drop, rng_stream):
        L = _lib.lib()
        x = x.contiguous()
        r = None if r is None else r.contiguous()
        cols = x.shape[-1]
        rows = x.numel() // cols
        out = torch.empty_like(x)
        seed = next_seed() if p_drop > 0 else 0
        _lib.check(L.mpg_residual_dropout_fwd(_lib.ptr(x), _lib.ptr(r), _lib.ptr(out), rows, cols, float(p_drop),
                                              seed, _seed_ptr(), int(rng_stream), _lib.stream()),
                   "mpg_residual_dropout_fwd")
        ctx.cfg = (rows, cols, float(p_drop), seed, _seed_ptr(), int(rng_stream))
        ctx.has_r = r is not None
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        L = _lib.lib()
        rows, cols, p, seed, sptr, rstream = ctx.cfg
        dout = dout.contiguous()
        if p == 0:
            return dout, (dout if ctx.has_r else None), None, None
        dx = torch.empty_like(dout)
        _lib.check(L.mpg_residual_dropout_bwd(_lib.ptr(dout), _lib.ptr(dx), rows, cols, p, seed, sptr, rstream,
                                              _lib.stream()), "mpg_residual_dropout_bwd")
        return dx, (dx if ctx.has_r else None), None, None


def residual_dropout(x, r, p_drop: float, rng_stream: int = 48):
    return ResidualDropoutFn.apply(x, r, p_drop, rng_stream)


# --------------------------------------------------------------------------------------------------
# GAPT LayerNorm
# --------------------------------------------------------------------------------------------------
class LayerNormFn(torch.autograd.Function):
    """nn.LayerNorm(C) over the last dimension (gapt/model.py:116-118, 130-136)."""

    @staticmethod
    def forward(ctx, x, w, b, eps):
        L = _lib.lib()
        xc = x.contiguous()
        Cc = xc.shape[-1]
        rows = xc.numel() // Cc
        y = torch.empty_like(xc)
        mean = torch.empty(rows, device=x.device, dtype=torch.float32)
        rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
        wc, bc = w.contiguous(), b.contiguous()
        _lib.check(L.mpg_layernorm_fwd(_lib.ptr(xc), _lib.ptr(wc), _lib.ptr(bc), _lib.ptr(y), _lib.ptr(mean),
                                       _lib.ptr(rstd), rows, Cc, float(eps), _lib.stream()), "mpg_layernorm_fwd")
        ctx.save_for_backward(xc, wc, mean, rstd)
        ctx.params = (w, b)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        L = _lib.lib()
        xc, wc, mean, rstd = ctx.saved_tensors
        Cc = xc.shape[-1]
        rows = xc.numel() // Cc
        dy = dy.contiguous()
        dx = torch.empty_like(xc)
        need_w = _want_wgrad(ctx, 1, 3)
        sinks = [_grad_sink(q) for q in ctx.params] if need_w else [None, None]
        direct = all(g is not None for g in sinks)
        if direct:
            dw, db = sinks
        elif need_w:
            dw, db = torch.zeros_like(wc), torch.zeros_like(wc)
        else:
            dw = db = None
        _lib.check(L.mpg_layernorm_bwd(_lib.ptr(dy), _lib.ptr(xc), _lib.ptr(wc), _lib.ptr(mean), _lib.ptr(rstd),
                                       _lib.ptr(dx), _lib.ptr(dw), _lib.ptr(db), rows, Cc, _lib.stream()),
                   "mpg_layernorm_bwd")
        if direct:
            dw = db = None
        return dx, dw, db, None


def layer_norm(x, w, b, eps: float = 1e-5):
    return LayerNormFn.apply(x, w, b, eps)


# --------------------------------------------------------------------------------------------------
# kNN message passing (mpgan/model.py:319-381)
# --------------------------------------------------------------------------------------------------
def knn_select(x, mask, k: int, nd: int, self_loops: bool = True):
    """int32 [B, N, k]: the k nearest senders of every receiver (see mpg_knn_select)."""
    L = _lib.lib()
    x3, ldx = _rows(x.detach())
    B, N = x3.shape[0], x3.shape[1]
    m = None if mask is None else mask.detach().reshape(B, N).contiguous()
    idx = torch.empty(B, N, k, device=x.device, dtype=torch.int32)
    _lib.check(L.mpg_knn_select(_lib.ptr(x3), ldx, _lib.ptr(m), B, N, int(nd), int(k), int(bool(self_loops)),
                                _lib.ptr(idx), _lib.stream()), "mpg_knn_select")
    return idx


class EdgeNbrFn(torch.autograd.Function):
    """agg[b,i] = scale * sum_m mask[b, nbr[b,i,m]] fe(x_i | x_nbr | [dist] | [cond]) over the listed neighbours
    (``nbr`` None: all particles of the jet) on the fp32 kernels; ``lc`` [B, H0]: first-layer contribution of the
    conditioning columns (see mpg_edge_nbr_fwd)."""

    @staticmethod
    def forward(ctx, x, mask, nbr, lc, w0, b0, w1, b1, w2, b2, ef_mode, nd, mean, alpha, p_drop):
        L = _lib.lib()
        x3, ldx = _rows(x)
        B, N, F = x3.shape
        K = N if nbr is None else nbr.shape[2]
        lc = None if lc is None else lc.contiguous()
        H0, H1, H2 = w0.shape[0], w1.shape[0], w2.shape[0]
        ws_bytes = L.mpg_edge_workspace_bytes(B, N, F, H0, H1, H2)
        ws = torch.empty(ws_bytes, device=x.device, dtype=torch.uint8)
        agg = torch.empty(B, N, H2, device=x.device, dtype=torch.float32)
        m = None if mask is None else mask.reshape(B, N).contiguous()
        ws_ = [t.contiguous() for t in (w0, b0, w1, b1, w2, b2)]
        seed = next_seed() if p_drop > 0 else 0
        scale = int(m is not None and nbr is not None)
        _lib.check(L.mpg_edge_nbr_fwd(_lib.ptr(nbr), K, scale, _lib.ptr(lc), _lib.ptr(x3), ldx, _lib.ptr(m), *[_lib.ptr(t) for t in ws_],
                                      B, N, F, H0, H1, H2, int(ef_mode), int(nd), int(mean), float(alpha), float(p_drop),
                                      seed, _seed_ptr(), ws.data_ptr(), ws_bytes, _lib.ptr(agg), _lib.stream()),
                   "mpg_edge_nbr_fwd")
        ctx.save_for_backward(x3, m, nbr, lc, *ws_)
        ctx.params = (w0, b0, w1, b1, w2, b2)
        ctx.cfg = (ldx, B, N, F, K, H0, H1, H2, int(ef_mode), int(nd), int(mean), float(alpha), float(p_drop), seed,
                   scale, _seed_ptr())
        return agg

    @staticmethod
    @once_differentiable
    def backward(ctx, dagg):
        L = _lib.lib()
        x3, m, nbr, lc, w0, b0, w1, b1, w2, b2 = ctx.saved_tensors
        ldx, B, N, F, K, H0, H1, H2, ef_mode, nd, mean, alpha, p, seed, scale, sptr = ctx.cfg
        dagg = dagg.contiguous()
        ws_bytes = L.mpg_edge_workspace_bytes(B, N, F, H0, H1, H2)
        ws = torch.empty(ws_bytes, device=dagg.device, dtype=torch.uint8)
        dx = torch.empty(B, N, F, device=dagg.device, dtype=torch.float32)
        dlc = torch.empty(B, H0, device=dagg.device, dtype=torch.float32) if (lc is not None and ctx.needs_input_grad[3]) else None
        want_w = _want_wgrad(ctx, 4, 10)
        sinks = [_grad_sink(q) for q in ctx.params] if (want_w and all(ctx.needs_input_grad[4:10])) else [None] * 6
        direct = all(g is not None for g in sinks)
        if direct:
            grads = sinks
        elif want_w:
            grads = [torch.zeros_like(t) for t in (w0, b0, w1, b1, w2, b2)]
        else:
            grads = [None] * 6
        _lib.check(L.mpg_edge_nbr_bwd(_lib.ptr(nbr), K, scale, _lib.ptr(lc), _lib.ptr(x3), ldx, _lib.ptr(m),
                                      *[_lib.ptr(t) for t in (w0, b0, w1, b1, w2, b2)], B, N, F, H0, H1, H2, ef_mode, nd,
                                      mean, alpha, p, seed, sptr, ws.data_ptr(), ws_bytes, _lib.ptr(dagg), _lib.ptr(dx), F,
                                      None, _lib.ptr(dlc), *[_lib.ptr(g) for g in grads], _lib.stream()), "mpg_edge_nbr_bwd")
        if direct:
            grads = [None] * 6
        return (dx, None, None, dlc, *grads, None, None, None, None, None)


def edge_aggregate_knn(x, mask, nbr, w0, b0, w1, b1, w2, b2, ef_mode=0, nd=0, mean=False, alpha=0.2, p_drop=0.0, lc=None):
    """``nbr`` None + ``lc``: the fully connected op with conditioning columns (fp32 kernels)."""
    return EdgeNbrFn.apply(x, mask, nbr, lc, w0, b0, w1, b1, w2, b2, ef_mode, nd, mean, alpha, p_drop)


class CondColsFn(torch.autograd.Function):
    """(x | cond[r % B]) for the node network's input (mpgan/model.py:270-276); the conditioning values are data."""

    @staticmethod
    def forward(ctx, x, cond):
        L = _lib.lib()
        x3, ldx = _rows(x)
        B, N, F = x3.shape
        c = cond.detach().contiguous().float()
        out = torch.empty(B, N, F + c.shape[1], device=x.device, dtype=torch.float32)
        _lib.check(L.mpg_cond_columns(_lib.ptr(x3), ldx, _lib.ptr(c), c.shape[1], _lib.ptr(out), B * N, F, B, _lib.stream()),
                   "mpg_cond_columns")
        ctx.F = F
        return out

    @staticmethod
    def backward(ctx, dout):
        return dout[..., :ctx.F], None


def cond_columns(x, cond):
    return CondColsFn.apply(x, cond)


# --------------------------------------------------------------------------------------------------
# generation post-processing (gen.py:126-141)
# --------------------------------------------------------------------------------------------------
def gen_postprocess(jets, out, shift, norm, maxv, use_mask: bool = True):
    """out[r, i] = ((jets[r, i] - shift[i]) / norm[i]) * maxv[i], masked particles zeroed, feature 2 clamped at 0.
    ``out`` ([rows..., nfeat], fp32) may be a pinned host tensor: the kernel writes straight into it."""
    import ctypes as C
    L = _lib.lib()
    j3, ldj = _rows(jets.detach())
    nfeat = out.shape[-1]
    rows = j3.numel() // j3.shape[-1]
    if out.numel() != rows * nfeat or not out.is_contiguous() or out.dtype != torch.float32:
        raise ValueError("gen_postprocess: out must be a contiguous fp32 tensor of shape [..., nfeat] matching jets")
    if not (out.is_cuda or out.is_pinned()):
        raise RuntimeError("gen_postprocess: out must live on the GPU or in pinned host memory")
    nan = float("nan")
    arr = lambda v: (C.c_float * nfeat)(*[nan if (v is None or v[i] is None) else float(v[i]) for i in range(nfeat)])
    _lib.check(L.mpg_gen_postprocess(_lib.ptr(j3), ldj, out.data_ptr(), nfeat, rows, nfeat, arr(shift), arr(norm),
                                     arr(maxv), int(bool(use_mask)), _lib.stream()), "mpg_gen_postprocess")
    return out


# --------------------------------------------------------------------------------------------------
# fused GAPT attention block
# --------------------------------------------------------------------------------------------------
def mab_supported(E, heads, Nq, Nk) -> bool:
    return bool(_lib.lib().mpg_mab_supported(int(E), int(heads), int(Nq), int(Nk)))


class MabFn(torch.autograd.Function):
    """One MAB (gapt/model.py:124-139 without LayerNorm) as one kernel per direction; see mpg_mab_fwd."""

    @staticmethod
    def forward(ctx, x, y, key_mask, w_in, b_in, w_out, b_out, w_ff, b_ff, heads, alpha, p_res, p_ff):
        L = _lib.lib()
        self_attn = y is None
        x3, ldx = _rows(x)
        y3, ldy = (x3, ldx) if self_attn else _rows(y)
        B, Nq, E = x3.shape
        Nk = y3.shape[1]
        dev = x.device
        km = None if key_mask is None else key_mask.reshape(B, Nk).contiguous().float()
        ws_ = [t.contiguous() for t in (w_in, b_in, w_out, b_out, w_ff, b_ff)]
        ws_bytes = L.mpg_mab_workspace_bytes(B)
        ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
        q = torch.empty(B, Nq, E, device=dev, dtype=torch.float32)
        kv = torch.empty(B, Nk, 2 * E, device=dev, dtype=torch.float32)
        o, h, f, out = (torch.empty(B, Nq, E, device=dev, dtype=torch.float32) for _ in range(4))
        seed = next_seed() if (p_res > 0 or p_ff > 0) else 0
        _lib.check(L.mpg_mab_fwd(_lib.ptr(x3), ldx, _lib.ptr(y3), ldy, _lib.ptr(km), *[_lib.ptr(t) for t in ws_], B, Nq, Nk,
                                 E, int(heads), float(alpha), float(p_res), float(p_ff), seed, _seed_ptr(), _PRECISION,
                                 ws.data_ptr(), ws_bytes, _lib.ptr(q), _lib.ptr(kv), _lib.ptr(o), _lib.ptr(h), _lib.ptr(f),
                                 _lib.ptr(out), _lib.stream()), "mpg_mab_fwd")
        ctx.save_for_backward(x3, y3 if not self_attn else None, km, q, kv, o, h, f, *ws_)
        ctx.params = (w_in, b_in, w_out, b_out, w_ff, b_ff)
        ctx.cfg = (ldx, ldy, B, Nq, Nk, E, int(heads), float(alpha), float(p_res), float(p_ff), seed, _seed_ptr(), self_attn,
                   _PRECISION)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        L = _lib.lib()
        x3, y3, km, q, kv, o, h, f, *ws_ = ctx.saved_tensors
        ldx, ldy, B, Nq, Nk, E, heads, alpha, p_res, p_ff, seed, sptr, self_attn, prec = ctx.cfg
        if self_attn:
            y3 = x3
        dev = dout.device
        dout = dout.contiguous()
        ws_bytes = L.mpg_mab_workspace_bytes(B)
        ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
        dx = torch.empty(B, Nq, E, device=dev, dtype=torch.float32)
        dy = None if self_attn else torch.empty(B, Nk, E, device=dev, dtype=torch.float32)
        want_w = _want_wgrad(ctx, 3, 9)
        sinks = [_grad_sink(p) for p in ctx.params] if (want_w and all(ctx.needs_input_grad[3:9])) else [None] * 6
        direct = all(g is not None for g in sinks)
        if direct:
            grads = sinks
        elif want_w:
            flat = torch.zeros(sum(t.numel() for t in ws_), device=dev, dtype=torch.float32)
            grads, off = [], 0
            for t in ws_:
                grads.append(flat[off:off + t.numel()].view_as(t))
                off += t.numel()
        else:
            grads = [None] * 6
        _lib.check(L.mpg_mab_bwd(_lib.ptr(x3), ldx, _lib.ptr(y3), ldy, _lib.ptr(km), *[_lib.ptr(t) for t in ws_], B, Nq, Nk,
                                 E, heads, alpha, p_res, p_ff, seed, sptr, prec, ws.data_ptr(), ws_bytes, _lib.ptr(q), _lib.ptr(kv),
                                 _lib.ptr(o), _lib.ptr(h), _lib.ptr(f), _lib.ptr(dout), _lib.ptr(dx), _lib.ptr(dy),
                                 *[_lib.ptr(g) for g in grads], _lib.stream()), "mpg_mab_bwd")
        if direct:
            grads = [None] * 6
        return (dx, dy, None, *grads, None, None, None, None)


def mab(x, y, key_mask, w_in, b_in, w_out, b_out, w_ff, b_ff, heads, alpha, p_res, p_ff):
    """``y`` None = self attention (keys / values from ``x``)."""
    return MabFn.apply(x, y, key_mask, w_in, b_in, w_out, b_out, w_ff, b_ff, heads, alpha, p_res, p_ff)
