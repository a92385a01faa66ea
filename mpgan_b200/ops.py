"""Autograd bindings of the C-ABI kernels (include/mpgan_b200.h).

Each op is a ``torch.autograd.Function`` whose forward/backward call straight into
``libmpgan_b200.so`` on the current CUDA stream.  PyTorch is used for device memory and autograd
bookkeeping only.  Backward passes are ``once_differentiable``: double backward (WGAN-GP,
``create_graph=True``) raises instead of silently producing wrong gradients.
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


def edge_active_fraction(mask, B, N):
    """Share of the (128-receiver tile, sender) steps the tcgen05 kernels execute: a step is dropped when
    the sender is masked in every jet the tile touches (csrc/edge_tc_common.cuh: step_list_kernel).
    Reporting helper (synchronises); not on the compute path."""
    if mask is None:
        return 1.0
    m = (mask.reshape(B, N) != 0).to(torch.int32)
    BN = B * N
    tiles = (BN + 127) // 128
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
        b = b.contiguous()
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
    @once_differentiable
    def backward(ctx, dy):
        L = _lib.lib()
        x2, w, y = ctx.saved_tensors
        ldx, M, K, N, act, alpha, p, seed, rstream, prec, sptr = ctx.cfg
        dy = dy.contiguous()
        dz = torch.empty_like(dy) if (act or p > 0) else None
        dx = torch.empty(*x2.shape, device=dy.device, dtype=torch.float32) if ctx.needs_input_grad[0] else None
        wp, bp = ctx.params
        dw_sink = _grad_sink(wp) if ctx.needs_input_grad[1] else None
        db_sink = _grad_sink(bp) if ctx.needs_input_grad[2] else None
        dw = dw_sink if dw_sink is not None else (torch.zeros_like(w) if ctx.needs_input_grad[1] else None)
        db = db_sink if db_sink is not None else (
            torch.zeros(N, device=dy.device, dtype=torch.float32) if ctx.needs_input_grad[2] else None)
        _lib.check(L.mpg_linear_bwd(_lib.ptr(dy), _lib.ptr(y), _lib.ptr(x2), ldx, _lib.ptr(w), _lib.ptr(dz),
                                    _lib.ptr(dx), K, 0, _lib.ptr(dw), _lib.ptr(db), M, K, N, act, alpha, p, seed,
                                    sptr, rstream, prec, _lib.stream()), "mpg_linear_bwd")
        return (dx, None if dw_sink is not None else dw, None if db_sink is not None else db, None, None, None, None,
                None)


def linear(x, w, b, act: bool, alpha: float, p_drop: float, rng_stream: int = 16, seed=None):
    """``seed``: dropout seed (None draws a fresh one); layers of one LinearNet call share a seed and differ
    in ``rng_stream``, which is what lets the fused node-network kernel reproduce their masks."""
    return LinearFn.apply(x, w, b, act, alpha, p_drop, rng_stream, seed)


# --------------------------------------------------------------------------------------------------
# fused edge network + aggregation
# --------------------------------------------------------------------------------------------------
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
        frac = edge_active_fraction(m, B, N) if _profile is not None else 1.0
        with _Timed("edge_fwd", edge_flops(B, N, F, H0, H1, H2),
                    [(1, "edge_tc_fwd_kernel", _pair_flops(B, N, H0, H1) + _pair_flops(B, N, H1, H2))], frac):
            _lib.check(L.mpg_edge_fwd(_lib.ptr(x3), ldx, _lib.ptr(m), *[_lib.ptr(t) for t in ws_], B, N, F, H0, H1,
                                      H2, int(ef_mode), int(nd), int(mean), float(alpha), float(p_drop), seed,
                                      _seed_ptr(), _PRECISION, ws.data_ptr(), ws_bytes, _lib.ptr(agg),
                                      _lib.stream()), "mpg_edge_fwd")
        ctx.save_for_backward(x3, m, *ws_)
        ctx.params = (w0, b0, w1, b1, w2, b2)
        ctx.fwd_ws = ws if _SAVE_EDGE_WS else None
        ctx.cfg = (ldx, B, N, F, H0, H1, H2, int(ef_mode), int(nd), int(mean), float(alpha), float(p_drop), seed,
                   _PRECISION, _seed_ptr())
        return agg

    @staticmethod
    @once_differentiable
    def backward(ctx, dagg):
        L = _lib.lib()
        x3, m, w0, b0, w1, b1, w2, b2 = ctx.saved_tensors
        ldx, B, N, F, H0, H1, H2, ef_mode, nd, mean, alpha, p, seed, prec, sptr = ctx.cfg
        dagg = dagg.contiguous()
        ws_bytes = L.mpg_edge_workspace_bytes(B, N, F, H0, H1, H2)
        ws = torch.empty(ws_bytes, device=dagg.device, dtype=torch.uint8)
        dx = torch.empty(B, N, F, device=dagg.device, dtype=torch.float32)
        sinks = [_grad_sink(p) for p in ctx.params] if all(ctx.needs_input_grad[2:8]) else [None] * 6
        direct = all(g is not None for g in sinks)
        if direct:
            grads = sinks
        elif any(ctx.needs_input_grad[2:8]):
            # one zero-filled buffer, six views (one fill kernel instead of six)
            ws_ = (w0, b0, w1, b1, w2, b2)
            flat = torch.zeros(sum(t.numel() for t in ws_), device=dagg.device, dtype=torch.float32)
            grads, off = [], 0
            for t in ws_:
                grads.append(flat[off:off + t.numel()].view_as(t))
                off += t.numel()
        else:   # frozen weights (train_G back-propagating through D): input gradient only
            grads = [None] * 6
        frac = edge_active_fraction(m, B, N) if _profile is not None else 1.0
        with _Timed("edge_bwd", 2.0 * edge_flops(B, N, F, H0, H1, H2),
                    [(2, "edge_tc_bwd_chain_kernel", 2 * _pair_flops(B, N, H0, H1) + _pair_flops(B, N, H1, H2)),
                     (3, "edge_tc_bwd_dw2_kernel", _pair_flops(B, N, H1, H2))], frac):
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
        return (dx, None, *grads, None, None, None, None, None)


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
    @once_differentiable
    def backward(ctx, dout):
        L = _lib.lib()
        a2, x2, y0, y1, w0, w1, w2 = ctx.saved_tensors
        lda, ldx, Ka, Kb, M, H1, H2, NO, alpha, p, seed, sptr = ctx.cfg
        dev = dout.device
        dout = dout.contiguous()
        ws_bytes = L.mpg_fn_workspace_bytes(Ka, Kb, H1, H2, NO)
        ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
        dz0 = torch.empty(M, H1, device=dev, dtype=torch.float32)
        dz1 = torch.empty(M, H2, device=dev, dtype=torch.float32)
        dz2 = torch.empty(M, NO, device=dev, dtype=torch.float32) if p > 0 else None
        da = torch.empty(*a2.shape, device=dev, dtype=torch.float32)
        db = torch.empty(*x2.shape, device=dev, dtype=torch.float32)
        need_w = any(ctx.needs_input_grad[2:8])
        sinks = [_grad_sink(q) for q in ctx.params] if all(ctx.needs_input_grad[2:8]) else [None] * 6
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


def split_mask(x):
    """mask = x[..., -1:] + 0.5 (fp32 multiplier) for the discriminator input."""
    L = _lib.lib()
    x3, ldx = _rows(x.detach())
    if ldx != x3.shape[-1]:
        x3 = x3.contiguous()
        ldx = x3.shape[-1]
    B, N = x3.shape[0], x3.shape[1]
    mask = torch.empty(B, N, 1, device=x.device, dtype=torch.float32)
    _lib.check(L.mpg_split_mask(_lib.ptr(x3), ldx, B * N, _lib.ptr(mask), _lib.stream()), "mpg_split_mask")
    return mask


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
    return GenTailFn.apply(h, mask, _ACT[activation])


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
    @once_differentiable
    def backward(ctx, dout):
        L = _lib.lib()
        m = ctx.saved_tensors[0] if ctx.has_mask else None
        B, N, Cc, mean = ctx.cfg
        dout = dout.contiguous()
        dh = torch.empty(B, N, Cc, device=dout.device, dtype=torch.float32)
        _lib.check(L.mpg_pool_bwd(_lib.ptr(dout), _lib.ptr(m), _lib.ptr(dh), B, N, Cc, mean, _lib.stream()),
                   "mpg_pool_bwd")
        return dh, None, None


def masked_pool(h, mask, mean: bool):
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
        r = r.contiguous()
        cols = x.shape[-1]
        rows = x.numel() // cols
        out = torch.empty_like(x)
        seed = next_seed() if p_drop > 0 else 0
        _lib.check(L.mpg_residual_dropout_fwd(_lib.ptr(x), _lib.ptr(r), _lib.ptr(out), rows, cols, float(p_drop),
                                              seed, _seed_ptr(), int(rng_stream), _lib.stream()),
                   "mpg_residual_dropout_fwd")
        ctx.cfg = (rows, cols, float(p_drop), seed, _seed_ptr(), int(rng_stream))
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        L = _lib.lib()
        rows, cols, p, seed, sptr, rstream = ctx.cfg
        dout = dout.contiguous()
        if p == 0:
            return dout, dout, None, None
        dx = torch.empty_like(dout)
        _lib.check(L.mpg_residual_dropout_bwd(_lib.ptr(dout), _lib.ptr(dx), rows, cols, p, seed, sptr, rstream,
                                              _lib.stream()), "mpg_residual_dropout_bwd")
        return dx, dx, None, None


def residual_dropout(x, r, p_drop: float, rng_stream: int = 48):
    return ResidualDropoutFn.apply(x, r, p_drop, rng_stream)
