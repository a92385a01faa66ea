"""G+D training step and generation on top of the drop-in modules.

Mirrors the reference's step functions (``train.py``: ``get_gen_noise`` :100-141, ``gen`` :144-223,
``calc_D_loss`` :331-395, ``train_D`` :398-462, ``calc_G_loss`` :465-476, ``train_G`` :479-523) for
the default configuration: least-squares loss, ``gp=0``, RMSprop (setup_training.py:1511-1513),
one critic and one generator update per batch.

Differences that do not change any gradient an optimizer consumes:
  * ``train_D`` generates its fake batch under ``no_grad`` -- the reference keeps the graph and
    back-propagates D's loss into G, then discards those gradients (train.py:428-437, 495);
  * ``train_G`` does not compute D's weight gradients (they are discarded by the next zero_grad).

Data parallel: one process per GPU; parameters and gradients live in one flat fp32 buffer per
network, all-reduced (averaged) over NCCL after each backward; losses are batch means,
so the averaged gradient equals the single-process gradient of the global batch.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops


class FlatParams:
    """Re-homes a module's trainable parameters and their grads into two flat fp32 buffers.  ``grad_alloc`` lets the
    gradient buffer live in symmetric (peer-mapped) memory for the fused all-reduce + RMSprop kernel."""

    def __init__(self, module: torch.nn.Module, grad_alloc=None):
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.names = [n for n, p in module.named_parameters() if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(n, device=dev, dtype=torch.float32) if grad_alloc is None else grad_alloc(n, dev).zero_()
        off = 0
        for p in self.params:
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view_as(p)
            p.grad = self.grad[off:off + k].view_as(p)
            off += k

    def zero_grad(self):
        self.grad.zero_()
        off = 0
        for p in self.params:  # re-attach in case something replaced .grad
            k = p.numel()
            if p.grad is None or p.grad.data_ptr() != self.grad[off:off + k].data_ptr():
                p.grad = self.grad[off:off + k].view_as(p)
            off += k

    def named_grads(self):
        return {n: p.grad.detach().clone() for n, p in zip(self.names, self.params)}


class FusedRMSprop:
    """torch.optim.RMSprop(lr, alpha=0.99, eps=1e-8) on a flat buffer, one kernel per step."""

    def __init__(self, fp: FlatParams, lr: float, alpha: float = 0.99, eps: float = 1e-8):
        self.fp, self.lr, self.alpha, self.eps = fp, lr, alpha, eps
        self.square_avg = torch.zeros_like(fp.flat)

    def step(self, grad_scale: float = 1.0):
        ops.rmsprop_(self.fp.flat, self.fp.grad, self.square_avg, self.lr, self.alpha, self.eps, grad_scale)


class PeerReducer:
    """Peer-memory plumbing of ``mpg_allreduce_rmsprop``: the flat gradient buffer and a flag block of one network
    allocated in symmetric memory (torch.distributed._symmetric_memory), rendezvoused over the process group, and the
    pointer tables the kernel takes.  Construction is collective; it raises if the GPUs cannot map each other's
    memory (the trainer then keeps ncclAllReduce + the RMSprop kernel)."""

    CTAS = 64    # measured on 2 B200s (profiles/r2_bench_peer_2gpu.txt): 21 us per 1.4 MB update vs 32 us for NCCL + RMSprop

    def __init__(self, module, group):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        grp = group if group is not None else dist.group.WORLD
        handles = []

        def alloc(n, dev):
            g = symm_mem.empty(n, dtype=torch.float32, device=dev)
            handles.append(symm_mem.rendezvous(g, grp))
            return g

        self.fp = FlatParams(module, grad_alloc=alloc)
        dev = self.fp.flat.device
        words = int(_lib.lib().mpg_peer_flag_words(self.CTAS, self.world))
        self.flags = symm_mem.empty(words, dtype=torch.int32, device=dev).zero_()
        hf = symm_mem.rendezvous(self.flags, grp)
        torch.cuda.synchronize()
        dist.barrier(group)   # every rank's flags are zero before anyone signals
        self._handles = (handles[0], hf)
        self.grad_ptrs = (C.c_void_p * self.world)(*[int(p) for p in handles[0].buffer_ptrs])
        self.flag_ptrs = (C.c_void_p * self.world)(*[int(p) for p in hf.buffer_ptrs])
        # NVLink-SHARP multicast mapping of the gradient buffers (in-switch reduction): measured on 8 B200s the
        # one-shot peer-load form loses to NCCL (every rank reads 8 x 1.4 MB), so from MULTICAST_MIN_WORLD ranks on the
        # kernel sums with multimem.ld_reduce when the fabric offers it
        self.multicast = None
        try:
            mc = int(handles[0].multicast_ptr)
            if mc and self.world >= self.MULTICAST_MIN_WORLD and self.use_multicast:
                self.multicast = mc
        except Exception:
            self.multicast = None

    MULTICAST_MIN_WORLD = int(__import__('os').environ.get('MPG_MULTICAST_MIN_WORLD', '4'))
    use_multicast = True

    def step(self, opt):
        ops.allreduce_rmsprop_(self.fp.flat, opt.square_avg, self.grad_ptrs, self.flag_ptrs, self.rank, self.world,
                               self.CTAS, opt.lr, opt.alpha, opt.eps, multicast=self.multicast)


def get_gen_noise(batch_size, num_particles, latent_node_size, sd=0.2, device="cuda", generator=None):
    """Normal(0, sd) noise [B, N, latent] (train.py:113-127, ``--sd 0.2``)."""
    # one kernel (normal_ scales in place) instead of randn followed by a multiply
    return torch.empty(batch_size, num_particles, latent_node_size, device=device).normal_(0.0, sd, generator=generator)


def gradient_penalty(gp_lambda, D, real_data, generated_data, alpha=None, labels=None):
    """WGAN-GP term (reference ``gradient_penalty``, train.py:286-324): D at the interpolate
    x = alpha * real + (1 - alpha) * fake (alpha ~ U[0,1) per jet), the gradient of sum(D(x)) w.r.t. x taken with the
    graph kept (``create_graph``: the second-order kernels of ops.EdgeAggBwdFn / LinearBwdFn), and
    gp = lambda * mean((sqrt(sum_jet grad^2 + 1e-12) - 1)^2).  The reference calls D without labels (:303)."""
    B = real_data.shape[0]
    if alpha is None:
        alpha = torch.rand(B, 1, 1, device=real_data.device)
    x = (alpha * real_data + (1 - alpha) * generated_data).detach().requires_grad_(True)
    out = D(x, labels)
    with ops.input_grad_only():   # dD/dx only: the parameters' first-order gradients are not part of the penalty
        (g,) = torch.autograd.grad(out, x, grad_outputs=torch.ones_like(out), create_graph=True, retain_graph=True)
    gn = torch.sqrt(torch.sum(g.reshape(B, -1) ** 2, dim=1) + 1e-12)
    return gp_lambda * ((gn - 1) ** 2).mean()


def d_loss(loss, real_out, fake_out):
    """Critic losses of calc_D_loss (train.py:364-378) without label smoothing / noise."""
    if loss == "ls":
        return ops.ls_loss(torch.cat((real_out, fake_out), 0), real_out.shape[0], 1.0, 0.0)
    if loss == "w":
        return fake_out.mean() - real_out.mean()
    if loss == "hinge":
        return torch.relu(1.0 - real_out).mean() + torch.relu(1.0 + fake_out).mean()
    if loss == "og":
        return torch.nn.functional.binary_cross_entropy(real_out, torch.ones_like(real_out)) + \
            torch.nn.functional.binary_cross_entropy(fake_out, torch.zeros_like(fake_out))
    raise ValueError(f"unknown loss {loss!r}")


def g_loss(loss, fake_out):
    """calc_G_loss (train.py:465-476)."""
    if loss == "ls":
        return ops.ls_loss(fake_out, fake_out.shape[0], 1.0)
    if loss in ("w", "hinge"):
        return -fake_out.mean()
    if loss == "og":
        return torch.nn.functional.binary_cross_entropy(fake_out, torch.ones_like(fake_out))
    raise ValueError(f"unknown loss {loss!r}")


class GANTrainer:
    def __init__(self, G, D, lr_gen=1e-5, lr_disc=3e-5, num_particles=30, latent_node_size=32, sd=0.2,
                 process_group=None, batch_real_fake=True, sort_by_count=True, world_override=None, loss="ls", gp=0.0,
                 fused_allreduce=True):
        self.G, self.D = G, D
        # loss in {"ls", "w", "hinge", "og"} and gp = the gradient-penalty weight (train.py --loss / --gp); the losses
        # other than "ls" are a few torch reductions over the [B, 1] discriminator outputs
        self.loss, self.gp = loss, float(gp)
        # With spectral norm every D forward runs one power iteration (u, v advance; spectral_normalization.py:21-33):
        # the reference's two calls D(real), D(fake) advance them twice per train_D and use different sigmas, so one
        # batched pass would not be the same update -> keep the two calls.
        from .spectral_normalization import SpectralNorm
        has_sn = any(isinstance(m, SpectralNorm) for m in D.modules())
        self.batch_real_fake = batch_real_fake and not has_sn
        # step(): order the jets of a batch by particle count.  Jets never interact (no BatchNorm) and both losses
        # are batch means, so the update is the same; what changes is that every 128-particle tile of the edge
        # kernels then holds jets with (nearly) the same padding, and the (tile, sender) steps whose sender is
        # padded in ALL of the tile's jets -- which the kernels drop -- go from ~5-30 % to ~50 % of the steps
        # for n ~ U{1..N}.
        self.sort_by_count = sort_by_count and not (getattr(G, "order_dependent", False) or getattr(D, "order_dependent", False))
        world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        if world_override is not None:
            world = int(world_override)
        # Data parallel: the gradient all-reduce and the RMSprop update of each network are ONE kernel over NVLink
        # peer memory (PeerReducer / mpg_allreduce_rmsprop) when the GPUs can map each other's memory, else
        # ncclAllReduce followed by the RMSprop kernel.  self.collective says which.
        self.peerG = self.peerD = None
        self.collective = "none"
        if world > 1 and fused_allreduce and next(G.parameters()).is_cuda:
            try:
                self.peerG, self.peerD = PeerReducer(G, process_group), PeerReducer(D, process_group)
                self.collective = "fused peer-memory all-reduce + RMSprop (mpg_allreduce_rmsprop, " + (
                    "multimem.ld_reduce in-switch sums)" if self.peerG.multicast else "one-shot peer loads)")
            except Exception as e:   # no peer access / symmetric memory unavailable
                self.peerG = self.peerD = None
                self.collective = f"ncclAllReduce + RMSprop kernel (peer memory unavailable: {type(e).__name__})"
        elif world > 1:
            self.collective = "ncclAllReduce + RMSprop kernel"
        self.fpG = self.peerG.fp if self.peerG else FlatParams(G)
        self.fpD = self.peerD.fp if self.peerD else FlatParams(D)
        # inside train_D / train_G the kernels accumulate weight gradients straight into the flat .grad buffers
        # (ops.direct_grad, scoped to the trainer's own backward passes)
        self.optG, self.optD = FusedRMSprop(self.fpG, lr_gen), FusedRMSprop(self.fpD, lr_disc)
        self.num_particles, self.latent, self.sd = num_particles, latent_node_size, sd
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        if world_override is not None:   # e.g. 1: a single-process trainer inside a multi-rank job (bench.py --check)
            self.world = int(world_override)
        if self.world > 1:  # identical weights on every rank (reference: DataParallel replicate)
            dist.broadcast(self.fpG.flat, 0, group=self.pg)
            dist.broadcast(self.fpD.flat, 0, group=self.pg)
            # state that is not in the flat buffers: spectral-norm u / v (requires_grad=False) and module buffers
            for mod in (G, D):
                extra = [p for p in mod.parameters() if not p.requires_grad] + list(mod.buffers())
                for t in extra:
                    dist.broadcast(t.data, 0, group=self.pg)
        # step(): D's gradient all-reduce + RMSprop run on a side stream under train_G's generator forward (which
        # does not read D); train_G joins before its D forward.  Captured as a parallel branch of the step's graph.
        self._comm_stream = None
        self._comm_pending = False

    # -- helpers ---------------------------------------------------------------------------------
    def _allreduce(self, fp: FlatParams) -> float:
        """Sum gradients over ranks; returns the scale the optimizer applies (1/world = average)."""
        if self.world == 1:
            return 1.0
        dist.all_reduce(fp.grad, op=dist.ReduceOp.SUM, group=self.pg)
        return 1.0 / self.world

    def _update(self, fp, opt, peer):
        if peer is not None and self.world > 1:
            peer.step(opt)            # all-reduce + RMSprop in one kernel over peer memory
        else:
            opt.step(self._allreduce(fp))

    def named_grads(self, which):
        return (self.fpG if which == "G" else self.fpD).named_grads()

    def gen(self, batch_size, labels, noise=None):
        if noise is None:
            noise = get_gen_noise(batch_size, self.num_particles, self.latent, self.sd, labels.device)
        return self.G(noise, labels)

    # -- train.py:398-462 ------------------------------------------------------------------------
    def _join_comm(self):
        if self._comm_pending:
            torch.cuda.current_stream().wait_stream(self._comm_stream)
            self._comm_pending = False

    def train_D(self, data, labels, noise=None, overlap_update=False, gp_alpha=None):
        self._join_comm()
        self.D.train()
        self.fpD.zero_grad()
        self.G.eval()
        with torch.no_grad():
            fake = self.gen(data.shape[0], labels, noise)
        if self.batch_real_fake:
            # D(real) and D(fake) as ONE forward/backward over 2B jets: jets never interact inside D (no
            # BatchNorm; per-row dropout), so outputs and gradients equal the reference's two calls
            # (train.py:425,446) while every kernel sees twice the rows per launch
            B = data.shape[0]
            d_both = self.D(torch.cat((data, fake), 0), torch.cat((labels, labels), 0))
        else:
            B = data.shape[0]
            d_both = torch.cat((self.D(data, labels), self.D(fake, labels)), 0)
        # least squares: real -> 1, fake -> 0 (train.py:357-358, 369-370, 378), one kernel
        loss = ops.ls_loss(d_both, B, 1.0, 0.0) if self.loss == "ls" else d_loss(self.loss, d_both[:B], d_both[B:])
        with ops.direct_grad():
            loss.backward()
        if self.gp:   # train.py:380-383: the penalty's gradients add to the critic loss's
            gp = gradient_penalty(self.gp, self.D, data, fake, alpha=gp_alpha)
            gp.backward()
            loss = loss.detach() + gp.detach()
        if overlap_update and self.world > 1 and data.is_cuda:
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream(device=data.device)
            self._comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._comm_stream):
                self._update(self.fpD, self.optD, self.peerD)
            self._comm_pending = True
        else:
            self._update(self.fpD, self.optD, self.peerD)
        return loss.detach()

    # -- train.py:479-523 ------------------------------------------------------------------------
    def train_G(self, labels, noise=None, batch_size=None):
        self.G.train()
        self.fpG.zero_grad()
        fake = self.gen(batch_size or labels.shape[0], labels, noise)
        self._join_comm()   # D's update (side stream) must have landed before D runs
        for p in self.fpD.params:
            p.requires_grad_(False)
        try:
            d_fake = self.D(fake, labels)  # D stays in train mode: its dropout is active (train.py:419,494)
            loss = g_loss(self.loss, d_fake)  # train.py:467,472
            with ops.direct_grad():
                loss.backward()
        finally:
            for p in self.fpD.params:
                p.requires_grad_(True)
        self._update(self.fpG, self.optG, self.peerG)
        return loss.detach()

    def step(self, data, labels, noise_d=None, noise_g=None):
        """One critic + one generator update (train.py:841-878 with num_critic = num_gen = 1).  ``noise_d`` /
        ``noise_g``: explicit generator noise per jet (reproducible runs, parity tests); drawn when None."""
        if self.sort_by_count:
            if noise_d is not None or noise_g is not None:
                pos = ops.batch_order(labels)
                noise_d = None if noise_d is None else ops.permute_batch(noise_d, pos, 0)
                noise_g = None if noise_g is None else ops.permute_batch(noise_g, pos, 0)
                data, labels = ops.permute_batch(data, pos, 0), ops.permute_batch(labels, pos, 0)
            else:
                data, labels = sort_by_count(data, labels)
        return self.train_D(data, labels, noise_d, overlap_update=True), self.train_G(labels, noise_g)

    def state(self):
        """Clones of everything a step changes (weights and RMSprop accumulators)."""
        return [t.clone() for t in (self.fpG.flat, self.fpD.flat, self.optG.square_avg, self.optD.square_avg)]

    def load_state(self, st):
        for t, v in zip((self.fpG.flat, self.fpD.flat, self.optG.square_avg, self.optD.square_avg), st):
            t.copy_(v)

    # -- whole-step CUDA graph ---------------------------------------------------
    def capture(self, data, labels, warmup=3, explicit_noise=False, keep_state=False):
        """Captures train_D + train_G (kernels, NCCL all-reduces, RMSprop) into one CUDA graph.

        Noise is drawn inside the graph (torch's graph-safe Philox state) unless ``explicit_noise``: then
        ``step_graphed`` takes the two noise tensors as inputs.  Dropout masks change on every replay because the
        kernels add a device-resident counter, bumped in-graph, to their seeds.  The warm-up and capture passes are
        real steps on the static batch; ``keep_state`` restores weights and optimizer state afterwards.  Returns
        self; afterwards ``step_graphed`` replays the graph on new batches.
        """
        dev = data.device
        self._static_data = data.clone()
        self._static_labels = labels.clone()
        self._static_noise = None
        if explicit_noise:
            self._static_noise = tuple(get_gen_noise(data.shape[0], self.num_particles, self.latent, self.sd, dev)
                                       for _ in range(2))
        self._seed_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        ops.set_device_seed(self._seed_dev)
        saved = self.state() if keep_state else None

        def body():
            self._seed_dev.add_(0x9E3779B97F4A7C15 >> 1)
            if self._static_noise is not None:
                return self.step(self._static_data, self._static_labels, *self._static_noise)
            return self.step(self._static_data, self._static_labels)

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from . import _lib
        n0 = _lib.lib().mpg_launch_count()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._static_losses = body()
        self.launches_per_step = int(_lib.lib().mpg_launch_count() - n0)  # our kernels inside one replay
        if saved is not None:
            torch.cuda.synchronize()
            self.load_state(saved)
            self._seed_dev.zero_()
        return self

    def release(self):
        """Drops the captured graph and its static buffers (the graph references the NCCL communicator: release it
        before ``destroy_process_group``)."""
        self._graph = None
        self._static_losses = self._static_data = self._static_labels = self._static_noise = None
        ops.set_device_seed(None)

    def step_graphed(self, data=None, labels=None, noise_d=None, noise_g=None):
        """Replays the captured step; ``data``/``labels`` (host or device) are copied into the graph's
        static input buffers first (``non_blocking`` so pinned host batches stream in)."""
        if data is not None:
            self._static_data.copy_(data, non_blocking=True)
        if labels is not None:
            self._static_labels.copy_(labels, non_blocking=True)
        if noise_d is not None or noise_g is not None:
            if self._static_noise is None:
                raise RuntimeError("the step was captured with in-graph noise; capture(explicit_noise=True) to pass it")
            self._static_noise[0].copy_(noise_d, non_blocking=True)
            self._static_noise[1].copy_(noise_g, non_blocking=True)
        self._graph.replay()
        return self._static_losses


def sort_by_count(data, labels):
    """Reorders a batch by descending particle count (labels[:, -1] = n / N); see GANTrainer.sort_by_count."""
    pos = ops.batch_order(labels)                      # one kernel; pos[b] = new index of jet b
    return ops.permute_batch(data, pos, 0), ops.permute_batch(labels, pos, 0)


@torch.no_grad()
def generate(G, labels, num_particles, latent_node_size=32, sd=0.2, noise=None):
    """G(noise, labels) with the jets processed in order of particle count (fewer live edge-kernel steps, see
    GANTrainer.sort_by_count) and returned in the caller's order."""
    B = labels.shape[0]
    if getattr(G, "order_dependent", False):   # conditioning columns depend on the jet order (model.MPNet)
        if noise is None:
            noise = get_gen_noise(B, num_particles, latent_node_size, sd, labels.device)
        return G(noise, labels)
    pos = ops.batch_order(labels)
    if noise is None:
        noise = get_gen_noise(B, num_particles, latent_node_size, sd, labels.device)
    else:
        noise = ops.permute_batch(noise, pos, 0)
    out_sorted = G(noise, ops.permute_batch(labels, pos, 0))
    return ops.permute_batch(out_sorted, pos, 1)       # out[b] = out_sorted[pos[b]]


class GraphedGenerator:
    """``generate`` for a fixed batch size captured into one CUDA graph: a 30-particle batch of 1024 jets is ~25
    kernels of 3-60 us, i.e. host-launch bound when issued eagerly.  ``__call__(labels)`` copies the labels into the
    graph's static input, replays, and returns the static output buffer (valid until the next call; clone it to
    keep it).  Noise is drawn inside the graph (torch's graph-safe Philox state), so every replay is a fresh batch.
    """

    def __init__(self, G, batch_size, num_particles, latent_node_size=32, sd=0.2, label_width=1, warmup=3):
        self.G, self.B, self.N, self.latent, self.sd = G, batch_size, num_particles, latent_node_size, sd
        dev = next(G.parameters()).device
        self._labels = torch.full((batch_size, label_width), 1.0, device=dev)
        G.eval()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                generate(G, self._labels, num_particles, latent_node_size, sd)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from . import _lib
        n0 = _lib.lib().mpg_launch_count()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._out = generate(G, self._labels, num_particles, latent_node_size, sd)
        self.launches_per_call = int(_lib.lib().mpg_launch_count() - n0)

    def __call__(self, labels):
        if labels.shape != self._labels.shape:
            raise ValueError(f"GraphedGenerator was captured for labels of shape {tuple(self._labels.shape)}, "
                             f"got {tuple(labels.shape)}")
        self._labels.copy_(labels, non_blocking=True)
        self._graph.replay()
        return self._out


def synthetic_jets(B, N, device="cuda", generator=None, all_real=False, count_generator=None):
    """SURVEY 8(d) synthetic batch: features U(-.5,.5) zeroed on padded rows, 4th channel mask-0.5;
    labels n * fp32(1/N).  ``count_generator``: separate RNG for the particle counts (a multi-GPU weak-scaling bench
    seeds it identically on every rank so that each GPU's shard carries the same amount of work)."""
    if all_real:
        n = torch.full((B,), N, device=device)
    else:
        n = torch.randint(1, N + 1, (B,), device=device, generator=count_generator or generator)
    real = (torch.arange(N, device=device)[None, :] < n[:, None]).float().unsqueeze(2)
    feats = (torch.rand(B, N, 3, device=device, generator=generator) - 0.5) * real
    x = torch.cat((feats, real - 0.5), dim=2)
    labels = (n.float() * torch.tensor(1.0 / N, dtype=torch.float32, device=device)).unsqueeze(1)
    return x, labels, n


# un-normalisation constants of gen.py:10-17 (JetNet gluon / light-quark / top jets)
FEATURE_MAXES = {
    "g": [1.4532885551452637, 0.520724892616272, 0.8537549376487732, 1.0],
    "q": [1.6211985349655151, 0.4568111002445221, 0.8896132111549377, 1.0],
    "t": [1.4242753982543945, 0.4949831962585449, 0.8774275183677673, 1.0],
}
FEATURE_NORMS = [1.0, 1.0, 1.0, 1.0]
FEATURE_SHIFTS = [0.0, 0.0, -0.5, -0.5]


def rank_shard(num_samples, rank=None, world=None):
    """[start, end) of the samples rank ``rank`` of ``world`` generates: generation shards over jets with no
    communication (every rank holds the full generator)."""
    if world is None:
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank() if world > 1 else 0
    per = (num_samples + world - 1) // world
    return min(rank * per, num_samples), min((rank + 1) * per, num_samples)


@torch.no_grad()
def gen_multi_batch(G, num_samples, batch_size, num_particles, labels=None, latent_node_size=32, sd=0.2,
                    out_device="cpu", pin=True, noise=None, jets=None, mask=True, rank=None, world=None):
    """Generates ``num_samples`` jets in batches into ONE pre-allocated output (train.py:226-282
    without its O(n^2) ``torch.cat`` and without the duplicated last batch when
    ``num_samples % batch_size == 0``).

    ``jets`` in {"g", "q", "t"}: also apply gen.py's post-processing (:126-141: un-normalise, zero the masked
    particles, clamp the third feature, drop the mask channel) -- one kernel per batch writing straight into the
    (pinned) output, which is then [n, N, 3] and ready for ``np.save``.  ``noise``: explicit [num_samples, N, latent]
    noise (reproducible runs).  ``rank`` / ``world``: generate only this rank's shard (``rank_shard``); the returned
    tensor holds the shard's jets in order."""
    G.eval()
    dev = next(G.parameters()).device
    start0, end0 = (0, num_samples) if (world is None and rank is None and not (dist.is_available() and dist.is_initialized())) \
        else rank_shard(num_samples, rank, world)
    n_out = end0 - start0
    out_feats = 3 if jets is not None else G.output_node_size + 1
    out = torch.empty(n_out, num_particles, out_feats, device=out_device,
                      pin_memory=(pin and out_device == "cpu"))
    direct = jets is not None and (out.is_cuda or out.is_pinned())
    for start in range(start0, end0, batch_size):
        n = min(batch_size, end0 - start)
        lab = None if labels is None else labels[start:start + n].to(dev, non_blocking=True)
        nz = None if noise is None else noise[start:start + n].to(dev, non_blocking=True)
        if lab is None:
            batch = G(get_gen_noise(n, num_particles, latent_node_size, sd, dev) if nz is None else nz, lab)
        else:   # eager: at 4096-jet batches the launches are not the bound (a graph's capture cost loses on a 1M sweep)
            batch = generate(G, lab, num_particles, latent_node_size, sd, noise=nz)
        dst = out[start - start0:start - start0 + n]
        if jets is not None:
            tgt = dst if direct else torch.empty(n, num_particles, 3, device=dev)
            ops.gen_postprocess(batch, tgt, FEATURE_SHIFTS[:3], FEATURE_NORMS[:3], FEATURE_MAXES[jets][:3], use_mask=mask)
            if not direct:
                dst.copy_(tgt, non_blocking=True)
        else:
            dst.copy_(batch, non_blocking=True)
    if dev.type == "cuda":
        torch.cuda.current_stream().synchronize()
    return out


def save_jets(path, gen_jets):
    """gen.py:143: ``np.save(output_file, gen_jets[:, :, :3])``."""
    import numpy as np
    np.save(path, gen_jets[:, :, :3].cpu().numpy())
