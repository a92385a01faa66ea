"""Benchmark of the MPGAN hot path (BASELINE.json metric: jets/s per G+D train step).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Workloads (config.workload):
  train_n30_b256   MPGAN G+D training step, 30-particle jets, batch 256 per GPU  (BASELINE configs[1];
                   the default: the configuration the metric is quoted on)
  train_n150_b32   same at 150 particles, batch 32 per GPU (configs[2], reference default batch)
  train_n150_b256  same at batch 256 per GPU (what the fused kernels make possible)
  train_n100_b256  100-point clouds, batch 256 (configs[4]: sparsified-MNIST shapes with masking)
  gen_n30_b1024    generator inference from the mp_g weights, batch 1024 (configs[0])
  gen_n150_b1024   the same at 150 particles
  train_gapt_*     GAPT (SAB / ISAB) training step, 30 particles, batch 512 (configs[3])

A "step" is one train_D + train_G on one synthetic batch (train.py:841-878, num_critic=num_gen=1,
LS loss, RMSprop, D dropout 0.5).  `value` times K steps on the device with the batch resident in
HBM (per-step CUDA events, L2 flushed between steps, max over ranks); `e2e` times the same K steps
through the public API starting from pinned HOST buffers, host->device copies and the device->host
read of the two losses inside the timed region.  `--impl reference` times the CPU oracle port of the
same step (oracle/mpgan_oracle.gd_step) on the host cores, on a bounded sample of the batch.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "train_n30_b256": dict(kind="train", N=30, B=256),
    "train_n150_b32": dict(kind="train", N=150, B=32),
    "train_n150_b256": dict(kind="train", N=150, B=256),
    # BASELINE configs[4]: sparsified-MNIST point clouds (train_mnist.py shapes: 100 points x (x, y, intensity) + mask
    # channel); same networks at num_hits = 100
    "train_n100_b256": dict(kind="train", N=100, B=256),
    "gen_n30_b1024": dict(kind="gen", N=30, B=1024),
    "gen_n150_b1024": dict(kind="gen", N=150, B=1024),
    # BASELINE configs[3]: GAPT (masked set attention) training step, reference batch 512 (setup_training.py:836-838)
    "train_gapt_n30_b512": dict(kind="train", N=30, B=512, model="gapt"),
    "train_gapt_isab_n30_b512": dict(kind="train", N=30, B=512, model="gapt", isab=True),
}
# GAPT is HBM/latency-bound: algorithmic bytes per jet: each MAB reads x, y and writes its output,
# (Nq + Nk + Nq) * 64 * 4 B; a G+D step costs 8 D-forward-equivalents + 4 G-forward-equivalents as for MPGAN
def gapt_step_bytes(N, isab, M=10):
    mab = lambda nq, nk: (2 * nq + nk) * 64 * 4.0
    block = (mab(M, N) + mab(N, M)) if isab else mab(N, N)
    g = 4 * block + N * (64 + 4) * 4.0
    d = 2 * block + mab(1, N) + N * (4 + 64) * 4.0
    return 8.0 * d + 4.0 * g
H = (96, 160, 192)
FN = (256, 256)
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the ncu --set full
# captures summarised under profiles/ (filled in per round; null where no capture exists for the workload)
DRAM_TRAFFIC = {}


def layer_flops(N, F, Fout):
    """Algorithmic forward FLOPs of one MPLayer per jet (first fe layer factorised)."""
    return 4.0 * N * F * H[0] + 2.0 * N * N * (H[0] * H[1] + H[1] * H[2]) + \
        2.0 * N * ((H[2] + F) * FN[0] + FN[0] * FN[1] + FN[1] * Fout)


def net_flops(N):
    fg = layer_flops(N, 32, 32) + layer_flops(N, 32, 3)
    fd = layer_flops(N, 3, 32) + layer_flops(N, 32, 32) + 2.0 * 32
    return fg, fd


def step_flops(N):
    """Minimal G+D step: D 3 fwd + 2 full bwd + 1 dX bwd, G 2 fwd + 1 bwd = 8 F_D + 4 F_G."""
    fg, fd = net_flops(N)
    return 8.0 * fd + 4.0 * fg


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=d["bf16_tflops_sustained"], tflops_burst=d["bf16_tflops"], hbm=d["hbm_gbs"], src="measured")
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0, src="fallback")


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons (100 ms period) from before the warm-up on; stop(t0, t1)
    reports the samples that fall inside the timed window [t0, t1] (wall clock)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def wait_first(self, timeout=5.0):
        t_end = time.time() + timeout
        while self.proc is not None and not self.rows and time.time() < t_end:
            time.sleep(0.05)

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)   # let the sample covering the end of the window arrive
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        ok = [(t, r) for t, r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        win = [r for t, r in ok if t0 <= t <= t1 + 0.12]
        note = "samples inside the timed window"
        if len(win) < 2:   # window shorter than two sampling periods: use every sample taken under load
            win = [r for t, r in ok if t >= t0 - 2.0]
            note = "timed window < 2 sampling periods: samples from 2 s before it to the end of the bench"
        sm = sorted(float(r[1]) for r in win)
        mx = [float(r[2]) for r in win if r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in win:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "note": note}


# ----------------------------------------------------------------------------------------------------
# CPU oracle arm
# ----------------------------------------------------------------------------------------------------
def cpu_step_time(N, B_sample, reps, warm=1, kind="train", device="cpu"):
    """Times the oracle port of the reference's eager PyTorch path (oracle/mpgan_oracle.py) on B_sample jets;
    returns seconds/step.  device="cpu": the CPU baseline; device="cuda": the same eager fp32 PyTorch code on the
    GPU (SURVEY 8d's like-for-like GPU baseline) -- a baseline leg either way, never the product path."""
    import torch
    from oracle import mpgan_oracle as mo
    torch.set_num_threads(os.cpu_count() or 1)
    gold = os.path.join(ROOT, "tests", "golden")
    sdG = torch.load(os.path.join(gold, "mp_g_weights.pt"), map_location="cpu")
    sdD = torch.load(os.path.join(gold, "mp_d_seed4_weights.pt"), map_location="cpu")
    if device != "cpu":
        return _gpu_eager_step_time(mo, sdG, sdD, N, B_sample, reps, warm, kind, device)
    cfgG = mo.NetCfg(num_particles=N, final_activation="tanh")
    cfgD = mo.NetCfg(num_particles=N, final_activation="sigmoid", dropout_p=0.5,
                     layers=[mo.EdgeCfg(all_ef=False), mo.EdgeCfg()])
    g = torch.Generator().manual_seed(4)
    data, labels, _ = mo.synthetic_jets(B_sample, N, g)
    times = []
    if kind == "gen":
        for i in range(warm + reps):
            noise = torch.randn(B_sample, N, 32, generator=g) * 0.2
            t0 = time.perf_counter()
            with torch.no_grad():
                mo.generator(sdG, noise, labels, cfgG)
            times.append(time.perf_counter() - t0)
        return min(times[warm:]), torch.get_num_threads()
    pG = {k: v.clone().requires_grad_(True) for k, v in sdG.items()}
    pD = {k: v.clone().requires_grad_(True) for k, v in sdD.items()}
    stD, stG = {}, {}
    for i in range(warm + reps):
        nd = torch.randn(B_sample, N, 32, generator=g) * 0.2
        ng = torch.randn(B_sample, N, 32, generator=g) * 0.2
        t0 = time.perf_counter()
        mo.gd_step(pG, pD, cfgG, cfgD, data, labels, nd, ng, stateD=stD, stateG=stG)
        times.append(time.perf_counter() - t0)
    return sum(times[warm:]) / reps, torch.get_num_threads()


def _gpu_eager_step_time(mo, sdG, sdD, N, B_sample, reps, warm, kind, device):
    import torch
    sdG = {k: v.to(device) for k, v in sdG.items()}
    sdD = {k: v.to(device) for k, v in sdD.items()}
    cfgG = mo.NetCfg(num_particles=N, final_activation="tanh")
    cfgD = mo.NetCfg(num_particles=N, final_activation="sigmoid", dropout_p=0.5,
                     layers=[mo.EdgeCfg(all_ef=False), mo.EdgeCfg()])
    g = torch.Generator().manual_seed(4)
    data, labels, _ = mo.synthetic_jets(B_sample, N, g)
    data, labels = data.to(device), labels.to(device)
    pG = {k: v.clone().requires_grad_(True) for k, v in sdG.items()}
    pD = {k: v.clone().requires_grad_(True) for k, v in sdD.items()}
    stD, stG = {}, {}
    times = []
    for i in range(warm + reps):
        nd = torch.randn(B_sample, N, 32, device=device) * 0.2
        ng = torch.randn(B_sample, N, 32, device=device) * 0.2
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if kind == "gen":
            with torch.no_grad():
                mo.generator(sdG, nd, labels, cfgG)
        else:
            mo.gd_step(pG, pD, cfgG, cfgD, data, labels, nd, ng, stateD=stD, stateG=stG)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    return min(times[warm:]), 0


def gpu_eager_baseline(wl, N, kind):
    """Eager fp32 PyTorch (oracle port of the reference modules) on this GPU, fp32 and TF32-allowed matmuls."""
    import torch
    if wl.get("model") == "gapt":
        return None
    bs = {30: 256, 150: 8}.get(N, 8) if kind == "train" else {30: 256, 150: 16}.get(N, 16)
    out = {"unit": "jets/s", "sample": f"{bs} jets/step, best of 3 (oracle port of the reference's eager PyTorch path on this GPU)"}
    for name, tf32 in (("fp32", False), ("tf32_allowed", True)):
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        try:
            sec, _ = cpu_step_time(N, bs, 3, 1, kind, device="cuda")
            out[name] = bs / sec
        except Exception as e:  # pragma: no cover
            out[name] = None
            out["error"] = str(e)[:200]
        finally:
            torch.backends.cuda.matmul.allow_tf32 = old
            torch.cuda.empty_cache()
    return out


def cpu_gapt_step_time(N, B_sample, reps, warm=1, isab=False):
    """Times the CPU oracle port of one GAPT train_D + train_G step (oracle/gapt_oracle.py) on B_sample jets."""
    import torch
    from mpgan_b200 import presets
    from oracle import gapt_oracle as go
    from oracle import mpgan_oracle as mo
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(4)
    sdG = {k: v.clone().requires_grad_(True) for k, v in presets.gapt_generator(num_hits=N, use_isab=isab).state_dict().items()}
    sdD = {k: v.clone().requires_grad_(True) for k, v in presets.gapt_discriminator(num_hits=N, use_isab=isab).state_dict().items()}
    cfgG = go.GaptCfg(num_particles=N, sab_layers=4, use_isab=isab)
    cfgD = go.GaptCfg(num_particles=N, sab_layers=2, use_isab=isab, dropout_p=0.5, linear_dropout_p=0.5)
    g = torch.Generator().manual_seed(4)
    data, labels, _ = mo.synthetic_jets(B_sample, N, g)
    stD, stG, times = {}, {}, []
    for i in range(warm + reps):
        nd = torch.randn(B_sample, N, 64, generator=g) * 0.2
        ng = torch.randn(B_sample, N, 64, generator=g) * 0.2
        t0 = time.perf_counter()
        real_out = go.gapt_d(sdD, data.clone(), labels, cfgD, training=True)
        fake_out = go.gapt_d(sdD, go.gapt_g(sdG, nd, labels, cfgG, training=False), labels, cfgD, training=True)
        gD = torch.autograd.grad(mo.d_loss_ls(real_out, fake_out), list(sdD.values()), allow_unused=True)
        with torch.no_grad():
            for (k, p), gr in zip(sdD.items(), gD):
                if gr is not None:
                    mo.rmsprop_step(p, gr, stD.setdefault(k, torch.zeros_like(gr)), 0.5e-4)
        fake_out = go.gapt_d(sdD, go.gapt_g(sdG, ng, labels, cfgG, training=True), labels, cfgD, training=True)
        gG = torch.autograd.grad(mo.g_loss_ls(fake_out), list(sdG.values()), allow_unused=True)
        with torch.no_grad():
            for (k, p), gr in zip(sdG.items(), gG):
                if gr is not None:
                    mo.rmsprop_step(p, gr, stG.setdefault(k, torch.zeros_like(gr)), 1.5e-4)
        times.append(time.perf_counter() - t0)
    return sum(times[warm:]) / reps, torch.get_num_threads()


def cpu_time(wl, N, B_sample, reps, warm, kind):
    if wl.get("model") == "gapt":
        return cpu_gapt_step_time(N, B_sample, reps, warm, isab=wl.get("isab", False))
    return cpu_step_time(N, B_sample, reps, warm=warm, kind=kind)


def cpu_sample_size(wl):
    N, kind = wl["N"], wl["kind"]
    if wl.get("model") == "gapt":
        return 512
    return ({30: 32, 150: 2}.get(N, 4)) if kind == "train" else ({30: 256, 150: 16}.get(N, 16))


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N, kind = wl["N"], wl["kind"]
    # bounded sample: sized so one step is ~1-2 s on a few cores
    B_sample = cpu_sample_size(wl)
    reps = max(1, min(args.steps, 8))
    sec, cores = cpu_time(wl, N, B_sample, reps, min(args.warmup, 1), kind)
    val = B_sample / sec
    line = {
        "impl": "reference", "metric": metric_name(wl), "value": val, "unit": "jets/s", "n_gpus": args.gpus,
        "steps": reps, "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "sample_jets_per_step": B_sample, "particles": N},
        "cpu_baseline": {"value": val, "unit": "jets/s", "cores": cores, "kind": "port",
                         "sample": f"{B_sample} jets/step x {reps} steps of the same workload (oracle port of the "
                                   "reference fp32 PyTorch path)"},
        "e2e": {"value": val, "unit": "jets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def metric_name(wl):
    return "jets/sec per G+D train step" if wl["kind"] == "train" else "generated jets/sec"


# ----------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------
def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    from mpgan_b200 import _lib, ops, presets, train

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries the ONE JSON line only
        dist.init_process_group("nccl", device_id=dev)
    N, B, kind = wl["N"], wl["B"], wl["kind"]
    L = _lib.lib()
    ops.set_precision(1)
    torch.manual_seed(4 + rank)

    gapt = wl.get("model") == "gapt"
    latent = 64 if gapt else 32
    torch.manual_seed(4)  # identical initial weights on every rank
    if gapt:
        G = presets.gapt_generator(num_hits=N, use_isab=wl.get("isab", False)).to(dev)
        D = presets.gapt_discriminator(num_hits=N, use_isab=wl.get("isab", False)).to(dev)
    else:
        G = presets.mp_generator(num_hits=N).to(dev)
        D = presets.mp_discriminator(num_hits=N).to(dev)
        gold = os.path.join(ROOT, "tests", "golden")
        G.load_state_dict(torch.load(os.path.join(gold, "mp_g_weights.pt"), map_location=dev))
        D.load_state_dict(torch.load(os.path.join(gold, "mp_d_seed4_weights.pt"), map_location=dev))
    gen = torch.Generator(device=dev).manual_seed(4 + rank)
    data, labels, _ = train.synthetic_jets(B, N, dev, gen, all_real=args.all_real)
    flush_buf = torch.empty(256 * 1024 * 1024 // 4, device=dev)  # > 126 MB L2

    eager_step = None
    if kind == "train":
        tr = train.GANTrainer(G, D, lr_gen=1.5e-4 if gapt else 1e-5, lr_disc=0.5e-4 if gapt else 3e-5, num_particles=N,
                              latent_node_size=latent)

        def eager_step(d, l):
            return tr.step(d, l)

        if args.graph:
            for _ in range(2):
                eager_step(data, labels)
            tr.capture(data, labels)

            def one_step(d, l):
                return tr.step_graphed(d, l)
        else:
            one_step = eager_step
    else:
        G.eval()

        def eager_step(d, l):   # the per-kernel probe leg needs eager launches
            return train.generate(G, l, N, latent, 0.2)

        if args.graph:
            gg = train.GraphedGenerator(G, B, N, latent, 0.2, label_width=labels.shape[1])   # public API, CUDA graph

            def one_step(d, l):
                return gg(l)
        else:
            def one_step(d, l):
                return train.generate(G, l, N, latent, 0.2)   # sorts by count, un-sorts the output

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        one_step(data, labels)
    # keep the GPUs under load for ~1.5 s so nvidia-smi (100 ms period, slow to start) has samples before the
    # timed window opens; every rank runs the SAME number of extra steps (the steps contain collectives)
    torch.cuda.synchronize()
    t0 = time.time()
    one_step(data, labels)
    torch.cuda.synchronize()
    n_extra = torch.tensor([min(3000, int(1.5 / max(time.time() - t0, 1e-4)) + 1)], device=dev)
    if world > 1:
        dist.all_reduce(n_extra, op=dist.ReduceOp.MAX)
    for _ in range(int(n_extra)):
        one_step(data, labels)
    barrier()

    # ---- device-resident timed region ---------------------------------------------------------------
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = L.mpg_launch_count()
    barrier()
    t_wall0 = time.time()
    for i in range(args.steps):
        flush_buf.zero_()  # evict L2 between timed steps (untimed)
        ev[i][0].record()
        one_step(data, labels)
        ev[i][1].record()
    barrier()
    t_wall1 = time.time()
    launches = L.mpg_launch_count() - launches0
    if kind == "train" and args.graph:
        launches = tr.launches_per_step * args.steps   # replayed kernels: counted once at capture
    elif kind != "train" and args.graph:
        launches = gg.launches_per_call * args.steps
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms)
    value = B * world * args.steps / (total_ms * 1e-3)

    # ---- end-to-end through the public API from pinned host buffers -------------------------------
    h_data, h_labels = data.cpu().pin_memory(), labels.cpu().pin_memory()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sink = 0.0
    for i in range(args.steps):
        d = h_data.to(dev, non_blocking=True)
        l = h_labels.to(dev, non_blocking=True)
        out = one_step(d, l)
        if kind == "train":
            sink += float(out[0]) + float(out[1])   # D2H read of both losses (train.py:390-393,523)
        else:
            sink += float(out[0, 0, 0].cpu())
    e1.record()
    barrier()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_val = B * world * args.steps / (float(e2e_ms) * 1e-3)
    h2d = h_data.numel() * 4 + h_labels.numel() * 4
    d2h = 8 if kind == "train" else 4

    # ---- per-kernel device time (roofline leg): eager steps with the library's event probes armed.
    # A queue of large GEMMs is enqueued first so the host runs ahead and the probed kernels execute
    # back to back on the device (no host-launch gaps inside the event pairs).
    prof_fn = eager_step if eager_step is not None else one_step
    big = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
    ops.profile_start()
    for _ in range(3):
        for _ in range(16):
            big @ big
        prof_fn(data, labels)
    prof = ops.profile_stop()
    barrier()
    del big

    if rank == 0:
        pk = peaks()
        # dominant kernel class by device time
        by = {}
        for name, a, b, fl, fl_exec in prof:
            t = by.setdefault(name, [0.0, 0, 0.0, 0.0])
            t[0] += a.elapsed_time(b)
            t[1] += 1
            t[2] += fl
            t[3] += fl_exec
        kernels = {k: v for k, v in by.items() if k.endswith("_kernel")} or by
        dom = max(kernels, key=lambda k: kernels[k][0]) if kernels else None
        prof_steps = 3
        roof = None
        if dom:
            ms, cnt, fl, fl_exec = by[dom]
            # `achieved` counts only the (tile, sender) steps the kernel executes (fully masked senders are
            # skipped); `achieved_dense` is SURVEY 8(d)'s dense N^2 figure over the same time
            ach = fl_exec / (ms * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": pk["tflops"], "unit": "TFLOP/s",
                    "frac": ach / pk["tflops"], "traffic": DRAM_TRAFFIC.get((args.workload, dom)),
                    "achieved_dense": fl / (ms * 1e-3) / 1e12, "executed_step_fraction": fl_exec / fl if fl else None,
                    "peak_source": pk["src"] + " bf16 sustained",
                    "launches": cnt, "avg_launch_ms": ms / cnt,
                    "share_of_step": (ms / prof_steps) / (sum(step_ms) / len(step_ms)),
                    "all": {k: {"ms_total": v[0], "launches": v[1], "tflops_executed": v[3] / (v[0] * 1e-3) / 1e12,
                                "tflops_dense": v[2] / (v[0] * 1e-3) / 1e12} for k, v in by.items()}}
        alg = step_flops(N) if kind == "train" else net_flops(N)[0]
        if gapt:   # HBM-bound path: the roofline is stated on the whole step against the measured copy bandwidth
            gb = gapt_step_bytes(N, wl.get("isab", False))
            ach = value / world * gb / 1e9
            roof = {"bound": "hbm", "kernel": "whole G+D step (attention, projection and dropout kernels)",
                    "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"], "traffic": None,
                    "algorithmic_bytes_per_jet": gb, "peak_source": pk["src"] + " HBM copy"}
        # CPU baseline: bounded sample of the same workload on the host cores
        try:
            if world > 1:   # the CPU baseline is reported at N = 1 only (the other ranks would idle behind it)
                raise RuntimeError("reported by the 1-GPU run only")
            bs = cpu_sample_size(wl)
            sec, cores = cpu_time(wl, N, bs, 2, 1, kind)
            cpu = {"value": bs / sec, "unit": "jets/s", "cores": cores, "kind": "port",
                   "sample": f"{bs} jets/step x 2 steps (oracle port of the reference fp32 PyTorch path)"}
            ge = gpu_eager_baseline(wl, N, kind)
            if ge is not None:
                cpu["gpu_eager"] = ge   # the same eager PyTorch code on this B200 (like-for-like GPU baseline)
        except Exception as e:  # pragma: no cover
            cpu = {"value": None, "unit": "jets/s", "cores": 0, "kind": "port", "sample": f"not measured: {e}"}
        line = {
            "metric": metric_name(wl), "value": value, "unit": "jets/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "tf32" if gapt else "bf16", "data": "synthetic",
            "config": {"workload": args.workload, "particles": N, "batch_per_gpu": B, "global_batch": B * world,
                       "particles_per_jet": "all N real" if args.all_real else "n ~ U{1..N} (padded rows masked)",
                       "batch_order": "jets ordered by particle count inside each batch (GANTrainer.sort_by_count / "
                                      "train.generate); fully padded (tile, sender) steps are dropped by the kernels",
                       "l2": "flushed between timed steps (256 MiB write)",
                       "precision": ("TF32 projections, fp32 attention core" if gapt else
                                     "bf16 tcgen05 edge network, TF32 node GEMMs, fp32 accumulate"), "parallelism": f"dp{world}",
                       "cuda_graph": bool(args.graph)},
            "e2e": {"value": e2e_val, "unit": "jets/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "roofline": roof,
            "step_roofline": None if gapt else {"algorithmic_gflop_per_jet": alg / 1e9,
                                                "achieved_tflops": value / world * alg / 1e12,
                                                "frac_of_peak": value / world * alg / 1e12 / pk["tflops"]},
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # every rank stays until rank 0 has printed (a rank that leaves early takes the job down with it), then:
        # NCCL communicators referenced by a captured CUDA graph do not tear down reliably
        # (destroy_process_group hung here): everything is measured and printed, so leave at once
        sys.stdout.flush()
        try:
            dist.barrier()
            torch.cuda.synchronize()
        except Exception:
            pass
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="run the training step eagerly instead of replaying the captured CUDA graph")
    ap.add_argument("--workload", default="train_n30_b256", choices=sorted(WORKLOADS))
    ap.add_argument("--all-real", action="store_true",
                    help="every jet has N real particles (no padding: the unmasked worst case of SURVEY 8d); default n ~ U{1..N}")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
