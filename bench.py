"""Benchmark of the MPGAN hot path (BASELINE.json metric: jets/s per G+D train step at 30 & 150 particles; gen jets/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--no-suite]
    torchrun ... bench.py --gpus N --check        # 1-rank vs N-rank gradient equality on real NCCL (prints one JSON line)

The ONE JSON line's top level is the headline workload (default `train_n30_b256`: BASELINE configs[1], the
configuration the metric is quoted on).  Unless `--no-suite` / `--workload` is given, the same run also times the
other configurations the metric names and embeds them under `"workloads"` (each with value / e2e / roofline / clocks /
gpu_launches), so one driver invocation per `--gpus N` covers N=150, generation, GAPT and the all-real variants:

  train_n30_b256[_allreal]    MPGAN G+D training step, 30-particle jets, batch 256 per GPU   (configs[1])
  train_n150_b32              150 particles, batch 32 per GPU (configs[2], the reference's default batch)
  train_n150_b256[_allreal]   150 particles, batch 256 per GPU (what the fused kernels make possible)
  train_n100_b256             100-point clouds, batch 256 (configs[4]: sparsified-MNIST shapes with masking)
  gen_n30_b1024, gen_n150_b1024   generator inference from the mp_g weights, batch 1024 per GPU (configs[0], [2])
  train_gapt_n30_b512, train_gapt_isab_n30_b512   GAPT (SAB / ISAB) training step, batch 512 (configs[3])

A "step" is one train_D + train_G on one synthetic batch (train.py:841-878, num_critic = num_gen = 1, LS loss,
RMSprop, D dropout 0.5) or one generator batch.  `value` times K steps on the device with the batch resident in HBM
(per-step CUDA events, L2 flushed between steps, max over ranks); `e2e` times the same K steps through the public API
starting from pinned HOST buffers, host->device copies and the device->host read of the result inside the timed
region.  `--impl reference` times the UNMODIFIED reference (oracle/_ref, see oracle/build_ref.py) on the host cores.
"""
import argparse
import gc
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
    "train_n30_b256_allreal": dict(kind="train", N=30, B=256, all_real=True),
    "train_n150_b32": dict(kind="train", N=150, B=32),
    "train_n150_b256": dict(kind="train", N=150, B=256),
    "train_n150_b256_allreal": dict(kind="train", N=150, B=256, all_real=True),
    # BASELINE configs[4]: sparsified-MNIST point clouds (train_mnist.py shapes: 100 points x (x, y, intensity) + mask
    # channel); same networks at num_hits = 100
    "train_n100_b256": dict(kind="train", N=100, B=256),
    "gen_n30_b1024": dict(kind="gen", N=30, B=1024),
    "gen_n150_b1024": dict(kind="gen", N=150, B=1024),
    # BASELINE configs[3]: GAPT (masked set attention) training step, reference batch 512 (setup_training.py:836-838)
    "train_gapt_n30_b512": dict(kind="train", N=30, B=512, model="gapt"),
    "train_gapt_isab_n30_b512": dict(kind="train", N=30, B=512, model="gapt", isab=True),
}
HEADLINE = "train_n30_b256"
SUITE = ["train_n30_b256_allreal", "train_n150_b32", "train_n150_b256", "train_n150_b256_allreal", "train_n100_b256",
         "gen_n30_b1024", "gen_n150_b1024", "train_gapt_n30_b512", "train_gapt_isab_n30_b512"]


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


def _dram_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of each tcgen05 edge kernel, from the `ncu --set full`
    captures summarised under profiles/ (profiles/dram_traffic.json: {workload: {kernel: bytes}}); null where no
    capture of that workload exists."""
    p = os.path.join(ROOT, "profiles", "dram_traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


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
    """Samples nvidia-smi clocks / throttle reasons (100 ms period) for the whole bench; window(t0, t1) reports the
    samples that fall inside one timed window [t0, t1] (wall clock)."""

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

    def window(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)   # let the sample covering the end of the window arrive
        ok = [(t, r) for t, r in list(self.rows) if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        win = [r for t, r in ok if t0 <= t <= t1 + 0.12]
        note = "samples inside the timed window"
        if len(win) < 2:   # window shorter than two sampling periods: use the samples taken under load just before it
            win = [r for t, r in ok if t0 - 1.5 <= t <= t1 + 0.25]
            note = "timed window < 2 sampling periods: samples from the 1.5 s of load before it to its end"
        sm = sorted(float(r[1]) for r in win)
        mx = [float(r[2]) for r in win if r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in win:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "note": note}

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                pass


# ----------------------------------------------------------------------------------------------------
# reference arms: the UNMODIFIED reference modules (oracle/_ref) on the host cores or, eager, on the GPU;
# the oracle port is the stand-in only when oracle/_ref is absent
# ----------------------------------------------------------------------------------------------------
def _golden_weights(device):
    import torch
    gold = os.path.join(ROOT, "tests", "golden")
    return (torch.load(os.path.join(gold, "mp_g_weights.pt"), map_location=device),
            torch.load(os.path.join(gold, "mp_d_seed4_weights.pt"), map_location=device))


def reference_step_time(wl, B_sample, reps, warm, device="cpu"):
    """Seconds per step of the reference's own train_D + train_G (or generator forward) on B_sample jets.
    Returns (seconds, threads, kind)."""
    import torch
    from oracle import mpgan_oracle as mo
    from oracle import ref_loader
    torch.set_num_threads(os.cpu_count() or 1)
    N, kind, gapt = wl["N"], wl["kind"], wl.get("model") == "gapt"
    g = torch.Generator().manual_seed(4)
    data, labels, _ = mo.synthetic_jets(B_sample, N, g, all_real=wl.get("all_real", False))
    data, labels = data.to(device), labels.to(device)
    sync = torch.cuda.synchronize if device != "cpu" else (lambda: None)
    if ref_loader.available():
        torch.manual_seed(4)
        weights = None if (gapt or N is None) else _golden_weights(device)
        rs = ref_loader.RefStep(N, device, "gapt" if gapt else "mpgan", wl.get("isab", False), weights)
        fn = (lambda: rs.generate(labels)) if kind == "gen" else (lambda: rs.step(data, labels))
        how = "reference"
    else:   # stand-in: the oracle port of the same step
        fn = _port_step(wl, data, labels, device, g)
        how = "port"
    times = []
    for _ in range(warm + reps):
        sync()
        t0 = time.perf_counter()
        fn()
        sync()
        times.append(time.perf_counter() - t0)
    best = min(times[warm:]) if (kind == "gen" or device != "cpu") else sum(times[warm:]) / reps
    return best, torch.get_num_threads(), how


def _port_step(wl, data, labels, device, g):
    import torch
    from oracle import mpgan_oracle as mo
    N, kind = wl["N"], wl["kind"]
    if wl.get("model") == "gapt":
        raise RuntimeError("oracle/_ref absent and no GAPT port step")
    sdG, sdD = _golden_weights(device)
    cfgG = mo.NetCfg(num_particles=N, final_activation="tanh")
    cfgD = mo.NetCfg(num_particles=N, final_activation="sigmoid", dropout_p=0.5,
                     layers=[mo.EdgeCfg(all_ef=False), mo.EdgeCfg()])
    pG = {k: v.clone().requires_grad_(True) for k, v in sdG.items()}
    pD = {k: v.clone().requires_grad_(True) for k, v in sdD.items()}
    stD, stG = {}, {}
    B = data.shape[0]

    def fn():
        nd = torch.randn(B, N, 32, device=device) * 0.2
        ng = torch.randn(B, N, 32, device=device) * 0.2
        if kind == "gen":
            with torch.no_grad():
                mo.generator(sdG, nd, labels, cfgG)
        else:
            mo.gd_step(pG, pD, cfgG, cfgD, data, labels, nd, ng, stateD=stD, stateG=stG)
    return fn


def gpu_eager_baseline(wl):
    """The reference's eager PyTorch modules on this GPU, fp32 and with TF32 matmuls allowed (like-for-like GPU
    baseline)."""
    import torch
    N, kind, gapt = wl["N"], wl["kind"], wl.get("model") == "gapt"
    if gapt:
        bs = wl["B"]
    elif kind == "train":
        bs = {30: 256, 100: 16, 150: 8}.get(N, 8)    # the N^2 x hidden tensors of larger batches do not fit comfortably
    else:
        bs = {30: 1024, 100: 64, 150: 32}.get(N, 32)
    out = {"unit": "jets/s"}
    for name, tf32 in (("fp32", False), ("tf32_allowed", True)):
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        try:
            sec, _, how = reference_step_time(wl, bs, 3, 1, device="cuda")
            out[name] = bs / sec
            out["sample"] = f"{bs} jets/step, best of 3 ({'unmodified reference modules' if how == 'reference' else 'oracle port'}, eager PyTorch on this GPU)"
        except Exception as e:  # pragma: no cover
            out[name] = None
            out["error"] = str(e)[:200]
        finally:
            torch.backends.cuda.matmul.allow_tf32 = old
            torch.cuda.empty_cache()
    return out


def cpu_sample_size(wl, budget_steps=None):
    """Jets per CPU step: the workload's own batch where a step stays near ~10 s on the host, else a bounded sample."""
    N, kind = wl["N"], wl["kind"]
    if wl.get("model") == "gapt":
        return wl["B"]
    if kind == "train":
        return {30: wl["B"], 100: 8, 150: 4}.get(N, 4)
    return {30: wl["B"], 100: 64, 150: 32}.get(N, 32)


def run_reference(args, name, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    N, kind = wl["N"], wl["kind"]
    B_sample = cpu_sample_size(wl)
    # bounded: time one probe step, then cut the step count, then the per-step sample, until the run fits ~3.5 minutes
    warm, reps = max(0, args.warmup), max(1, args.steps)
    probe_B = max(1, min(B_sample, 32))
    sec_probe, cores, how = reference_step_time(wl, probe_B, 1, 0)
    budget = 210.0
    per_step = lambda: sec_probe / probe_B * B_sample
    if per_step() * (warm + reps) > budget:   # keep the GPU arm's batch, run fewer steps of it (said in steps / warmup)
        warm = min(warm, 2)
        reps = min(reps, max(5, int(budget / per_step()) - warm))
    while B_sample > 8 and per_step() * (warm + reps) > budget:   # a step of the full batch is too long: bounded sample
        B_sample //= 2
    sec, cores, how = reference_step_time(wl, B_sample, reps, warm)
    val = B_sample / sec
    what = "unmodified reference (oracle/_ref: train.train_D + train.train_G, torch.optim.RMSprop)" if how == "reference" \
        else "oracle port of the reference fp32 PyTorch path (oracle/_ref absent)"
    line = {
        "impl": "reference", "metric": metric_name(wl), "value": val, "unit": "jets/s", "n_gpus": args.gpus,
        "steps": reps, "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": name, "particles": N, "batch_per_gpu": B_sample, "global_batch": B_sample,
                   "same_batch_as_gpu_arm": B_sample == wl["B"], "device": "host CPU", "threads": cores},
        "cpu_baseline": {"value": val, "unit": "jets/s", "cores": cores, "kind": how,
                         "sample": f"{B_sample} jets/step x {reps} steps of the same workload ({what})"},
        "e2e": {"value": val, "unit": "jets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def metric_name(wl):
    return "jets/sec per G+D train step" if wl["kind"] == "train" else "generated jets/sec"


# ----------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------
class Env:
    pass


def measure(name, wl, args, env, with_baselines):
    """Times one workload; returns its result dict (every rank runs it; rank 0's dict is the one reported)."""
    import torch
    import torch.distributed as dist
    from mpgan_b200 import _lib, ops, presets, train

    dev, rank, world = env.dev, env.rank, env.world
    N, B, kind = wl["N"], wl["B"], wl["kind"]
    all_real = wl.get("all_real", False) or args.all_real
    L = _lib.lib()
    ops.set_precision(1)
    gapt = wl.get("model") == "gapt"
    latent = 64 if gapt else 32
    torch.manual_seed(4)  # identical initial weights on every rank
    if gapt:
        G = presets.gapt_generator(num_hits=N, use_isab=wl.get("isab", False)).to(dev)
        D = presets.gapt_discriminator(num_hits=N, use_isab=wl.get("isab", False)).to(dev)
    else:
        G = presets.mp_generator(num_hits=N).to(dev)
        D = presets.mp_discriminator(num_hits=N).to(dev)
        sdG, sdD = _golden_weights(dev)
        G.load_state_dict(sdG)
        D.load_state_dict(sdD)
    # from here on every rank draws its own data, noise and dropout masks (an independent shard of the global batch)
    torch.manual_seed(4 + 1000 * rank)
    gen = torch.Generator(device=dev).manual_seed(4 + 1000 * rank)
    # weak scaling = fixed per-GPU work: every rank's shard has the same particle-count multiset (drawn from one seed),
    # its own feature values, noise and dropout masks
    data, labels, _ = train.synthetic_jets(B, N, dev, gen, all_real=all_real,
                                           count_generator=torch.Generator(device=dev).manual_seed(4))

    eager_step = None
    tr = gg = None
    tr_collective = None
    if kind == "train":
        tr = train.GANTrainer(G, D, lr_gen=1.5e-4 if gapt else 1e-5, lr_disc=0.5e-4 if gapt else 3e-5, num_particles=N,
                              latent_node_size=latent, fused_allreduce=args.fused)

        tr_collective = tr.collective if world > 1 else None

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
            one_step = eager_step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        one_step(data, labels)
    # keep the GPUs under load for ~1 s so the clock samples describe the loaded state when the timed window opens;
    # every rank runs the SAME number of extra steps (the steps contain collectives)
    torch.cuda.synchronize()
    t0 = time.time()
    one_step(data, labels)
    torch.cuda.synchronize()
    n_extra = torch.tensor([min(2000, int(args.preload_s / max(time.time() - t0, 1e-4)) + 1)], device=dev)
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
        env.flush_buf.zero_()  # evict L2 between timed steps (untimed)
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
    clocks = env.sampler.window(t_wall0, t_wall1) if rank == 0 else None
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms)
    value = B * world * args.steps / (total_ms * 1e-3)

    # ---- end-to-end through the public API from pinned host buffers -------------------------------
    h_data, h_labels = data.cpu().pin_memory(), labels.cpu().pin_memory()
    h_out = torch.empty(B, N, 4, pin_memory=True) if kind == "gen" else None
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sink = 0.0
    for i in range(args.steps):
        l = h_labels.to(dev, non_blocking=True)
        if kind == "train":
            d = h_data.to(dev, non_blocking=True)
            out = one_step(d, l)
            sink += float(out[0]) + float(out[1])   # D2H read of both losses (train.py:390-393,523)
        else:
            out = one_step(None, l)
            h_out.copy_(out, non_blocking=True)       # the generated jets land in pinned host memory (gen.py:143)
            torch.cuda.current_stream().synchronize()
            sink += float(h_out[0, 0, 0])
    e1.record()
    barrier()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_val = B * world * args.steps / (float(e2e_ms) * 1e-3)
    h2d = h_labels.numel() * 4 + (h_data.numel() * 4 if kind == "train" else 0)
    d2h = 8 if kind == "train" else h_out.numel() * 4

    # ---- per-kernel device time (roofline leg): eager steps with the library's event probes armed.
    # A queue of large GEMMs is enqueued first so the host runs ahead and the probed kernels execute
    # back to back on the device (no host-launch gaps inside the event pairs).
    prof, prof_steps = [], 3
    if not gapt:
        big = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
        ops.profile_start()
        for _ in range(prof_steps):
            for _ in range(16):
                big @ big
            eager_step(data, labels)
        prof = ops.profile_stop()
        barrier()
        del big

    res = None
    if rank == 0:
        pk = peaks()
        by = {}
        for kname, a, b, fl, fl_exec in prof:
            t = by.setdefault(kname, [0.0, 0, 0.0, 0.0])
            t[0] += a.elapsed_time(b)
            t[1] += 1
            t[2] += fl
            t[3] += fl_exec
        kernels = {k: v for k, v in by.items() if k.endswith("_kernel")}
        dom = max(kernels, key=lambda k: kernels[k][0]) if kernels else None
        roof = None
        if dom:
            ms, cnt, fl, fl_exec = by[dom]
            # `achieved` counts only the (tile, sender) steps the kernel executes (fully masked senders are skipped and
            # not credited).  Peak = the measured BURST bf16 figure: each kernel is timed alone by its own event pair,
            # for well under a second.
            ach = fl_exec / (ms * 1e-3) / 1e12
            traffic = _dram_traffic().get(name, {})
            roof = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": pk["tflops_burst"], "unit": "TFLOP/s",
                    "frac": ach / pk["tflops_burst"], "traffic": traffic.get(dom),
                    "frac_of_sustained_peak": ach / pk["tflops"], "executed_step_fraction": fl_exec / fl if fl else None,
                    "peak_source": pk["src"] + " bf16 burst (kernel timed alone by CUDA events; sustained figure: %.1f)" % pk["tflops"],
                    "launches": cnt, "avg_launch_ms": ms / cnt,
                    "share_of_step": (ms / prof_steps) / (sum(step_ms) / len(step_ms)),
                    "kernels": {k: {"ms_per_step": v[0] / prof_steps, "launches_per_step": v[1] / prof_steps,
                                    "tflops": v[3] / (v[0] * 1e-3) / 1e12, "frac": v[3] / (v[0] * 1e-3) / 1e12 / pk["tflops_burst"],
                                    "traffic": traffic.get(k)}
                                for k, v in kernels.items()}}
        if gapt:   # HBM-bound path: the roofline is stated on the whole step against the measured copy bandwidth
            gb = gapt_step_bytes(N, wl.get("isab", False))
            ach = value / world * gb / 1e9
            roof = {"bound": "hbm", "kernel": "whole G+D step (attention, projection and dropout kernels)",
                    "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"], "traffic": None,
                    "algorithmic_bytes_per_jet": gb, "peak_source": pk["src"] + " HBM copy"}
        cpu = None
        if with_baselines and world == 1:
            cpu = {}
            if with_baselines == "full":   # CPU baseline: bounded sample of the same workload on the host cores
                try:
                    bs = min(cpu_sample_size(wl), 64 if not gapt else wl["B"])
                    sec, cores, how = reference_step_time(wl, bs, 2, 1)
                    cpu = {"value": bs / sec, "unit": "jets/s", "cores": cores, "kind": how,
                           "sample": f"{bs} jets/step x 2 steps ({'unmodified reference, oracle/_ref' if how == 'reference' else 'oracle port'}; "
                                     "the --impl reference arm runs the full batch)"}
                except Exception as e:  # pragma: no cover
                    cpu = {"value": None, "unit": "jets/s", "cores": 0, "kind": "port", "sample": f"not measured: {e}"}
            try:
                cpu["gpu_eager"] = gpu_eager_baseline(wl)   # the reference's eager PyTorch code on this B200
            except Exception as e:  # pragma: no cover
                cpu["gpu_eager"] = {"error": str(e)[:200]}
        res = {
            "metric": metric_name(wl), "value": value, "unit": "jets/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "tf32" if gapt else "bf16", "data": "synthetic",
            "config": {"workload": name, "particles": N, "batch_per_gpu": B, "global_batch": B * world,
                       "particles_per_jet": "all N real" if all_real else "n ~ U{1..N} (padded rows masked); every rank's "
                                            "shard has the same particle-count multiset (fixed per-GPU work)",
                       "batch_order": "jets ordered by particle count inside each batch (GANTrainer.sort_by_count / "
                                      "train.generate); fully padded (tile, sender) steps are dropped by the kernels; the "
                                      "discriminator's edge tiles hold unmasked particles only (receiver compaction, exact: "
                                      "DESIGN.md 4.1)",
                       "l2": "flushed between timed steps (256 MiB write)",
                       "precision": ("TF32 projections, fp32 attention core" if gapt else
                                     "bf16 tcgen05 edge network, TF32 node GEMMs, fp32 accumulate"), "parallelism": f"dp{world}",
                       "cuda_graph": bool(args.graph),
                       **({"collective": tr_collective} if tr_collective else {})},
            "e2e": {"value": e2e_val, "unit": "jets/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "roofline": roof,
            "clocks": clocks,
        }
        if cpu is not None:
            res["cpu_baseline"] = cpu
    # release the graphs (they reference the NCCL communicator) before the next workload / teardown
    ops.set_device_seed(None)
    del one_step, eager_step
    if tr is not None:
        tr.release()
    tr = gg = None
    gc.collect()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist

    env = Env()
    env.rank = int(os.environ.get("RANK", "0"))
    env.world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    env.dev = torch.device("cuda", local)
    if env.world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries the ONE JSON line only
        dist.init_process_group("nccl", device_id=env.dev)
    env.flush_buf = torch.empty(256 * 1024 * 1024 // 4, device=env.dev)  # > 126 MB L2
    env.sampler = ClockSampler(local)
    if env.rank == 0:
        env.sampler.start()

    if args.check:
        line = dp_check(args, env)
    else:
        head = args.workload or HEADLINE
        line = measure(head, WORKLOADS[head], args, env, with_baselines="full" if args.baselines else None)
        if args.suite and args.workload is None:
            extra = {}
            for name in SUITE:
                try:
                    r = measure(name, WORKLOADS[name], args, env,
                                with_baselines="gpu" if args.baselines and name in ("train_n150_b256", "gen_n30_b1024", "gen_n150_b1024",
                                                                 "train_gapt_n30_b512", "train_gapt_isab_n30_b512") else None)
                except Exception as e:  # a failing secondary workload must not take the headline down with it
                    r = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
                    if env.world > 1:
                        raise
                if env.rank == 0:
                    extra[name] = r
            if env.rank == 0:
                line["workloads"] = extra
    if env.rank == 0:
        env.sampler.stop()
        print(json.dumps(line), flush=True)
    if env.world > 1:
        teardown(dist, torch)


def teardown(dist, torch):
    """Leave together: every rank stays until rank 0 has printed.  The graphs that referenced the NCCL communicator
    were released in measure(); destroy_process_group() then returns promptly.  A watchdog bounds the teardown in
    case a communicator still hangs (it did while captured graphs were alive): everything is measured and printed."""
    sys.stdout.flush()
    try:
        dist.barrier()
        torch.cuda.synchronize()
    except Exception:
        pass

    def _bail():
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)

    wd = threading.Timer(20.0, _bail)
    wd.daemon = True
    wd.start()
    try:
        dist.destroy_process_group()
    except Exception:
        pass
    wd.cancel()


def dp_check(args, env):
    """1 rank on the global batch vs `world` ranks on its shards, on real NVLink / NCCL: after one train_D + train_G
    the RMSprop accumulators (0.01 * g_avg^2 after the first step, i.e. the averaged gradients the optimizers consumed)
    and the weights must agree (setup_training.py:1418-1421 DataParallel semantics).  Both collectives are checked: the
    fused peer-memory all-reduce + RMSprop kernel and ncclAllReduce + the RMSprop kernel."""
    import torch
    import torch.distributed as dist
    from mpgan_b200 import ops, presets, train

    dev, rank, world = env.dev, env.rank, env.world
    N, Bs = 30, 64
    res = {"check": "dp_gradient_equality", "n_gpus": world, "shard_batch": Bs, "global_batch": Bs * world}
    worst = {0: 0.0, 1: 0.0}
    for prec in (0, 1):
        ops.set_precision(prec)
        g = torch.Generator(device=dev).manual_seed(77)
        data, labels, _ = train.synthetic_jets(Bs * world, N, dev, g)
        nd = train.get_gen_noise(Bs * world, N, 32, 0.2, dev, g)
        ng = train.get_gen_noise(Bs * world, N, 32, 0.2, dev, g)
        sl = slice(rank * Bs, (rank + 1) * Bs)
        out = {}
        for mode in ("single", "fused", "nccl"):
            torch.manual_seed(4)
            G = presets.mp_generator(num_hits=N).to(dev)
            D = presets.mp_discriminator(num_hits=N, disc_dropout=0.0).to(dev)
            sdG, sdD = _golden_weights(dev)
            G.load_state_dict(sdG)
            D.load_state_dict(sdD)
            tr = train.GANTrainer(G, D, num_particles=N, lr_gen=1e-4, lr_disc=3e-4, world_override=1 if mode == "single" else None,
                                  fused_allreduce=(mode == "fused"))
            w0 = (tr.fpD.flat.clone(), tr.fpG.flat.clone())
            if mode == "single":
                tr.train_D(data, labels, noise=nd)
                tr.train_G(labels, noise=ng)
            else:
                tr.train_D(data[sl], labels[sl], noise=nd[sl])
                tr.train_G(labels[sl], noise=ng[sl])
            torch.cuda.synchronize()
            out[mode] = (tr.optD.square_avg.clone(), tr.optG.square_avg.clone(), tr.fpD.flat.clone(), tr.fpG.flat.clone(), w0,
                         tr.collective)
        rel = lambda a, b: float((a - b).norm() / b.norm().clamp_min(1e-30))
        pr = {}
        for mode in ("fused", "nccl"):
            o, r = out[mode], out["single"]
            e = {"sqD": rel(o[0], r[0]), "sqG": rel(o[1], r[1]),
                 # weights: difference between the modes relative to the distance the step moved them (RMSprop's first
                 # update is ~10 lr sign(g): elements with g ~ 0 may step either way)
                 "weightsD": float((o[2] - r[2]).norm() / (r[2] - r[4][0]).norm()),
                 "weightsG": float((o[3] - r[3]).norm() / (r[3] - r[4][1]).norm()), "collective": o[5]}
            t = torch.tensor([max(e["sqD"], e["sqG"])], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e["sq_max_over_ranks"] = float(t)
            worst[prec] = max(worst[prec], float(t))
            # every rank must hold the same weights afterwards
            ref = o[2].clone()
            if world > 1:
                dist.broadcast(ref, 0)
            same = torch.tensor([float(torch.equal(ref, o[2]))], device=dev)
            if world > 1:
                dist.all_reduce(same, op=dist.ReduceOp.MIN)
            e["weights_identical_on_all_ranks"] = bool(same.item())
            pr[mode] = e
        res[f"precision{prec}"] = pr
    ops.set_precision(1)
    res["ok"] = bool(worst[0] < 2e-4 and worst[1] < 6e-2 and all(
        res[f"precision{p}"][m]["weights_identical_on_all_ranks"] for p in (0, 1) for m in ("fused", "nccl")))
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="run the training step eagerly instead of replaying the captured CUDA graph")
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="time this workload only (default: the headline train_n30_b256 plus the suite under 'workloads')")
    ap.add_argument("--no-suite", dest="suite", action="store_false", help="headline workload only")
    ap.add_argument("--all-real", action="store_true",
                    help="every jet has N real particles (no padding: the unmasked worst case of SURVEY 8d); default n ~ U{1..N}")
    ap.add_argument("--preload-s", type=float, default=1.0, help="seconds of untimed load before each timed window")
    ap.add_argument("--no-fused-allreduce", dest="fused", action="store_false",
                    help="multi-GPU: ncclAllReduce + RMSprop kernel instead of the fused peer-memory kernel")
    ap.add_argument("--no-baselines", dest="baselines", action="store_false",
                    help="skip the CPU / GPU-eager reference legs (profiling runs)")
    ap.add_argument("--check", action="store_true", help="data-parallel gradient-equality check (use under torchrun)")
    args = ap.parse_args()
    if args.impl == "reference":
        name = args.workload or HEADLINE
        run_reference(args, name, WORKLOADS[name])
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
