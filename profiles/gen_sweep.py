"""1M-jet generation sweep (BASELINE configs[2]; reference gen.py:113-143 + train.gen_multi_batch, train.py:226-282)
through train.gen_multi_batch: batches of 4096 jets, gen.py's post-processing (un-normalise, zero masked particles,
clamp, drop the mask channel) as one kernel per batch writing straight into ONE pinned host buffer.  Prints
end-to-end jets/s (wall clock around the whole sweep, including every device->host byte).

    python profiles/gen_sweep.py [N] [num_jets] [batch]
    torchrun --nproc-per-node 8 profiles/gen_sweep.py ...     # jets sharded over the ranks, no communication
"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist

from mpgan_b200 import ops, presets, train

N = int(sys.argv[1]) if len(sys.argv) > 1 else 150
total = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)   # used for the start / end barriers only
rank = dist.get_rank() if world > 1 else 0
ops.set_precision(1)
torch.manual_seed(4)
G = presets.mp_generator(num_hits=N).to(dev).eval()
gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
G.load_state_dict(torch.load(os.path.join(gold, "mp_g_weights.pt"), map_location=dev))
g = torch.Generator().manual_seed(4)
n = torch.randint(1, N + 1, (total,), generator=g)
labels = (n.float() * torch.tensor(1.0 / N)).unsqueeze(1).pin_memory()
lo, hi = train.rank_shard(total, rank, world)
train.gen_multi_batch(G, 4 * batch, batch, N, labels=labels[:4 * batch], jets="g", rank=0, world=1)   # warm-up
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
out = train.gen_multi_batch(G, total, batch, N, labels=labels, jets="g", rank=rank, world=world)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = time.perf_counter() - t0
nz = (out[:20000].abs().sum(2) > 0).sum(1)          # particles left non-zero by the mask
ok = bool((nz <= n[lo:lo + 20000]).all()) and out.shape == (hi - lo, N, 3)
if rank == 0:
    print(f"N={N}, {world} GPU(s): {total} jets in {dt:.3f} s = {total / dt:,.0f} jets/s end to end (batch {batch}, "
          f"un-normalised [jets, {N}, 3] output in pinned host memory, {out.numel() * 4 / 1e9:.2f} GB per rank); "
          f"masked particles zeroed: {ok}")
if world > 1:
    dist.destroy_process_group()
