"""Device time of one data-parallel update of a 355,617-parameter network (MPGAN's D): the fused peer-memory
all-reduce + RMSprop kernel at several CTA counts vs ncclAllReduce + the RMSprop kernel.  Run under torchrun."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist

from mpgan_b200 import ops, presets, train

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
D = presets.mp_discriminator().to(dev)
pr = train.PeerReducer(D, None)
opt = train.FusedRMSprop(pr.fp, 3e-5)
pr.fp.grad.normal_()
n = pr.fp.grad.numel()


def timeit(fn, iters=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


res = {}
for mc in (False, True):
  train.PeerReducer.use_multicast = mc
  train.PeerReducer.MULTICAST_MIN_WORLD = 2
  for ctas in ((8, 16, 32, 64, 128) if mc else (16, 32, 64, 128)):
    # every CTA count needs its own flag block layout: rebuild the reducer's flags
    train.PeerReducer.CTAS = ctas
    D2 = presets.mp_discriminator().to(dev)
    p2 = train.PeerReducer(D2, None)
    o2 = train.FusedRMSprop(p2.fp, 3e-5)
    p2.fp.grad.normal_()
    if mc and p2.multicast is None:
        continue
    res[f"fused_{'multimem' if mc else 'peer_loads'}_ctas{ctas}"] = timeit(lambda: p2.step(o2))
g = torch.randn(n, device=dev)
p = torch.randn(n, device=dev)
sq = torch.zeros(n, device=dev)


def nccl_step():
    dist.all_reduce(g)
    ops.rmsprop_(p, g, sq, 3e-5, gscale=1.0 / world)


res["nccl_allreduce_plus_rmsprop"] = timeit(nccl_step)
res["rmsprop_kernel_alone"] = timeit(lambda: ops.rmsprop_(p, g, sq, 3e-5))
if rank == 0:
    print(f"world {world}, {n} fp32 parameters ({n * 4 / 1e6:.2f} MB): microseconds per update (max over ranks)")
    for k, v in res.items():
        print(f"  {k:32s} {v:8.1f} us")
dist.barrier()
dist.destroy_process_group()
