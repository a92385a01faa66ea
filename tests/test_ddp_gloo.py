"""Data-parallel host logic on CPU (gloo, world_size 2): flat-buffer broadcast + gradient all-reduce
reproduce the single-process gradient of the global batch.  The CUDA kernels are not
involved: the per-rank gradients come from the CPU oracle, which here plays the role of 'a correct
backward' so that the sharding / averaging / optimizer-scale plumbing is what is under test."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_d_grads(sd, x, labels):
    from oracle import mpgan_oracle as mo
    cfg = mo.NetCfg(num_particles=30, final_activation="sigmoid", layers=[mo.EdgeCfg(all_ef=False), mo.EdgeCfg()])
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    out = mo.discriminator(leaf, x, labels, cfg, training=False)
    loss = ((out - 1.0) ** 2).mean()
    loss.backward()
    return {k: v.grad for k, v in leaf.items()}


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, ROOT)
    from mpgan_b200 import presets, train
    from oracle import mpgan_oracle as mo
    torch.manual_seed(100 + rank)                      # ranks start from DIFFERENT weights ...
    G, D = presets.mp_generator(), presets.mp_discriminator(disc_dropout=0.0)
    tr = train.GANTrainer(G, D)                        # ... and the trainer broadcasts rank 0's
    flats = [torch.zeros_like(tr.fpD.flat) for _ in range(world)]
    dist.all_gather(flats, tr.fpD.flat)
    assert all(torch.equal(flats[0], f) for f in flats), "weights must be identical after broadcast"
    # global batch of 8 jets, sharded 4 + 4
    g = torch.Generator().manual_seed(7)
    x, labels, _ = mo.synthetic_jets(8, 30, g)
    sd = {k: v.detach().clone() for k, v in D.state_dict().items()}
    shard = slice(rank * 4, rank * 4 + 4)
    grads = _oracle_d_grads(sd, x[shard], labels[shard])
    tr.fpD.zero_grad()
    for n, p in zip(tr.fpD.names, tr.fpD.params):
        p.grad.copy_(grads[n])
    scale = tr._allreduce(tr.fpD)                      # sum over ranks + optimizer-side 1/world
    assert scale == 1.0 / world
    if rank == 0:
        full = _oracle_d_grads(sd, x, labels)
        worst = 0.0
        for n, p in zip(tr.fpD.names, tr.fpD.params):
            ref = full[n]
            err = float((p.grad * scale - ref).abs().max()) / max(float(ref.abs().max()), 1e-12)
            worst = max(worst, err)
        torch.save({"worst": worst, "world": tr.world}, tmp)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_flat_grad_allreduce_matches_global_batch(tmp_path):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["world"] == 2
    assert res["worst"] < 1e-5, res


def test_flat_params_views_track_the_buffers():
    from mpgan_b200 import presets, train
    D = presets.mp_discriminator()
    before = {k: v.clone() for k, v in D.state_dict().items()}
    fp = train.FlatParams(D)
    assert fp.flat.numel() == 355617
    for k, v in D.state_dict().items():
        assert torch.equal(v, before[k])               # re-homing does not change values
    fp.flat.mul_(2.0)
    assert torch.equal(D.mp_layers[0].fe.net[0].weight.data, before["mp_layers.0.fe.net.0.weight"] * 2)
    p = fp.params[0]
    p.grad.add_(1.0)
    assert float(fp.grad[: p.numel()].sum()) == p.numel()
    fp.zero_grad()
    assert float(fp.grad.abs().sum()) == 0.0 and p.grad.data_ptr() == fp.grad.data_ptr()
