"""World-size-2 gloo test of the data-parallel step logic (CPU): sharding + global NT-Xent + DDP gradients
equal the single-process global-batch step."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn
import torch.nn.functional as F

from grafp_b200.simclr.distributed import global_ntxent_loss, shard_bounds
from grafp_b200.simclr.ntxent import ntxent_loss

CFG = {"tau": 0.05}


def tiny_model():
    torch.manual_seed(3)
    return nn.Sequential(nn.Linear(12, 16), nn.ELU(), nn.Linear(16, 8))


def data():
    g = torch.Generator().manual_seed(5)
    return torch.randn(6, 12, generator=g), torch.randn(6, 12, generator=g)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    model = nn.parallel.DistributedDataParallel(tiny_model())
    x_i, x_j = data()
    lo, hi = shard_bounds(x_i.shape[0], world, rank)
    z_i = F.normalize(model(x_i[lo:hi]), dim=1)
    z_j = F.normalize(model(x_j[lo:hi]), dim=1)
    loss = global_ntxent_loss(z_i, z_j, CFG)
    loss.backward()
    if rank == 0:
        torch.save({"loss": loss.detach(), "grads": [p.grad.clone() for p in model.module.parameters()]}, out)
    dist.barrier()
    dist.destroy_process_group()


def _flat_worker(rank, world, port, out):
    """The same step with grafp_b200.training.FlatGradients (what the CUDA-graphed data-parallel step uses) instead of
    DistributedDataParallel; rank 1 starts from different weights, which the constructor's broadcast must overwrite."""
    from grafp_b200.training import FlatGradients
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    model = tiny_model()
    if rank == 1:
        with torch.no_grad():
            for p in model.parameters():
                p.add_(1.0)
    flat = FlatGradients(model)
    x_i, x_j = data()
    lo, hi = shard_bounds(x_i.shape[0], world, rank)
    for _ in range(2):   # twice: the second pass must start from zeroed gradients
        flat.zero()
        z_i = F.normalize(model(x_i[lo:hi]), dim=1)
        z_j = F.normalize(model(x_j[lo:hi]), dim=1)
        loss = global_ntxent_loss(z_i, z_j, CFG)
        loss.backward()
        flat.all_reduce_mean()
    if rank == 0:
        views = all(p.grad.data_ptr() >= flat.flat.data_ptr() for p in model.parameters())
        torch.save({"loss": loss.detach(), "grads": [p.grad.clone() for p in model.parameters()], "views": views}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_all_reduce_equals_single_process_global_step(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "rank0_flat.pt")
    mp.spawn(_flat_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    model = tiny_model()
    x_i, x_j = data()
    loss = ntxent_loss(F.normalize(model(x_i), dim=1), F.normalize(model(x_j), dim=1), CFG)
    loss.backward()
    assert got["views"]
    assert torch.allclose(got["loss"], loss.detach(), rtol=1e-5, atol=1e-6)
    for g, p in zip(got["grads"], model.parameters()):
        assert torch.allclose(g, p.grad, rtol=1e-4, atol=1e-6)


def test_shard_bounds_cover_the_batch():
    for n, w in [(6, 2), (7, 3), (4096, 8), (5, 8)]:
        spans = [shard_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_two_rank_step_equals_single_process_global_step(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "rank0.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    model = tiny_model()
    x_i, x_j = data()
    loss = ntxent_loss(F.normalize(model(x_i), dim=1), F.normalize(model(x_j), dim=1), CFG)
    loss.backward()
    assert torch.allclose(got["loss"], loss.detach(), rtol=1e-5, atol=1e-6)
    for g, p in zip(got["grads"], model.parameters()):
        assert torch.allclose(g, p.grad, rtol=1e-4, atol=1e-6)
