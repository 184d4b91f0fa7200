"""World-size-2 NCCL test (needs two GPUs; skipped otherwise): the row-sharded global NT-Xent of the data-parallel step
against the replicated form (autograd all_gather + the full problem on every rank) - same loss, same embedding gradients
under the same contract (world x the rank's share, averaged by the gradient reduction that follows)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

CFG = {"tau": 0.05}


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import torch.distributed.nn.functional as dfn
    from grafp_b200.simclr.distributed import global_ntxent_loss
    from grafp_b200.simclr.ntxent import ntxent_loss
    g = torch.Generator().manual_seed(100 + rank)
    z_i0 = torch.nn.functional.normalize(torch.randn(96, 128, generator=g), dim=1).cuda()
    z_j0 = torch.nn.functional.normalize(z_i0.cpu() + 0.3 * torch.randn(96, 128, generator=g), dim=1).cuda()
    # sharded (the CUDA path of global_ntxent_loss)
    z_i, z_j = z_i0.clone().requires_grad_(True), z_j0.clone().requires_grad_(True)
    loss_s = global_ntxent_loss(z_i, z_j, CFG)
    loss_s.backward()
    # replicated reference: gather with autograd, full problem on every rank
    r_i, r_j = z_i0.clone().requires_grad_(True), z_j0.clone().requires_grad_(True)
    loss_r = ntxent_loss(torch.cat(dfn.all_gather(r_i), dim=0), torch.cat(dfn.all_gather(r_j), dim=0), CFG)
    loss_r.backward()
    torch.cuda.synchronize()
    ok = (abs(float(loss_s) - float(loss_r)) < 1e-5 * abs(float(loss_r))
          and torch.allclose(z_i.grad, r_i.grad, rtol=1e-4, atol=1e-7) and torch.allclose(z_j.grad, r_j.grad, rtol=1e-4, atol=1e-7))
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        torch.save({"ok": bool(flag.item() > 0.5), "loss_s": float(loss_s), "loss_r": float(loss_r),
                    "gerr": float((z_i.grad - r_i.grad).abs().max() / r_i.grad.abs().max())}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_global_ntxent_equals_the_replicated_form(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "nccl_rank0.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    assert got["ok"], got
