"""Data-parallel pieces of the training step (one process per GPU; NCCL on the box, gloo in the CPU tests).

Segments are independent, so the hot path needs no collective; what needs one is the loss, which the reference
computes over the *global* batch (its DataParallel gathers z_i, z_j to GPU 0, train.py:69-71).
"""
import torch
import torch.distributed as dist

from .ntxent import ntxent_loss


def shard_bounds(global_batch: int, world_size: int, rank: int):
    """Contiguous equal shards of the batch dimension (the remainder goes to the first ranks)."""
    base, rem = divmod(global_batch, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_ntxent_loss(z_i, z_j, cfg):
    """NT-Xent over the embeddings of all ranks.

    Autograd-aware all_gather: every rank evaluates the same global loss; the gather's backward sums the
    per-rank gradients of each slot, so after DistributedDataParallel's mean over ranks every parameter holds
    exactly the gradient of the global-batch loss (tests/test_ddp_gloo.py).
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return ntxent_loss(z_i, z_j, cfg)
    from .. import ops
    z = torch.stack((z_i, z_j), dim=1).view(2 * z_i.shape[0], z_i.shape[1])
    if ops.ntxent_supported(z):
        # CUDA: every rank evaluates only its own anchor rows against the gathered embeddings (1 / world of the work;
        # the replicated form below costs world^2 x the single-GPU loss per rank) - same value, same gradient contract
        return ops.ntxent_sharded(z, cfg['tau'])
    import torch.distributed.nn.functional as dfn
    zi_all = torch.cat(dfn.all_gather(z_i), dim=0)
    zj_all = torch.cat(dfn.all_gather(z_j), dim=0)
    return ntxent_loss(zi_all, zj_all, cfg)
