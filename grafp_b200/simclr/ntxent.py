"""NT-Xent loss (reference: simclr/ntxent.py:4-29) without the per-row Python loop."""
import torch
import torch.nn.functional as F

from .. import ops


def ntxent_loss(z_i, z_j, cfg):
    """z_i, z_j: (B, d) L2-normalised embeddings of the two views; cfg['tau'] is the temperature.

    Rows are interleaved (i0, j0, i1, j1, ...); each row's positive is its partner and its
    negatives are all other rows; the loss is the mean negative log-softmax of the positive.
    On CUDA fp32 embeddings the loss and its gradient are two fused kernels (ops.ntxent) that never build
    the (2B, 2B) similarity matrix; CPU tensors (the gloo tests of the data-parallel logic) take the
    equivalent loop-free PyTorch expression.
    """
    n2 = 2 * z_i.shape[0]
    z = torch.stack((z_i, z_j), dim=1).view(n2, z_i.shape[1])
    if ops.ntxent_supported(z):
        return ops.ntxent(z, cfg['tau'])
    logits = torch.matmul(z, z.T) / cfg['tau']
    logits = logits.masked_fill(torch.eye(n2, dtype=torch.bool, device=z.device), float('-inf'))
    partner = torch.arange(n2, device=z.device) ^ 1
    return F.cross_entropy(logits, partner, reduction='sum') / n2
