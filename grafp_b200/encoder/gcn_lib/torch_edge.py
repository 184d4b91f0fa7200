"""Dynamic dilated k-NN graph construction (reference: encoder/gcn_lib/torch_edge.py).

``DenseDilatedKnnGraph`` / ``dense_knn_matrix`` / ``xy_dense_knn_matrix`` keep the reference
signatures and return the same ``edge_index`` (2, B, N, k) int64, but run one fused B200
kernel sequence (normalise -> tensor-core Gram -> per-row top-k -> dilation) that never writes
the N x M distance matrix to memory.  CUDA tensors only.
"""
import torch
from torch import nn

from ... import ops


class KnnTag:
    """Side information attached to an ``edge_index`` produced here (attribute ``_grafp_knn``):
    the int32 copy of the neighbour ids and the fact that ``edge_index[1]`` is arange(N), which lets
    the graph convolutions skip the centre gather."""
    __slots__ = ("nbr32",)

    def __init__(self, nbr32):
        self.nbr32 = nbr32


def _edge_index_from_knn(x, k, dilation=1, y=None, relative_pos=None, emit_all=False, normalize=True, metric="l2"):
    B, _, N, _ = x.shape
    k_out = k * dilation if emit_all else k
    edge_index = torch.empty((2, B, N, k_out), dtype=torch.int64, device=x.device)
    _, nbr32 = ops.knn_graph(x, k, dilation, y, relative_pos, emit_all=emit_all, normalize=normalize,
                             out=edge_index[0], metric=metric)
    edge_index[1] = torch.arange(N, device=x.device).view(1, N, 1)
    edge_index._grafp_knn = KnnTag(nbr32)
    return edge_index


def pairwise_distance(x):
    """(B, N, C) -> (B, N, N) squared distances, |x|^2 - 2 x x^T + |x|^2^T (reference: torch_edge.py:7-18).

    Compatibility helper for callers that want the matrix itself; the graph ops never build it.
    """
    with torch.no_grad():
        sq = (x * x).sum(dim=-1, keepdim=True)
        return sq + (-2 * torch.matmul(x, x.transpose(2, 1))) + sq.transpose(2, 1)


def part_pairwise_distance(x, start_idx=0, end_idx=1):
    """Rows [start_idx, end_idx) of :func:`pairwise_distance` (reference: torch_edge.py:21-34)."""
    with torch.no_grad():
        part = x[:, start_idx:end_idx]
        sq_part = (part * part).sum(dim=-1, keepdim=True)
        sq = (x * x).sum(dim=-1, keepdim=True)
        return sq_part + (-2 * torch.matmul(part, x.transpose(2, 1))) + sq.transpose(2, 1)


def xy_pairwise_distance(x, y):
    """(B, N, C), (B, M, C) -> (B, N, M) squared distances (reference: torch_edge.py:37-53)."""
    with torch.no_grad():
        return (x * x).sum(dim=-1, keepdim=True) + (-2 * torch.matmul(x, y.transpose(2, 1))) \
            + (y * y).sum(dim=-1, keepdim=True).transpose(2, 1)


def dense_knn_matrix(x, k=16, relative_pos=None):
    """k nearest neighbours of every node of x (B, C, N, 1), features used as given
    (reference: torch_edge.py:70-103).  Returns edge_index (2, B, N, k) int64."""
    return _edge_index_from_knn(x, k, 1, None, relative_pos, normalize=False)


def xy_dense_knn_matrix(x, y, k=16, relative_pos=None):
    """k nearest key nodes (y) of every query node (x) (reference: torch_edge.py:144-164)."""
    return _edge_index_from_knn(x, k, 1, y, relative_pos, normalize=False)


# ---- cosine variants (reference: torch_edge.py:55-68, 106-141, 166-231; unused by GraphEncoder) ----

def xy_pairwise_distance_cos(x, y):
    """The reference computes x y^T and returns an EMPTY LIST (torch_edge.py:55-68: `cos_sim = []` is what it returns);
    kept bug-compatible for callers that import it."""
    return []


def pair_cos_sim(x, y):
    """(B, N, C), (B, M, C) -> (B, N, M) inner products (reference: torch_edge.py:221-227)."""
    return torch.matmul(x, y.transpose(-2, -1))


def cos_sim_x(x):
    """(B, N, C) -> (B, N, N) inner products (reference: torch_edge.py:229-231)."""
    return torch.matmul(x, x.transpose(-2, -1))


def dense_knn_matrix_plg(x, k=16, relative_pos=None):
    """k nearest neighbours by 1 - x x^T (+ relative_pos), features used as given (reference: torch_edge.py:106-141).
    Like the reference, point clouds of more than 10 000 nodes switch to the squared Euclidean distance (:119-131)."""
    metric = "cosine" if x.shape[2] <= 10000 else "l2"
    return _edge_index_from_knn(x, k, 1, None, relative_pos, normalize=False, metric=metric)


def xy_dense_knn_matrix_plg(x, y, k=16, relative_pos=None):
    """k nearest key nodes by 1 - x y^T (+ relative_pos) (reference: torch_edge.py:166-192)."""
    return _edge_index_from_knn(x, k, 1, y, relative_pos, normalize=False, metric="cosine")


def xy_dense_knn_matrix_plg_new(x, y, k=16, relative_pos=None):
    """Same map as :func:`xy_dense_knn_matrix_plg` (reference: torch_edge.py:194-219)."""
    return _edge_index_from_knn(x, k, 1, y, relative_pos, normalize=False, metric="cosine")


class DenseDilated(nn.Module):
    """Pick the dilated neighbours out of a (2, B, N, k*d) list (reference: torch_edge.py:233-255)."""

    def __init__(self, k=9, dilation=1, stochastic=False, epsilon=0.0):
        super().__init__()
        self.dilation = dilation
        self.stochastic = stochastic
        self.epsilon = epsilon
        self.k = k

    def forward(self, edge_index):
        if self.stochastic and torch.rand(1) < self.epsilon and self.training:
            keep = torch.randperm(self.k * self.dilation)[:self.k]
            return edge_index[:, :, :, keep]
        return edge_index[:, :, :, ::self.dilation]


class DenseDilatedKnnGraph(nn.Module):
    """Dilated k-NN graph of L2-normalised node features (reference: torch_edge.py:258-284).

    forward(x (B, C, N, 1), y=None, relative_pos=None) -> edge_index (2, B, N, k) int64 with
    edge_index[0] the neighbour ids (ranks 0, d, 2d, ... of the k*d nearest, nearest first) and
    edge_index[1] the centre ids.  No parameters or buffers.
    """

    def __init__(self, k=9, dilation=1, stochastic=False, epsilon=0.0):
        super().__init__()
        self.dilation = dilation
        self.stochastic = stochastic
        self.epsilon = epsilon
        self.k = k
        self._dilated = DenseDilated(k, dilation, stochastic, epsilon)

    def forward(self, x, y=None, relative_pos=None):
        if self.stochastic:
            # the random branch needs the full k*d list; DenseDilated then draws from it
            full = _edge_index_from_knn(x, self.k, self.dilation, y, relative_pos, emit_all=True)
            return self._dilated(full)
        # deterministic: the kernel emits ranks 0, d, 2d, ... directly (== full[..., ::d])
        return _edge_index_from_knn(x, self.k, self.dilation, y, relative_pos)


class _CosineKnnGraph(nn.Module):
    """Shared body of the two cosine graph builders: L2-normalise, rank by 1 - x_hat . y_hat (+ relative_pos), dilate.
    ``sim_alpha`` / ``sim_beta`` are parameters of the reference modules that their forward never uses
    (torch_edge.py:297-298, 333-334); they are kept for state_dict compatibility."""

    def __init__(self, k=9, dilation=1, stochastic=False, epsilon=0.0):
        super().__init__()
        self.dilation = dilation
        self.stochastic = stochastic
        self.epsilon = epsilon
        self.k = k
        self._dilated = DenseDilated(k, dilation, stochastic, epsilon)
        self.sim_alpha = nn.Parameter(torch.ones(1))
        self.sim_beta = nn.Parameter(torch.zeros(1))

    def _graph(self, x, y, relative_pos):
        if self.stochastic:
            full = _edge_index_from_knn(x, self.k, self.dilation, y, relative_pos, emit_all=True, metric="cosine")
            return self._dilated(full)
        return _edge_index_from_knn(x, self.k, self.dilation, y, relative_pos, metric="cosine")


class DenseDilatedKnnGraph_plg(_CosineKnnGraph):
    """Dilated k-NN graph by cosine distance (reference: torch_edge.py:323-361)."""

    def forward(self, x, y=None, relative_pos=None):
        return self._graph(x, y, relative_pos)


class DenseDilatedKnnGraph_new(_CosineKnnGraph):
    """Cosine k-NN graph of queries x against keys y (reference: torch_edge.py:286-321; its forward normalises y
    unconditionally, so y is required there too)."""

    def forward(self, x, y=None, relative_pos=None):
        if y is None:
            raise TypeError("DenseDilatedKnnGraph_new needs the key set y (the reference normalises it unconditionally)")
        return self._graph(x, y, relative_pos)
