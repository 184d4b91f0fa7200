"""Graph convolutions of gcn_lib (reference: encoder/gcn_lib/torch_vertex.py) on the B200 kernels.

Class names, constructor / forward signatures and parameter names are the reference's.  The
gather / subtract / max / concat chains are single fused kernels (forward and backward); the
1x1 convolutions, norms and activations stay PyTorch.  Activations flow channels-last.
"""
import numpy as np
import torch
from torch import nn
import torch.nn.functional as F

from ... import ops
from .torch_nn import BasicConv, batched_index_select, act_layer
from .torch_edge import DenseDilatedKnnGraph
from .pos_embed import get_2d_relative_pos_embed


class DropPath(nn.Module):
    """Stochastic depth per sample (the timm layer the reference imports, torch_vertex.py:8)."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x * mask / keep


def _split_edge_index(edge_index):
    """-> (neighbour ids, centre ids or None).  Graphs built by DenseDilatedKnnGraph carry a tag that
    says the centre row is arange(N) and holds an int32 copy of the neighbour ids."""
    tag = getattr(edge_index, "_grafp_knn", None)
    if tag is not None:
        return (tag.nbr32 if tag.nbr32 is not None else edge_index[0]), None
    return edge_index[0], edge_index[1]


def _basic_conv(seq, x):
    """Run a BasicConv stack; a trailing [Conv2d, BatchNorm2d, ReLU] goes through the fused BatchNorm + ReLU op."""
    if len(seq) == 3 and isinstance(seq[0], nn.Conv2d) and isinstance(seq[1], nn.BatchNorm2d) \
            and isinstance(seq[2], nn.ReLU):
        return ops.conv_batch_norm_act(x, seq[0], seq[1], relu=True)
    return seq(x)


class MRConv2d(nn.Module):
    """Max-relative graph convolution (reference: torch_vertex.py:11-34)."""

    def __init__(self, in_channels, out_channels, act='relu', norm=None, bias=True):
        super().__init__()
        self.nn = BasicConv([in_channels * 2, out_channels], act, norm, bias)

    def forward(self, x, edge_index, y=None):
        nbr, ctr = _split_edge_index(edge_index)
        return _basic_conv(self.nn, ops.mr_aggregate(x, nbr, y, ctr))


class EdgeConv2d(nn.Module):
    """Edge convolution: max over neighbours of nn([x_i, x_j - x_i]) (reference: torch_vertex.py:37-52)."""

    def __init__(self, in_channels, out_channels, act='relu', norm=None, bias=True):
        super().__init__()
        self.nn = BasicConv([in_channels * 2, out_channels], act, norm, bias)

    def forward(self, x, edge_index, y=None):
        nbr, ctr = _split_edge_index(edge_index)
        return ops.max_over_k(self.nn(ops.edge_features(x, nbr, y, ctr)))


class GraphSAGE(nn.Module):
    """GraphSAGE convolution (reference: torch_vertex.py:55-70)."""

    def __init__(self, in_channels, out_channels, act='relu', norm=None, bias=True):
        super().__init__()
        self.nn1 = BasicConv([in_channels, in_channels], act, norm, bias)
        self.nn2 = BasicConv([in_channels * 2, out_channels], act, norm, bias)

    def forward(self, x, edge_index, y=None):
        nbr, _ = _split_edge_index(edge_index)
        x_j = batched_index_select(x if y is None else y, nbr)
        x_j = ops.max_over_k(self.nn1(x_j))
        return self.nn2(torch.cat([x, x_j], dim=1))


class GINConv2d(nn.Module):
    """GIN convolution (reference: torch_vertex.py:73-89)."""

    def __init__(self, in_channels, out_channels, act='relu', norm=None, bias=True):
        super().__init__()
        self.nn = BasicConv([in_channels, out_channels], act, norm, bias)
        self.eps = nn.Parameter(torch.Tensor([0.0]))

    def forward(self, x, edge_index, y=None):
        nbr, _ = _split_edge_index(edge_index)
        x_j = ops.neighbor_sum(x if y is None else y, nbr)   # gather + sum over the neighbour axis, fused
        return self.nn((1 + self.eps) * x + x_j)


_CONVS = {'edge': EdgeConv2d, 'mr': MRConv2d, 'sage': GraphSAGE, 'gin': GINConv2d}


class GraphConv2d(nn.Module):
    """Static graph convolution dispatcher (reference: torch_vertex.py:92-111)."""

    def __init__(self, in_channels, out_channels, conv='edge', act='relu', norm=None, bias=True):
        super().__init__()
        if conv not in _CONVS:
            raise NotImplementedError('conv:{} is not supported'.format(conv))
        self.gconv = _CONVS[conv](in_channels, out_channels, act, norm, bias)

    def forward(self, x, edge_index, y=None):
        return self.gconv(x, edge_index, y)


class DyGraphConv2d(GraphConv2d):
    """Dynamic graph convolution: build the k-NN graph of the input, then convolve on it
    (reference: torch_vertex.py:114-139)."""

    def __init__(self, in_channels, out_channels, kernel_size=9, dilation=1, conv='edge', act='relu',
                 norm=None, bias=True, stochastic=False, epsilon=0.0, r=1):
        super().__init__(in_channels, out_channels, conv, act, norm, bias)
        self.k = kernel_size
        self.d = dilation
        self.r = r
        self.dilated_knn_graph = DenseDilatedKnnGraph(kernel_size, dilation, stochastic, epsilon)

    def forward(self, x, relative_pos=None):
        B, C, H, W = x.shape
        y = None
        if self.r > 1:
            y = F.avg_pool2d(x, self.r, self.r)
            y = ops.as_rows(y.reshape(B, C, -1, 1))
        # (B, C, H*W, 1) node list; as_rows is a no-op for channels-last activations
        x = ops.as_rows(x.reshape(B, C, -1, 1))
        edge_index = self.dilated_knn_graph(x, y, relative_pos)
        x = super().forward(x, edge_index, y)
        return x.reshape(B, -1, H, W)


class Grapher(nn.Module):
    """fc1 -> dynamic graph convolution -> fc2 with a residual (reference: torch_vertex.py:142-194)."""

    def __init__(self, in_channels, kernel_size=9, dilation=1, conv='edge', act='relu', norm=None,
                 bias=True, stochastic=False, epsilon=0.0, r=1, n=196, drop_path=0.0, relative_pos=False):
        super().__init__()
        self.channels = in_channels
        self.n = n
        self.r = r
        self.fc1 = nn.Sequential(
            nn.Conv2d(in_channels, in_channels, 1, stride=1, padding=0),
            nn.BatchNorm2d(in_channels),
        )
        self.graph_conv = DyGraphConv2d(in_channels, in_channels * 2, kernel_size, dilation, conv,
                                        act, norm, bias, stochastic, epsilon, r)
        self.fc2 = nn.Sequential(
            nn.Conv2d(in_channels * 2, in_channels, 1, stride=1, padding=0),
            nn.BatchNorm2d(in_channels),
        )
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.relative_pos = None
        if relative_pos:
            # frozen table, present in the state_dict but unused by forward (torch_vertex.py:165-172,190)
            table = torch.from_numpy(np.float32(get_2d_relative_pos_embed(in_channels, int(n ** 0.5))))
            table = F.interpolate(table[None, None], size=(n, n // (r * r)), mode='bicubic', align_corners=False)
            self.relative_pos = nn.Parameter(-table.squeeze(1), requires_grad=False)

    def _get_relative_pos(self, relative_pos, H, W):
        if relative_pos is None or H * W == self.n:
            return relative_pos
        N = H * W
        return F.interpolate(relative_pos.unsqueeze(0), size=(N, N // (self.r * self.r)), mode="bicubic").squeeze(0)

    def forward(self, x):
        shortcut = x
        x = ops.conv_batch_norm_act(x, self.fc1[0], self.fc1[1])
        x = self.graph_conv(x, relative_pos=None)
        if isinstance(self.drop_path, nn.Identity):
            return ops.conv_batch_norm_act(x, self.fc2[0], self.fc2[1], residual=shortcut)  # BatchNorm + residual fused
        x = self.fc2(x)
        return self.drop_path(x) + shortcut
