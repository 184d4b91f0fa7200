"""GraphEncoder (reference: encoder/graph_encoder.py) on the B200 graph kernels.

Same constructor, forward signature, module tree and ``state_dict`` keys as the reference:
stem -> 12 x Seq(Grapher, FFN) with 3 Downsample stages -> 1x1 projection -> mean over nodes.
Internally every activation is channels-last, i.e. node rows (B, N, C), which is what the
k-NN / aggregation kernels and cuDNN's 1x1 convolutions both want.
"""
import torch
import torch.nn as nn
from torch.nn import Sequential as Seq

from .. import ops
from .gcn_lib.torch_vertex import Grapher, DropPath
from .gcn_lib.torch_nn import act_layer, norm_layer, MLP, BasicConv  # noqa: F401  (re-exported like the reference)

_SIZES = {
    # size: (blocks per stage, channels per stage)
    't': ([2, 2, 6, 2], [64, 128, 256, 512]),
    's': ([2, 2, 6, 2], [80, 160, 400, 640]),
    'm': ([2, 2, 16, 2], [96, 192, 384, 768]),
}
_SIZE_LARGE = ([2, 2, 18, 2], [128, 256, 512, 1024])


class Downsample(nn.Module):
    """3x3 stride-2 convolution + BN; halves the node count (reference: graph_encoder.py:16-28)."""

    def __init__(self, in_dim=3, out_dim=768):
        super().__init__()
        self.conv = nn.Sequential(
            nn.Conv2d(in_dim, out_dim, 3, stride=2, padding=1),
            nn.BatchNorm2d(out_dim),
        )

    def forward(self, x):
        if x.is_cuda:
            y = ops.downsample_rows(x, self.conv[0], self.conv[1])  # 3-tap stride-2 form (see ops.downsample_rows)
            if y is not None:
                return y
        return self.conv(x)


class ChannelConv(nn.Module):
    """1x1 convolution + BN (reference: graph_encoder.py:31-43; unused by GraphEncoder)."""

    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(in_dim, out_dim, kernel_size=1, bias=False), nn.BatchNorm2d(out_dim))

    def forward(self, x):
        return self.conv(x)


class FFN(nn.Module):
    """Two 1x1 conv + BN layers with a residual (reference: graph_encoder.py:45-67)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act='relu', drop_path=0.0):
        super().__init__()
        out_features = out_features if out_features is not None else in_features
        hidden_features = hidden_features if hidden_features is not None else in_features
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.act = act_layer(act)
        self.fc1 = Seq(nn.Conv2d(in_features, hidden_features, 1, stride=1, bias=False, padding=0),
                       nn.BatchNorm2d(hidden_features))
        self.fc2 = Seq(nn.Conv2d(hidden_features, out_features, 1, stride=1, bias=False, padding=0),
                       nn.BatchNorm2d(out_features))

    def forward(self, x):
        if isinstance(self.drop_path, nn.Identity) and isinstance(self.act, nn.ReLU):
            # BatchNorm + ReLU and BatchNorm + residual as fused ops (train mode; eval falls through to PyTorch)
            h = ops.conv_batch_norm_act(x, self.fc1[0], self.fc1[1], relu=True)
            return ops.conv_batch_norm_act(h, self.fc2[0], self.fc2[1], residual=x)
        return self.drop_path(self.fc2(self.act(self.fc1(x)))) + x


class GraphEncoder(nn.Module):
    """Point-cloud graph encoder: (B, in_channels, N) -> (B, emb_dims) (reference: graph_encoder.py:69-191).

    cfg needs 'n_mels', 'n_frames', 'peak_stride' (N = n_mels * n_frames // peak_stride).
    As in the reference every Grapher block uses the same k, dilation 1 and no drop-path, because
    its block counter never advances (graph_encoder.py:138,147-150).
    """

    def __init__(self, cfg, k=3, conv='mr', act='relu', norm='batch', bias=True, dropout=0.0, dilation=True,
                 epsilon=0.2, drop_path=0.1, size='t', emb_dims=1024, in_channels=3):
        super().__init__()
        self.blocks, self.channels = (list(v) for v in _SIZES.get(size, _SIZE_LARGE))
        self.k = int(k)
        self.act = act
        self.norm = norm
        self.bias = bias
        self.drop_path = drop_path
        self.emb_dims = emb_dims
        self.epsilon = epsilon
        self.dilation = dilation
        self.dropout = dropout
        self.num_blocks = sum(self.blocks)
        self.conv = 'mr'  # the reference ignores its `conv` argument (graph_encoder.py:123)
        stochastic = False
        n_nodes = cfg['n_mels'] * cfg['n_frames'] // cfg['peak_stride']

        num_k = [int(v.item()) for v in torch.linspace(k, k, self.num_blocks)]
        max_dilation = 128 // max(num_k)
        dpr = [v.item() for v in torch.linspace(0, drop_path, self.num_blocks)]

        self.stem = nn.Sequential(nn.Conv2d(in_channels, self.channels[0], kernel_size=1, bias=False),
                                  nn.BatchNorm2d(self.channels[0]),
                                  nn.LeakyReLU(negative_slope=0.2))

        idx = 0  # never advanced, exactly like the reference: every block is (num_k[0], dilation 1, dpr[0] = 0)
        layers = []
        for stage, reps in enumerate(self.blocks):
            if stage > 0:
                layers.append(Downsample(self.channels[stage - 1], self.channels[stage]))
                n_nodes = n_nodes // 4
            for _ in range(reps):
                layers.append(Seq(
                    Grapher(self.channels[stage], num_k[idx], min(idx // 4 + 1, max_dilation), self.conv, self.act,
                            self.norm, self.bias, stochastic, epsilon, 1, n=n_nodes, drop_path=dpr[idx],
                            relative_pos=True),
                    FFN(in_features=self.channels[stage], hidden_features=self.channels[stage] * 4,
                        out_features=self.channels[stage], act=act, drop_path=dpr[idx]),
                ))
        self.backbone = Seq(*layers)
        self.proj = nn.Conv2d(self.channels[-1], 1024, 1, bias=True)
        ops.match_grad_strides(self)  # weight gradients keep the parameters' strides (DistributedDataParallel buckets)

    def model_init(self):
        for m in self.modules():
            if isinstance(m, torch.nn.Conv2d):
                torch.nn.init.kaiming_normal_(m.weight)
                m.weight.requires_grad = True
                if m.bias is not None:
                    m.bias.data.zero_()
                    m.bias.requires_grad = True

    def forward(self, x):
        """x: (B, C, num_points) -> (B, 1024)."""
        # (B, C, N) -> logical (B, C, N, 1) stored as node rows (B, N, C) (canonical channels-last strides):
        # one transposing copy of the 8-channel input, after which every layer keeps that layout
        # (no copy at all when the fused peak extractor produced the point cloud: it writes node rows directly)
        x = ops.canonical_rows(x.unsqueeze(-1)) if x.is_cuda else x.unsqueeze(-1).contiguous(memory_format=torch.channels_last)
        x = self.stem(x)
        for block in self.backbone:
            x = block(x)
        # The reference projects every node and then averages (graph_encoder.py:186-187).  proj is a 1x1 convolution -
        # linear - so the mean over the nodes commutes with it: averaging first is the same map (differences are fp32
        # rounding of the mean, ~1e-7) and leaves a (B, C) x (C, 1024) product instead of a convolution over all B * N
        # rows, its bias add, the layout copy and the reduction over the 4x larger projected tensor (and their backwards):
        # ~3 ms of a 105 ms training step at batch 512.
        x = self.proj(ops.mean_over_nodes(x))
        return x.flatten(1)
