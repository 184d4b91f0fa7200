"""Layer factories and the neighbour gather of gcn_lib (reference: encoder/gcn_lib/torch_nn.py).

Module / parameter names are identical to the reference so checkpoints load with
``strict=True``; the convolution / norm / activation layers stay PyTorch (cuDNN), the gather
is the B200 kernel.
"""
import torch
from torch import nn
from torch.nn import Sequential as Seq, Linear as Lin, Conv2d

from ... import ops

_ACTIVATIONS = {
    'relu': lambda inplace, slope, n: nn.ReLU(inplace),
    'leakyrelu': lambda inplace, slope, n: nn.LeakyReLU(slope, inplace),
    'prelu': lambda inplace, slope, n: nn.PReLU(num_parameters=n, init=slope),
    'gelu': lambda inplace, slope, n: nn.GELU(),
    'hswish': lambda inplace, slope, n: nn.Hardswish(inplace),
}


def act_layer(act, inplace=False, neg_slope=0.2, n_prelu=1):
    """Activation by name (reference: torch_nn.py:9-25)."""
    try:
        make = _ACTIVATIONS[act.lower()]
    except KeyError:
        raise NotImplementedError('activation layer [%s] is not found' % act) from None
    return make(inplace, neg_slope, n_prelu)


def norm_layer(norm, nc):
    """2-D normalisation by name (reference: torch_nn.py:28-37)."""
    kind = norm.lower()
    if kind == 'batch':
        return nn.BatchNorm2d(nc, affine=True)
    if kind == 'instance':
        return nn.InstanceNorm2d(nc, affine=False)
    raise NotImplementedError('normalization layer [%s] is not found' % norm)


def _wanted(name):
    return name is not None and name.lower() != 'none'


class MLP(Seq):
    """Linear (+act) (+norm) stack (reference: torch_nn.py:40-49)."""

    def __init__(self, channels, act='relu', norm=None, bias=True):
        layers = []
        for c_in, c_out in zip(channels[:-1], channels[1:]):
            layers.append(Lin(c_in, c_out, bias))
            if _wanted(act):
                layers.append(act_layer(act))
            if _wanted(norm):
                layers.append(norm_layer(norm, channels[-1]))
        super().__init__(*layers)


class BasicConv(Seq):
    """Grouped (4) 1x1 Conv2d (+norm) (+act) (+dropout) stack (reference: torch_nn.py:52-76)."""

    def __init__(self, channels, act='relu', norm=None, bias=True, drop=0.):
        layers = []
        for c_in, c_out in zip(channels[:-1], channels[1:]):
            layers.append(Conv2d(c_in, c_out, 1, bias=bias, groups=4))
            if _wanted(norm):
                layers.append(norm_layer(norm, channels[-1]))
            if _wanted(act):
                layers.append(act_layer(act))
            if drop > 0:
                layers.append(nn.Dropout2d(drop))
        super().__init__(*layers)
        self.reset_parameters()

    def reset_parameters(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, (nn.BatchNorm2d, nn.InstanceNorm2d)) and m.weight is not None:
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)


def batched_index_select(x, idx):
    r"""Fetch neighbour features (reference: torch_nn.py:79-98).

    x: (B, C, M, 1), idx: (B, N, k) -> (B, C, N, k) with out[b, c, n, j] = x[b, c, idx[b, n, j]].
    Runs the coalesced row-gather kernel (backward: scatter-add); the result is channels-last.
    """
    return ops.gather_neighbors(x, idx)
