"""Seeded synthetic workloads: spectrograms, peak point clouds, weights.

Everything is generated on the CPU from an explicit ``torch.Generator`` so the same
seed gives the same bytes in the build container and on the GPU box.
"""
from __future__ import annotations

import hashlib
import math
from typing import Dict, Iterable, Tuple

import torch

DEFAULT_CFG = {
    # the keys of config/grafp.yaml that shape the hot path (grafp.yaml:17-52)
    "arch": "grafp", "n_mels": 64, "n_frames": 32, "peak_stride": 2, "n_filters": 8,
    "blur_kernel": [7, 7], "bsz_train": 256, "tau": 0.05, "lr": 8.0e-5,
    "d": 128, "h": 1024, "u": 32,
}


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def synth_spec(batch: int, seed: int, n_mels: int = 64, n_frames: int = 32) -> Tuple[torch.Tensor, torch.Tensor]:
    """Two views of ``batch`` dB-like log-mel segments, each (batch, n_mels, n_frames) fp32.

    View 1: noise floor N(-60, 5^2) plus ~40 Gaussian bumps (20-60 dB, sigma 0.5-2 bins).
    View 2: same bumps shifted by {-1, 0, 1} frames plus white noise (mirrors tr_snr 0-20 dB).
    """
    g = _gen(seed)
    n_bumps = 40
    f = torch.arange(n_mels, dtype=torch.float32).view(1, 1, n_mels, 1)
    t = torch.arange(n_frames, dtype=torch.float32).view(1, 1, 1, n_frames)
    cf = torch.rand(batch, n_bumps, 1, 1, generator=g) * (n_mels - 1)
    ct = torch.rand(batch, n_bumps, 1, 1, generator=g) * (n_frames - 1)
    amp = 20 + 40 * torch.rand(batch, n_bumps, 1, 1, generator=g)
    sf = 0.5 + 1.5 * torch.rand(batch, n_bumps, 1, 1, generator=g)
    st = 0.5 + 1.5 * torch.rand(batch, n_bumps, 1, 1, generator=g)
    shift = torch.randint(-1, 2, (batch, 1, 1, 1), generator=g).float()

    def render(dt):
        bumps = amp * torch.exp(-0.5 * (((f - cf) / sf) ** 2 + ((t - ct - dt) / st) ** 2))
        return bumps.sum(1)

    floor1 = -60 + 5 * torch.randn(batch, n_mels, n_frames, generator=g)
    floor2 = -60 + 5 * torch.randn(batch, n_mels, n_frames, generator=g)
    snr_db = 20 * torch.rand(batch, 1, 1, generator=g)
    v1 = floor1 + render(0.0)
    clean2 = render(shift)
    noise = torch.randn(batch, n_mels, n_frames, generator=g) * clean2.std(dim=(1, 2), keepdim=True) \
        * torch.pow(10.0, -snr_db / 20)
    v2 = floor2 + clean2 + noise
    return v1.contiguous(), v2.contiguous()


def synth_point_cloud(batch: int, channels: int, nodes: int, seed: int, relu: bool = False) -> torch.Tensor:
    """(batch, channels, nodes, 1) N(0,1) node features (BASELINE config 4); optional ReLU sparsity."""
    x = torch.randn(batch, channels, nodes, 1, generator=_gen(seed))
    return torch.relu(x) if relu else x


def _key_seed(seed: int, key: str) -> int:
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    return int.from_bytes(h[:8], "little") & 0x7FFFFFFFFFFFFFFF


def synth_state_dict(shapes: Dict[str, Iterable[int]], seed: int, keep: Dict[str, torch.Tensor] | None = None
                     ) -> Dict[str, torch.Tensor]:
    """Deterministic weights for a reference-keyed ``state_dict`` (any module order).

    Every tensor depends only on (seed, key name, shape): conv / linear weights are
    Kaiming-scaled normals, norm scales ~ 1, biases / running means small, running_var
    in [0.5, 1.5].  Keys listed in ``keep`` (e.g. the frozen ``relative_pos`` tables) are
    passed through unchanged.
    """
    out: Dict[str, torch.Tensor] = {}
    for key, shape in shapes.items():
        shape = tuple(shape)
        if keep is not None and key in keep:
            out[key] = keep[key].clone()
            continue
        g = _gen(_key_seed(seed, key))
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            out[key] = torch.zeros(shape, dtype=torch.int64)
        elif leaf == "running_mean":
            out[key] = 0.1 * torch.randn(shape, generator=g)
        elif leaf == "running_var":
            out[key] = 0.5 + torch.rand(shape, generator=g)
        elif leaf == "eps":
            out[key] = 0.1 * torch.randn(shape, generator=g)
        elif leaf == "bias":
            out[key] = 0.05 * torch.randn(shape, generator=g)
        elif leaf == "weight" and len(shape) == 1:
            out[key] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif leaf == "weight":
            fan_in = max(1, math.prod(shape[1:]))
            out[key] = math.sqrt(2.0 / fan_in) * torch.randn(shape, generator=g)
        else:
            out[key] = 0.1 * torch.randn(shape, generator=g)
    return out
