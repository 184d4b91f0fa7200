"""Helpers to read the committed golden fixtures (tests/golden/*.npz)."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def shapes_from(gold):
    return {str(k): tuple(int(s) for s in sh.split(",") if s) for k, sh in zip(gold["keys"], gold["shapes"])}


def rel_err(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def close(a, b, rel=1e-4, scale=0.0):
    """||a - b|| <= rel * max(||b||, 0.1 * scale).

    ``scale`` (the largest gradient norm of the module under test) gives a floor for gradients that
    are mathematically zero - e.g. a conv bias feeding a train-mode BatchNorm - where both sides hold
    nothing but rounding noise proportional to the other gradients.
    """
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm()) <= rel * max(float(b.norm()), 0.1 * scale)
