"""The UNMODIFIED upstream GraFP modules, staged under the git-ignored ``baseline/_ref`` (TEST INFRASTRUCTURE).

Used by ``bench.py`` only: the ``--impl reference`` arm / ``cpu_baseline`` (the reference's own PyTorch code on the
host cores) and ``gpu_eager_baseline`` (the same code run eager on the B200 - the "kernel to beat" of SURVEY.md
section 8d).  Nothing under ``grafp_b200/`` imports this.

``stage()`` copies the few upstream files the training step needs from ``/root/reference`` (build container only;
``__graft_entry__.build()`` calls it).  ``baseline/_ref`` is git-ignored - no upstream source enters the history -
but not gpurun-ignored, so it travels to the GPU box, where ``/root/reference`` does not exist.  Three third-party
modules the upstream files import but never use on this path (timm, torchmetrics, librosa) are stubbed, exactly as in
``tests/_reference_import.py``.
"""
import os
import shutil
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
UPSTREAM = os.environ.get("GRAFP_REFERENCE_ROOT", "/root/reference")
_FILES = ["encoder", "simclr", "peak_extractor.py", os.path.join("config", "grafp.yaml")]


def stage(force: bool = False) -> bool:
    """Copy the upstream files into baseline/_ref; returns True when the staged copy exists afterwards."""
    if not os.path.isfile(os.path.join(UPSTREAM, "encoder", "gcn_lib", "torch_edge.py")):
        return available()
    if available() and not force:
        return True
    os.makedirs(REF_DIR, exist_ok=True)
    for rel in _FILES:
        src, dst = os.path.join(UPSTREAM, rel), os.path.join(REF_DIR, rel)
        if os.path.isdir(src):
            shutil.copytree(src, dst, dirs_exist_ok=True, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        else:
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
    return available()


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "encoder", "gcn_lib", "torch_edge.py"))


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def _install_stubs():
    import torch.nn as nn

    class DropPath(nn.Module):  # GraphEncoder never builds it with p > 0 (dpr[0] == 0, graph_encoder.py:135,148)
        def __init__(self, p=0.0):
            super().__init__()
            if p > 0:
                raise NotImplementedError("stub DropPath only supports p == 0")

        def forward(self, x):
            return x

    if "timm" not in sys.modules:
        layers = _stub("timm.models.layers", DropPath=DropPath, to_2tuple=lambda v: (v, v),
                       trunc_normal_=lambda t, std=1.0: nn.init.trunc_normal_(t, std=std))
        _stub("timm", models=_stub("timm.models", layers=layers))
    if "torchmetrics" not in sys.modules:
        _stub("torchmetrics", functional=_stub("torchmetrics.functional", pairwise_cosine_similarity=None))
    if "librosa" not in sys.modules:
        _stub("librosa")


def load():
    """Namespace with the upstream GraphEncoder, SimCLR and ntxent_loss imported from baseline/_ref."""
    if not available():
        raise RuntimeError(f"no staged reference under {REF_DIR} (run __graft_entry__.build() where /root/reference exists)")
    _install_stubs()
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    from encoder import graph_encoder
    from simclr import simclr as simclr_mod, ntxent
    for mod in (graph_encoder, simclr_mod, ntxent):
        assert os.path.abspath(mod.__file__).startswith(os.path.abspath(REF_DIR)), mod.__file__
    return types.SimpleNamespace(GraphEncoder=graph_encoder.GraphEncoder, SimCLR=simclr_mod.SimCLR,
                                 ntxent_loss=ntxent.ntxent_loss)
