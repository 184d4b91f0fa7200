"""Import the upstream GraFP modules from /root/reference (build container only).

The upstream code needs three third-party modules that are absent here and unused
at run time on this path (timm, torchmetrics, librosa); they are stubbed.  The
upstream package names (``encoder``, ``simclr``, ``peak_extractor``) are imported
into ``sys.modules`` under their own names, so call :func:`load` only from a
process that does not use ``grafp_b200.install_dropin()`` at the same time.

/root/reference does not exist on the GPU box: nothing marked ``gpu`` and neither
``smoke()`` nor ``bench.py`` may call this.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GRAFP_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "encoder", "gcn_lib", "torch_edge.py"))


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def install_stubs():
    import torch.nn as nn

    class DropPath(nn.Module):  # never built with p > 0 by GraphEncoder (dpr[0] == 0)
        def __init__(self, p=0.0):
            super().__init__()
            if p > 0:
                raise NotImplementedError("stub DropPath only supports p == 0")

        def forward(self, x):
            return x

    def to_2tuple(v):
        return (v, v)

    def trunc_normal_(t, std=1.0):
        return nn.init.trunc_normal_(t, std=std)

    if "timm" not in sys.modules:
        layers = _stub("timm.models.layers", DropPath=DropPath, to_2tuple=to_2tuple, trunc_normal_=trunc_normal_)
        models = _stub("timm.models", layers=layers)
        _stub("timm", models=models)
    if "torchmetrics" not in sys.modules:
        fn = _stub("torchmetrics.functional", pairwise_cosine_similarity=None)
        _stub("torchmetrics", functional=fn)
    if "librosa" not in sys.modules:
        _stub("librosa")


def load():
    """Return a namespace with the upstream classes/functions used as the parity anchor."""
    if not available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from encoder.gcn_lib import torch_edge, torch_nn, torch_vertex
    from encoder import graph_encoder
    from simclr import simclr as simclr_mod, ntxent
    import peak_extractor

    for mod in (torch_edge, torch_nn, torch_vertex, graph_encoder):
        assert os.path.abspath(mod.__file__).startswith(os.path.abspath(REFERENCE_ROOT)), mod.__file__
    return types.SimpleNamespace(
        torch_edge=torch_edge, torch_nn=torch_nn, torch_vertex=torch_vertex,
        graph_encoder=graph_encoder, simclr=simclr_mod, ntxent=ntxent, peak_extractor=peak_extractor,
    )
