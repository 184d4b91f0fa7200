"""grafp_b200 - B200-native GraphEncoder hot path of GraFP (chymaera96/GraFP).

The dynamic dilated k-NN graph and the max-relative / edge graph-convolution aggregation run as
hand-written sm_100a CUDA kernels behind a C ABI (``include/grafp_b200.h``,
``grafp_b200/lib/libgrafp_b200.so``); ``grafp_b200.encoder`` mirrors the reference's
``encoder.gcn_lib`` / ``encoder.graph_encoder`` modules on top of them.  CUDA only - there is no
CPU or PyTorch-eager fallback for the graph ops.
"""
import importlib
import sys

__version__ = "0.1.0"

_DROPIN_MODULES = {
    # reference import path -> module of this package
    "encoder": "grafp_b200.encoder",
    "encoder.gcn_lib": "grafp_b200.encoder.gcn_lib",
    "encoder.gcn_lib.torch_nn": "grafp_b200.encoder.gcn_lib.torch_nn",
    "encoder.gcn_lib.torch_edge": "grafp_b200.encoder.gcn_lib.torch_edge",
    "encoder.gcn_lib.torch_vertex": "grafp_b200.encoder.gcn_lib.torch_vertex",
    "encoder.gcn_lib.pos_embed": "grafp_b200.encoder.gcn_lib.pos_embed",
    "encoder.graph_encoder": "grafp_b200.encoder.graph_encoder",
}


def install_dropin(include_callers: bool = False) -> None:
    """Make the reference's import paths resolve to this package.

    After this call ``from encoder.graph_encoder import GraphEncoder`` and
    ``from encoder.gcn_lib.torch_vertex import Grapher`` (what the reference's train.py,
    generate.py and test_fp.py do) import the B200 implementations, so those scripts run
    unchanged.  With ``include_callers`` the ``simclr`` and ``peak_extractor`` modules are
    aliased too.  Call it before the reference modules are imported.
    """
    table = dict(_DROPIN_MODULES)
    if include_callers:
        table.update({
            "simclr": "grafp_b200.simclr",
            "simclr.simclr": "grafp_b200.simclr.simclr",
            "simclr.ntxent": "grafp_b200.simclr.ntxent",
            "peak_extractor": "grafp_b200.peak_extractor",
        })
    for alias, target in table.items():
        sys.modules[alias] = importlib.import_module(target)
