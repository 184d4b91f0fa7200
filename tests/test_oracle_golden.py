"""Pin oracle/grafp_oracle.py to the upstream reference through the committed fixtures (CPU)."""
import numpy as np
import pytest
import torch

import golden_io as gio
from grafp_b200 import synth
from oracle import grafp_oracle as O

torch.set_num_threads(4)


def test_knn_matches_reference_bit_exact():
    gold = gio.load("knn")
    for name in gold["names"]:
        name = str(name)
        x = gio.t(gold[f"{name}.x"])
        y = gio.t(gold[f"{name}.y"]) if f"{name}.y" in gold else None
        rp = gio.t(gold[f"{name}.relative_pos"]) if f"{name}.relative_pos" in gold else None
        k, d = (int(v) for v in gold[f"{name}.kd"])
        edge = O.dilated_knn_graph(x, k, d, y, rp)
        assert edge.dtype == torch.int64
        assert torch.equal(edge, gio.t(gold[f"{name}.edge_index"])), name


def test_self_is_rank0_neighbour():
    gold = gio.load("knn")
    edge = gio.t(gold["plain_k3.edge_index"])
    assert torch.equal(edge[0][..., 0], edge[1][..., 0])


@pytest.mark.parametrize("tag", ["self", "xy"])
def test_aggregation_matches_reference(tag):
    gold = gio.load("aggregate")
    x = gio.t(gold[f"{tag}.x"]).requires_grad_(True)
    y = gio.t(gold[f"{tag}.y"]).requires_grad_(True) if f"{tag}.y" in gold else None
    edge = gio.t(gold[f"{tag}.edge_index"])
    src = x if y is None else y
    assert torch.equal(O.gather_neighbors(src, edge[0]), gio.t(gold[f"{tag}.gather"]))
    feat = O.max_relative_features(x, edge, y)
    assert torch.equal(feat, gio.t(gold[f"{tag}.mr_features"]))
    feat.backward(gio.t(gold[f"{tag}.mr_upstream"]))
    assert torch.equal(x.grad, gio.t(gold[f"{tag}.mr_grad_x"]))
    if y is not None:
        assert torch.equal(y.grad, gio.t(gold[f"{tag}.mr_grad_y"]))
    x.grad = None
    if y is not None:
        y.grad = None
    ef = O.edge_features(x, edge, y)
    assert torch.equal(ef, gio.t(gold[f"{tag}.edge_features"]))
    ef.backward(gio.t(gold[f"{tag}.edge_upstream"]))
    assert torch.equal(x.grad, gio.t(gold[f"{tag}.edge_grad_x"]))


def _params_for(gold, tag, prefix_strip=""):
    p = {}
    for key in gold:
        if key.startswith(f"{tag}.grad.") or key.startswith(f"{tag}.buf."):
            name = key.split(".", 2)[2]
            p[name] = gold[key].shape
    return p


@pytest.mark.parametrize("name", [str(n) for n in gio.load("knn_cos")["names"]])
def test_cosine_knn_graphs_match_reference(name):
    """DenseDilatedKnnGraph_plg / _new (torch_edge.py:286-361): the oracle's cosine graph is the reference's, bit for bit."""
    gold = gio.load("knn_cos")
    x = gio.t(gold[f"{name}.x"])
    y = gio.t(gold[f"{name}.y"]) if f"{name}.y" in gold else None
    rp = gio.t(gold[f"{name}.relative_pos"]) if f"{name}.relative_pos" in gold else None
    k, d = (int(v) for v in gold[f"{name}.kd"])
    assert torch.equal(O.cosine_dilated_knn_graph(x, k, d, y, rp), gio.t(gold[f"{name}.edge_index"]))


@pytest.mark.parametrize("conv", ["mr", "edge", "sage", "gin"])
@pytest.mark.parametrize("d", [1, 2])
def test_graph_conv_modules_match_reference(conv, d):
    gold = gio.load("gconv")
    B, C, N, k = (int(v) for v in gold["cfg"])
    tag = f"{conv}_d{d}"
    shapes = _params_for(gold, tag)
    p = synth.synth_state_dict(shapes, 40 + d)
    for name in list(p):
        if f"{tag}.grad.{name}" in gold:
            p[name].requires_grad_(True)
    x = gio.t(gold["x"]).requires_grad_(True)
    out = O.dy_graph_conv(p, "", x, True, k, d, conv)  # keys start with "gconv." -> prefix "" + ".gconv"
    assert gio.rel_err(out, gio.t(gold[f"{tag}.out"])) < 1e-6
    out.backward(gio.t(gold[f"{tag}.upstream"]))
    assert gio.rel_err(x.grad, gio.t(gold[f"{tag}.grad_x"])) < 1e-5
    for name, v in p.items():
        if v.requires_grad:
            assert gio.rel_err(v.grad, gio.t(gold[f"{tag}.grad.{name}"])) < 1e-5, name
        elif name.endswith("running_mean") or name.endswith("running_var"):
            assert gio.rel_err(v, gio.t(gold[f"{tag}.buf.{name}"])) < 1e-6, name


@pytest.mark.parametrize("conv", ["mr", "edge", "sage", "gin"])
@pytest.mark.parametrize("d", [1, 2])
def test_graph_conv_modules_r2_match_reference(conv, d):
    """DyGraphConv2d(r=2) on an H x W map: queries against the average-pooled key set (torch_vertex.py:130-132)."""
    gold = gio.load("gconv_r2")
    B, C, H, W, k = (int(v) for v in gold["cfg"])
    tag = f"{conv}_d{d}"
    shapes = _params_for(gold, tag)
    p = synth.synth_state_dict(shapes, 50 + d)
    for name in list(p):
        if f"{tag}.grad.{name}" in gold:
            p[name].requires_grad_(True)
    x = gio.t(gold["x"]).requires_grad_(True)
    out = O.dy_graph_conv(p, "", x, True, k, d, conv, r=2)
    assert out.shape == (B, 2 * C, H, W)
    assert gio.rel_err(out, gio.t(gold[f"{tag}.out"])) < 1e-6
    out.backward(gio.t(gold[f"{tag}.upstream"]))
    assert gio.rel_err(x.grad, gio.t(gold[f"{tag}.grad_x"])) < 1e-5
    for name, v in p.items():
        if v.requires_grad:
            assert gio.rel_err(v.grad, gio.t(gold[f"{tag}.grad.{name}"])) < 1e-5, name


def test_grapher_block_matches_reference():
    gold = gio.load("grapher")
    B, C, N, k, d = (int(v) for v in gold["cfg"])
    shapes = {key[5:]: gold[key].shape for key in gold if key.startswith("grad.")}
    for bn in ("fc1.1", "graph_conv.gconv.nn.1", "fc2.1"):
        ch = shapes[bn + ".weight"]
        shapes[bn + ".running_mean"] = ch
        shapes[bn + ".running_var"] = ch
    p = synth.synth_state_dict(shapes, 61)
    for name, v in p.items():
        if f"grad.{name}" in gold:
            v.requires_grad_(True)
    x = gio.t(gold["x"]).requires_grad_(True)
    out = O.grapher(p, "", x, True, k, d)
    assert gio.rel_err(out, gio.t(gold["out"])) < 1e-6
    out.backward(gio.t(gold["upstream"]))
    assert gio.rel_err(x.grad, gio.t(gold["grad_x"])) < 1e-5
    for name, v in p.items():
        if v.requires_grad:
            assert gio.rel_err(v.grad, gio.t(gold[f"grad.{name}"])) < 1e-5, name
    with torch.no_grad():
        assert gio.rel_err(O.grapher(p, "", x.detach(), False, k, d), gio.t(gold["out_eval"])) < 1e-6


def encoder_params(gold, seed, prefix=""):
    shapes = gio.shapes_from(gold)
    p = synth.synth_state_dict({k: v for k, v in shapes.items() if not k.endswith("relative_pos")}, seed)
    trainable = set(str(s) for s in gold["requires_grad"]) if "requires_grad" in gold else None
    return p, trainable


def test_graph_encoder_matches_reference():
    gold = gio.load("encoder")
    p, trainable = encoder_params(gold, 81)
    assert len(gold["keys"]) == 437  # +6 SimCLR keys = 443
    for name in trainable:
        p[name].requires_grad_(True)
    x = gio.t(gold["x"]).requires_grad_(True)
    out = O.graph_encoder(p, x, True, k=3)
    assert out.shape == (4, 1024)
    assert gio.rel_err(out, gio.t(gold["out_train"])) < 1e-5
    (out * gio.t(gold["upstream"])).sum().backward()
    assert gio.rel_err(x.grad, gio.t(gold["grad_x"])) < 1e-4
    norms = dict(zip((str(n) for n in gold["grad_names"]), gold["grad_norm"]))
    for name in trainable:
        got = float(p[name].grad.double().norm())
        assert abs(got - norms[name]) <= 1e-4 * max(norms[name], 1e-12), name
    assert gio.rel_err(p["stem.0.weight"].grad, gio.t(gold["grad.stem.0.weight"])) < 1e-4
    assert gio.rel_err(p["backbone.0.0.graph_conv.gconv.nn.0.weight"].grad,
                       gio.t(gold["grad.backbone.0.0.graph_conv.gconv.nn.0.weight"])) < 1e-4
    assert gio.rel_err(p["stem.1.running_mean"], gio.t(gold["bn_running_mean.stem.1"])) < 1e-6
    with torch.no_grad():
        assert gio.rel_err(O.graph_encoder(p, x.detach(), False, k=3), gio.t(gold["out_eval"])) < 1e-5


def test_simclr_step_and_retrieval_match_reference():
    gold = gio.load("simclr")
    shapes = gio.shapes_from(gold)
    assert len(shapes) == 443
    p = synth.synth_state_dict({k: v for k, v in shapes.items() if not k.endswith("relative_pos")}, 101)
    s_i, s_j = gio.t(gold["spec_i"]), gio.t(gold["spec_j"])
    assert torch.equal(O.peak_extractor(p, s_i), gio.t(gold["peaks_i"]))
    h_i, h_j, z_i, z_j = O.simclr_forward(p, s_i, s_j, True)
    assert gio.rel_err(z_i, gio.t(gold["z_i"])) < 1e-5
    assert gio.rel_err(z_j, gio.t(gold["z_j"])) < 1e-5
    loss = O.ntxent_loss(z_i, z_j, synth.DEFAULT_CFG["tau"])
    assert abs(float(loss) - float(gold["loss"])) < 1e-5 * abs(float(gold["loss"]))
    with torch.no_grad():  # BatchNorm on batch statistics, as generate.py leaves the model (never .eval())
        db_specs, q_specs = synth.synth_spec(32, 121)
        _, _, db, _ = O.simclr_forward(p, db_specs, db_specs, True)
        _, _, q, _ = O.simclr_forward(p, q_specs[:16], q_specs[:16], True)
    assert gio.rel_err(db, gio.t(gold["db"])) < 1e-5
    assert torch.equal(O.top1_retrieval(db, q), gio.t(gold["top1"]))
    assert np.array_equal(O.top1_retrieval(gio.t(gold["db"]), gio.t(gold["queries"])).numpy(), gold["top1"])
