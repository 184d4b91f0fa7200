"""Host-side mirror of the reference interface: module tree, state_dict, drop-in aliases (CPU)."""
import inspect
import sys

import numpy as np
import pytest
import torch

import golden_io as gio
import grafp_b200
from grafp_b200 import ops, synth
from grafp_b200.encoder.gcn_lib import torch_edge, torch_nn, torch_vertex, pos_embed
from grafp_b200.encoder.graph_encoder import GraphEncoder, FFN, Downsample
from grafp_b200.simclr.simclr import SimCLR
from grafp_b200.simclr.ntxent import ntxent_loss
from oracle import grafp_oracle as O


@pytest.fixture(scope="module")
def model():
    cfg = dict(synth.DEFAULT_CFG)
    return SimCLR(cfg, GraphEncoder(cfg=cfg, in_channels=cfg["n_filters"], k=3))


def test_state_dict_keys_and_shapes_match_reference(model):
    gold = gio.load("simclr")
    sd = model.state_dict()
    assert [str(k) for k in gold["keys"]] == list(sd.keys())
    assert gio.shapes_from(gold) == {k: tuple(v.shape) for k, v in sd.items()}
    trainable = sum(p.numel() for p in model.parameters() if p.requires_grad)
    assert trainable == 18367264 and sum(p.numel() for p in model.parameters()) == 20620576


def test_relative_pos_tables_match_reference(model):
    gold = gio.load("encoder")
    sd = model.encoder.state_dict()
    names = [k[len("relpos_sum."):] for k in gold if k.startswith("relpos_sum.")]
    assert len(names) == 12
    for name in names:
        assert not dict(model.encoder.named_parameters())[name].requires_grad
        assert abs(float(sd[name].double().sum()) - float(gold["relpos_sum." + name])) < 1e-6 * abs(float(gold["relpos_sum." + name])) + 1e-9
        assert np.allclose(sd[name].flatten()[:64].numpy(), gold["relpos_head." + name], atol=1e-6)


def test_graph_modules_add_no_parameters():
    g = torch_edge.DenseDilatedKnnGraph(k=3, dilation=2)
    assert list(g.state_dict()) == [] and isinstance(g._dilated, torch_edge.DenseDilated)
    mr = torch_vertex.MRConv2d(16, 32, "relu", "batch", True)
    assert list(mr.state_dict()) == ["nn.0.weight", "nn.0.bias", "nn.1.weight", "nn.1.bias", "nn.1.running_mean",
                                     "nn.1.running_var", "nn.1.num_batches_tracked"]
    assert mr.nn[0].groups == 4 and mr.nn[0].weight.shape == (32, 8, 1, 1)


def test_signatures_match_reference():
    def params(fn):
        return [(p.name, p.default) for p in inspect.signature(fn).parameters.values() if p.name != "self"]

    assert params(torch_edge.DenseDilatedKnnGraph.__init__) == [("k", 9), ("dilation", 1), ("stochastic", False), ("epsilon", 0.0)]
    assert params(torch_edge.DenseDilatedKnnGraph.forward) == [("x", inspect._empty), ("y", None), ("relative_pos", None)]
    assert params(torch_vertex.MRConv2d.__init__) == [("in_channels", inspect._empty), ("out_channels", inspect._empty),
                                                      ("act", "relu"), ("norm", None), ("bias", True)]
    assert params(torch_vertex.MRConv2d.forward) == [("x", inspect._empty), ("edge_index", inspect._empty), ("y", None)]
    assert [n for n, _ in params(torch_vertex.DyGraphConv2d.__init__)] == [
        "in_channels", "out_channels", "kernel_size", "dilation", "conv", "act", "norm", "bias", "stochastic", "epsilon", "r"]
    assert [n for n, _ in params(torch_vertex.Grapher.__init__)] == [
        "in_channels", "kernel_size", "dilation", "conv", "act", "norm", "bias", "stochastic", "epsilon", "r", "n",
        "drop_path", "relative_pos"]
    assert params(GraphEncoder.__init__) == [
        ("cfg", inspect._empty), ("k", 3), ("conv", "mr"), ("act", "relu"), ("norm", "batch"), ("bias", True),
        ("dropout", 0.0), ("dilation", True), ("epsilon", 0.2), ("drop_path", 0.1), ("size", "t"), ("emb_dims", 1024),
        ("in_channels", 3)]
    assert params(torch_nn.batched_index_select) == [("x", inspect._empty), ("idx", inspect._empty)]
    for conv in ("edge", "mr", "sage", "gin"):
        torch_vertex.GraphConv2d(16, 32, conv, "relu", "batch", True)
    with pytest.raises(NotImplementedError):
        torch_vertex.GraphConv2d(16, 32, "gcn")


def test_every_grapher_block_uses_k3_dilation1(model):
    graphers = [m for m in model.encoder.modules() if isinstance(m, torch_vertex.Grapher)]
    assert len(graphers) == 12
    assert {(g.graph_conv.k, g.graph_conv.d, type(g.drop_path).__name__) for g in graphers} == {(3, 1, "Identity")}
    assert [g.relative_pos.shape[-1] for g in graphers] == [1024] * 2 + [256] * 2 + [64] * 6 + [16] * 2


def test_ops_refuse_cpu_tensors():
    x = torch.randn(1, 8, 16, 1)
    idx = torch.zeros(1, 16, 2, dtype=torch.int64)
    for call in (lambda: ops.knn_graph(x, 2), lambda: ops.mr_aggregate(x, idx), lambda: ops.gather_neighbors(x, idx),
                 lambda: ops.edge_features(x, idx), lambda: ops.max_over_k(torch.randn(1, 8, 16, 2)),
                 lambda: torch_edge.DenseDilatedKnnGraph(2)(x)):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()


def test_as_rows_layouts():
    x = torch.randn(2, 8, 16, 1)
    r = ops.as_rows(x)
    assert torch.equal(r, x) and r.permute(0, 2, 1, 3).is_contiguous() and r.is_contiguous(memory_format=torch.channels_last)
    assert ops._new_rows(2, 8, 16, x).is_contiguous(memory_format=torch.channels_last)
    assert ops._new_edge_rows(2, 8, 16, 3, x).permute(0, 2, 3, 1).is_contiguous()
    cl = x.contiguous(memory_format=torch.channels_last)
    assert ops.as_rows(cl).data_ptr() == cl.data_ptr()
    h = torch.randn(2, 8, 16, 3)
    hr = ops.as_edge_rows(h)
    assert torch.equal(hr, h) and hr.permute(0, 2, 3, 1).is_contiguous()


def test_install_dropin_aliases_reference_import_paths():
    saved = {k: v for k, v in sys.modules.items() if k == "encoder" or k.startswith("encoder.")}
    try:
        grafp_b200.install_dropin()
        from encoder.graph_encoder import GraphEncoder as Aliased
        from encoder.gcn_lib.torch_vertex import Grapher as AliasedGrapher
        assert Aliased is GraphEncoder and AliasedGrapher is torch_vertex.Grapher
    finally:
        for k in [k for k in sys.modules if k == "encoder" or k.startswith("encoder.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_ntxent_matches_reference_value():
    gold = gio.load("simclr")
    loss = ntxent_loss(gio.t(gold["z_i"]), gio.t(gold["z_j"]), synth.DEFAULT_CFG)
    assert abs(float(loss) - float(gold["loss"])) < 1e-5 * abs(float(gold["loss"]))
    z_i = torch.nn.functional.normalize(torch.randn(6, 16), dim=1).requires_grad_(True)
    z_j = torch.nn.functional.normalize(torch.randn(6, 16), dim=1)
    a = ntxent_loss(z_i, z_j, {"tau": 0.05})
    b = O.ntxent_loss(z_i, z_j, 0.05)
    assert torch.allclose(a, b, rtol=1e-5)


def test_pos_embed_shapes():
    t = pos_embed.get_2d_relative_pos_embed(16, 4)
    assert t.shape == (16, 16) and np.allclose(t, t.T)


def test_synth_is_deterministic():
    a, b = synth.synth_spec(3, 5)
    c, d = synth.synth_spec(3, 5)
    assert torch.equal(a, c) and torch.equal(b, d) and a.shape == (3, 64, 32) and not torch.equal(a, b)
    s1 = synth.synth_state_dict({"w.weight": (4, 3, 1, 1), "bn.running_var": (4,)}, 9)
    s2 = synth.synth_state_dict({"bn.running_var": (4,), "w.weight": (4, 3, 1, 1)}, 9)
    assert all(torch.equal(s1[k], s2[k]) for k in s1)


@pytest.mark.parametrize("groups,bias,relu,res", [(1, True, True, False), (4, True, True, False), (1, False, False, True),
                                                  (1, False, True, False), (1, True, False, False)])
def test_folded_conv_batchnorm_equals_the_eval_modules(groups, bias, relu, res, monkeypatch):
    """Inference path: conv -> BatchNorm(eval) [-> ReLU | + residual] with the BatchNorm folded into the convolution's
    weight and bias is the same map as the three modules (fp64, CPU: the algebra, not the kernels)."""
    monkeypatch.setenv("GRAFP_FOLD_BN", "2")     # plain conv2d form (cudnn_convolution_relu is CUDA-only)
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(16, 32, 1, groups=groups, bias=bias).double()
    bn = torch.nn.BatchNorm2d(32).double()
    with torch.no_grad():
        bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2.0); bn.weight.normal_(); bn.bias.normal_()
    bn.eval()
    x = torch.randn(3, 16, 50, 1, dtype=torch.float64)
    r = torch.randn(3, 32, 50, 1, dtype=torch.float64) if res else None
    with torch.no_grad():
        want = bn(conv(x))
        want = want + r if res else want
        want = torch.relu(want) if relu else want
        got = ops._folded_conv_bn_eval(x, conv.weight, conv.bias, bn, relu, r.clone() if res else None,
                                       (conv.stride, conv.padding, conv.dilation, conv.groups))
    assert float((got - want).abs().max()) < 1e-12


def test_fold_and_graph_paths_are_gated_to_cuda_inference():
    """The folded path is taken only in eval mode under no_grad on CUDA tensors; GraphedEncoder refuses CPU inputs."""
    from grafp_b200.inference import GraphedEncoder
    bn = torch.nn.BatchNorm2d(8)
    x = torch.randn(2, 8, 4, 1)
    with torch.no_grad():
        assert not ops._fold_eval_ok(x, bn.eval())          # CPU tensor
    assert not ops._fold_eval_ok(x, bn.train())
    with pytest.raises(RuntimeError, match="CUDA"):
        GraphedEncoder(lambda t: t, x)


def test_fingerprint_writer_matches_the_reference_memmap_format(tmp_path):
    """db.mm / db_shape.npy exactly as test_fp.py:108-125 writes them: raw float32 (n, d) + the shape tuple."""
    from grafp_b200.inference import FingerprintWriter, load_fingerprint_db
    g = torch.Generator().manual_seed(5)
    chunks = [torch.randn(n, 128, generator=g) for n in (128, 128, 37)]
    want = torch.cat(chunks).numpy()
    with FingerprintWriter(str(tmp_path), "db", want.shape[0], 128) as w:
        for c in chunks:
            w.append(c)
    # the reference's own way of writing / reading the same data
    ref = np.memmap(tmp_path / "ref.mm", dtype="float32", mode="w+", shape=want.shape)
    ref[:] = want[:]
    ref.flush(); del ref
    assert (tmp_path / "db.mm").read_bytes() == (tmp_path / "ref.mm").read_bytes()
    assert tuple(np.load(tmp_path / "db_shape.npy")) == want.shape
    data, shape = load_fingerprint_db(str(tmp_path), "db")
    assert shape == want.shape and np.array_equal(np.asarray(data), want)
    with pytest.raises(ValueError):
        with FingerprintWriter(str(tmp_path), "short", 10, 128) as w:
            w.append(torch.zeros(4, 128))


def test_block_diagonal_form_of_a_grouped_pointwise_convolution():
    """bf16 training runs BasicConv's grouped (4) 1x1 convolution as a dense one with a block-diagonal weight
    (ops._dense_form_of_grouped): same outputs, input gradient and - on the diagonal blocks - weight gradient (fp64)."""
    g = torch.Generator().manual_seed(0)
    cw = torch.randn(32, 4, 1, 1, generator=g, dtype=torch.float64, requires_grad=True)
    x = torch.randn(2, 16, 10, 1, generator=g, dtype=torch.float64, requires_grad=True)
    ref = torch.nn.functional.conv2d(x, cw, None, groups=4)
    up = torch.randn(ref.shape, generator=g, dtype=torch.float64)
    ref.backward(up)
    w = ops._block_diag_weight(cw.detach(), 4).requires_grad_(True)
    x2 = x.detach().clone().requires_grad_(True)
    out = torch.nn.functional.conv2d(x2, w)
    out.backward(up)
    assert torch.equal(out, ref) and torch.equal(x2.grad, x.grad)
    assert torch.equal(ops._block_diag_grad(w.grad, 4), cw.grad)
    assert ops._dense_form_of_grouped(x.bfloat16(), cw, 4) and not ops._dense_form_of_grouped(x, cw, 4)
