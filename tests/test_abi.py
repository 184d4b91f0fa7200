"""The C-ABI library builds, loads and exports exactly what include/grafp_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

from grafp_b200 import _native, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "grafp_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(grafp_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib_path():
    return build.build()


def test_header_and_binding_agree():
    assert declared_functions() == sorted(_native.SIGNATURES)


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in declared_functions():
        assert hasattr(lib, name), name


def test_library_links_without_the_cuda_driver(lib_path):
    import subprocess
    out = subprocess.run(["ldd", lib_path], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out and "not found" not in out


def test_abi_version_and_sizes(lib_path):
    lib = _native.load()
    assert lib.grafp_abi_version() == _native.ABI_VERSION == 7
    assert lib.grafp_knn_workspace_bytes(0, 1, 1, 1, 1, 0) == 0
    need = lib.grafp_knn_workspace_bytes(4, 256, 256, 64, 3, 0)
    assert need >= 4 * 4 * 256 * 64 * 4  # hi + lo for queries and keys


def test_calls_fail_loudly_without_a_device(lib_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = _native.load()
    rc = lib.grafp_max_over_k_fwd(None, None, None, 1, 1, 4, 1, 0, None)
    assert rc == -4  # GRAFP_ENODEVICE: no CPU path
    assert b"no CPU path" in lib.grafp_last_error()
    rc = lib.grafp_knn_fwd(None, None, None, None, None, 1, 8, 8, 4, 9, 1, 0, 1, 0, 0, 0, None, 0, None)
    assert rc != 0


def test_missing_library_is_an_error(monkeypatch, tmp_path):
    monkeypatch.setenv("GRAFP_B200_LIB", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_native, "_lib", None)
    with pytest.raises(RuntimeError, match="no PyTorch/CPU fallback"):
        _native.load()


def test_options_are_settable_and_default_from_the_header(lib_path):
    """grafp_set_option / grafp_get_option work without a device (they select kernels, they launch nothing)."""
    lib = _native.load()
    assert lib.grafp_get_option(b"mr_bwd_form") == 2 and lib.grafp_get_option(b"mr_fwd_form") == 2
    assert lib.grafp_get_option(b"knn_epilogue") == 0 and lib.grafp_get_option(b"check_index") == 0
    assert lib.grafp_set_option(b"mr_bwd_form", 3) == 0 and lib.grafp_get_option(b"mr_bwd_form") == 3
    assert lib.grafp_set_option(b"mr_bwd_form", 2) == 0
    assert lib.grafp_set_option(b"no_such_option", 1) == -1
    assert b"unknown option" in lib.grafp_last_error()
    assert lib.grafp_get_option(b"no_such_option") == -1


def test_conv1x1_envelope_is_answered_on_the_host(lib_path):
    """grafp_conv1x1_bn_stats_supported is pure host logic: the shapes of the encoder (dense and BasicConv's 4 groups) are
    in, shapes the 64 / 128 / 256-column tiling cannot express are refused before any launch."""
    lib = _native.load()
    ok = lib.grafp_conv1x1_bn_stats_supported
    F32, BF16 = 0, 1
    for C in (64, 128, 256, 512):
        R = 512 * 65536 // C
        for cin, cout, g in ((C, C, 1), (2 * C, 2 * C, 4), (2 * C, C, 1), (C, 4 * C, 1), (4 * C, C, 1), (3 * C // 2 * 2, 2 * C, 1)):
            assert ok(R, cin, cout, g, F32) == 1 and ok(R, cin, cout, g, BF16) == 1, (cin, cout, g)
    assert ok(1000, 8, 64, 1, F32) == 1 and ok(1000, 8, 64, 1, BF16) == 1      # the stem's 8 input channels: 32 / 16 bytes
    assert ok(1000, 6, 64, 1, F32) == 0                                        # 24-byte rows
    assert ok(1000, 4, 64, 1, BF16) == 0                                       # 8-byte rows
    assert ok(1000, 96, 96, 4, F32) == 0                                       # 24 output channels per group
    assert ok(1000, 128, 128, 3, F32) == 0                                     # channels not divisible by the groups
    assert ok(1000, 128, 192, 2, F32) == 0                                     # 96 per group: neither fills nor divides a tile
    assert ok(1000, 128, 128, 8, F32) == 1 and ok(1000, 64, 64, 4, F32) == 1   # 16 per group: four / four groups per tile
    assert ok(0, 64, 64, 1, F32) == 0 and ok(1 << 31, 64, 64, 1, F32) == 0
    assert ok(1000, 64, 64, 1, 7) == 0
