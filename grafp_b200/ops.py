"""PyTorch-facing operators over the C ABI: the dilated k-NN graph and the graph-conv
aggregations, forward and backward (torch.autograd.Function), on CUDA tensors only.

Logical tensor shapes follow the reference (node features (B, C, N, 1), edge features
(B, C, N, k)); physically every tensor handed to the library is channels-last, i.e. rows
(B, N, C) / (B, N, k, C).  Inputs in any other layout are converted once.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch

from . import _native

_DTYPES = {torch.float32: 0, torch.bfloat16: 1}


class KernelTimer:
    """Optional per-call instrumentation (bench.py): CUDA events around every C-ABI call on the
    launching stream, plus a count of the kernels this package launched."""

    def __init__(self, timing: bool = True):
        self.timing = timing
        self.launches = 0
        self.records = []  # (op name, meta dict, start event, end event)

    def summary(self):
        """name -> {calls, ms_total, meta of the costliest shape}; call after torch.cuda.synchronize()."""
        out = {}
        for name, meta, e0, e1 in self.records:
            ms = e0.elapsed_time(e1)
            slot = out.setdefault(name, {"calls": 0, "ms_total": 0.0, "by_shape": {}})
            slot["calls"] += 1
            slot["ms_total"] += ms
            key = tuple(sorted(meta.items()))
            sh = slot["by_shape"].setdefault(key, {"calls": 0, "ms_total": 0.0})
            sh["calls"] += 1
            sh["ms_total"] += ms
        return out


_TIMER: Optional[KernelTimer] = None


def set_timer(timer: Optional[KernelTimer]) -> None:
    global _TIMER
    _TIMER = timer


def _call(name: str, n_kernels: int, meta: dict, fn, dev: torch.device, *args) -> None:
    """Invoke one C-ABI entry point with ``dev`` (the tensors' device) current, raising on a non-zero return code.

    The library launches on the *current* CUDA device and configures its kernels per device, so a call for tensors
    on another GPU of the process (nn.DataParallel replicas, a model on cuda:1) switches to it for the call."""
    if dev.index is not None and dev.index != torch.cuda.current_device():
        with torch.cuda.device(dev):
            return _call(name, n_kernels, meta, fn, dev, *args)
    t = _TIMER
    if t is None:
        _native.check(fn(*args), name)
        return
    t.launches += n_kernels
    if t.timing:
        stream = torch.cuda.current_stream(dev)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        _native.check(fn(*args), name)
        e1.record(stream)
        t.records.append((name, meta, e0, e1))
    else:
        _native.check(fn(*args), name)


def set_option(name: str, value: int) -> None:
    """Process-wide kernel-selection / diagnostic option of the library (see include/grafp_b200.h)."""
    _native.check(_native.load().grafp_set_option(name.encode(), int(value)), "set_option")


def get_option(name: str) -> int:
    return int(_native.load().grafp_get_option(name.encode()))


_WARNED = set()


def _warn_once(key: str, message: str) -> None:
    """Performance cliffs are never silent: the first call that leaves a fast path says so (once per reason)."""
    if key not in _WARNED:
        _WARNED.add(key)
        import warnings
        warnings.warn("grafp_b200: " + message, RuntimeWarning, stacklevel=3)


def _dtype_code(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise RuntimeError(f"grafp_b200: dtype {t.dtype} is not supported (float32 and bfloat16 are)") from None


def _require_cuda(*tensors: Optional[torch.Tensor]) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("grafp_b200: expected CUDA tensors - the B200 path has no CPU fallback")


def _stream(t: torch.Tensor) -> int:
    """Handle of the current torch stream of the tensor's own device."""
    return torch.cuda.current_stream(t.device).cuda_stream


def _same_device(first: torch.Tensor, *rest: Optional[torch.Tensor]) -> None:
    for t in rest:
        if t is not None and t.device != first.device:
            raise RuntimeError(f"grafp_b200: tensors on different devices ({first.device} and {t.device})")


def _check_index(idx: torch.Tensor, limit: int, what: str) -> None:
    """With option check_index = 1: raise like the reference's advanced indexing (IndexError, torch_nn.py:92-96) when a
    user-supplied graph holds an id outside [0, limit).  One small kernel + a host read, so it is off by default;
    graphs produced by the k-NN op are in range by construction and are never checked."""
    if get_option("check_index") == 0:
        return
    lib = _native.load()
    bad = torch.empty(1, dtype=torch.int32, device=idx.device)
    idx_c = idx.contiguous()
    _call("check_index", 1, {}, lib.grafp_check_index, idx.device, idx_c.data_ptr(), int(idx_c.dtype == torch.int64),
          idx_c.numel(), int(limit), bad.data_ptr(), _stream(idx_c))
    n_bad = int(bad.item())
    if n_bad:
        raise IndexError(f"grafp_b200.{what}: {n_bad} index entries are outside [0, {limit})")


def as_rows(x: torch.Tensor) -> torch.Tensor:
    """Return ``x`` (B, C, N, 1) backed by row-major (B, N, C) memory, copying only if needed."""
    if x.dim() != 4 or x.shape[3] != 1:
        raise RuntimeError(f"grafp_b200: expected a (B, C, N, 1) tensor, got {tuple(x.shape)}")
    B, C, N, _ = x.shape
    s = x.stride()
    if (C == 1 or s[1] == 1) and (N == 1 or s[2] == C) and (B == 1 or s[0] == N * C):
        return x
    # canonical channels-last strides (N*C, 1, C, C): PyTorch / cuDNN recognise the result as NHWC
    out = torch.empty_like(x, memory_format=torch.channels_last)
    out.copy_(x)
    return out


def canonical_rows(x: torch.Tensor) -> torch.Tensor:
    """``x`` (B, C, N, 1) as node rows with the CANONICAL channels-last strides (N*C, 1, C, C).  A view whose size-1 last
    dimension carries another stride (e.g. ``t.unsqueeze(-1)``) is the same memory, but PyTorch / cuDNN only recognise
    the canonical form as NHWC; anything else makes the next convolution fall back to an NCHW copy."""
    x = as_rows(x)
    B, C, N, _ = x.shape
    want = (N * C, 1, C, C)
    return x if x.stride() == want else x.as_strided(x.shape, want)


def as_edge_rows(h: torch.Tensor) -> torch.Tensor:
    """Return ``h`` (B, C, N, k) backed by (B, N, k, C) memory."""
    B, C, N, k = h.shape
    s = h.stride()
    if (C == 1 or s[1] == 1) and (k == 1 or s[3] == C) and (N == 1 or s[2] == k * C) and (B == 1 or s[0] == N * k * C):
        return h
    out = torch.empty_like(h, memory_format=torch.channels_last)
    out.copy_(h)
    return out


def _new_rows(B: int, C: int, N: int, like: torch.Tensor) -> torch.Tensor:
    # strides (N*C, 1, C, C): rows (B, N, C) in memory *and* canonical torch.channels_last, so the
    # cuDNN convolutions / batch norms that follow consume it without a layout copy
    return torch.empty((B, C, N, 1), dtype=like.dtype, device=like.device, memory_format=torch.channels_last)


def _new_edge_rows(B: int, C: int, N: int, k: int, like: torch.Tensor) -> torch.Tensor:
    return torch.empty((B, C, N, k), dtype=like.dtype, device=like.device, memory_format=torch.channels_last)


def _index_arg(idx: torch.Tensor) -> Tuple[torch.Tensor, int]:
    if idx.dtype not in (torch.int64, torch.int32):
        raise RuntimeError(f"grafp_b200: index tensors must be int64 or int32, got {idx.dtype}")
    return idx.contiguous(), int(idx.dtype == torch.int64)


# --------------------------------------------------------------------------------------
# k-NN graph
# --------------------------------------------------------------------------------------

def knn_graph(x: torch.Tensor, k: int, dilation: int = 1, y: Optional[torch.Tensor] = None,
              relative_pos: Optional[torch.Tensor] = None, emit_all: bool = False, normalize: bool = True,
              want_i32: bool = True, algo: int = _native.KNN_AUTO,
              out: Optional[torch.Tensor] = None, metric: str = "l2") -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Neighbour ids of the dilated k-NN graph (reference: torch_edge.py:270-284).

    x: (B, C, N, 1) queries, y: (B, C, M, 1) keys or None, relative_pos: (1, N, M) or None.
    Returns (nn_idx int64 (B, N, k_out), nn_idx32 int32 or None), k_out = k (or k*dilation with
    ``emit_all``).  ``out`` may be a preallocated contiguous int64 (B, N, k_out) destination
    (e.g. ``edge_index[0]``).  Not differentiable, like the reference (no_grad + detach).
    """
    _require_cuda(x, y, relative_pos)
    _same_device(x, y, relative_pos, out)
    lib = _native.load()
    with torch.no_grad():
        xr = as_rows(x.detach())
        B, C, N, _ = xr.shape
        yr = None
        M = N
        if y is not None:
            yr = as_rows(y.detach().to(xr.dtype))
            if yr.shape[0] != B or yr.shape[1] != C:
                raise RuntimeError("grafp_b200.knn_graph: x and y must agree in batch and channel size")
            M = yr.shape[2]
        rp = None
        if relative_pos is not None:
            rp = relative_pos.detach().to(torch.float32).reshape(-1, relative_pos.shape[-1]).contiguous()
            if metric == "cosine":
                rp = rp * 2  # the kernels rank 2 (1 - x.y) + relpos' (see include/grafp_b200.h)
            if rp.shape != (N, M):
                raise RuntimeError(f"grafp_b200.knn_graph: relative_pos must be (1, {N}, {M}), got {tuple(relative_pos.shape)}")
        K = int(k) * int(dilation)
        k_out = K if emit_all else int(k)
        if out is None:
            out = torch.empty((B, N, k_out), dtype=torch.int64, device=x.device)
        elif out.shape != (B, N, k_out) or out.dtype != torch.int64 or not out.is_contiguous():
            raise RuntimeError("grafp_b200.knn_graph: `out` must be a contiguous int64 (B, N, k_out) tensor")
        out32 = torch.empty((B, N, k_out), dtype=torch.int32, device=x.device) if want_i32 else None
        dt = _dtype_code(xr)
        ws_bytes = lib.grafp_knn_workspace_bytes(B, N, M, C, K, dt)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        _call("knn_fwd", 2 if yr is None else 3, dict(B=B, N=N, M=M, C=C, K=K, dtype=dt), lib.grafp_knn_fwd, x.device,
              xr.data_ptr(), yr.data_ptr() if yr is not None else None,
              rp.data_ptr() if rp is not None else None, out.data_ptr(),
              out32.data_ptr() if out32 is not None else None,
              B, N, M, C, int(k), int(dilation), int(emit_all), int(normalize), dt, int(algo),
              _native.METRIC_COSINE if metric == "cosine" else _native.METRIC_L2, ws.data_ptr(), ws_bytes, _stream(xr))
        if algo == _native.KNN_AUTO and N >= 128 and M >= 128 and C >= 32:
            variant = lib.grafp_knn_last_variant().decode()
            if variant != "f16x3":
                _warn_once(f"knn-{variant}-{C % 8}-{int(normalize)}-{int(rp is not None)}",
                           f"k-NN on (N={N}, M={M}, C={C}) took the {variant} kernel, not the fp16-plane tcgen05 path (needs "
                           "normalised features, no relative_pos, C % 8 == 0, C >= 64, N, M >= 128): 2.7x (tf32x3) to 9x (SIMT) slower")
    return out, out32


def knn_last_algo() -> str:
    return _native.load().grafp_knn_last_algo().decode()


def knn_last_variant() -> str:
    return _native.load().grafp_knn_last_variant().decode()


# --------------------------------------------------------------------------------------
# max-relative aggregation (MRConv2d body)
# --------------------------------------------------------------------------------------

class _MRAggregate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, nbr, ctr):
        lib = _native.load()
        xr = as_rows(x)
        B, C, N, _ = xr.shape
        yr = as_rows(y.to(xr.dtype)) if y is not None else None
        M = yr.shape[2] if yr is not None else N
        nbr_c, i64 = _index_arg(nbr)
        ctr_c = None
        if ctr is not None:
            ctr_c = ctr.to(nbr_c.dtype).contiguous()
        k = nbr_c.shape[-1]
        if nbr_c.shape != (B, N, k):
            raise RuntimeError(f"grafp_b200.mr_aggregate: neighbour index must be ({B}, {N}, k), got {tuple(nbr.shape)}")
        need_grad = any(ctx.needs_input_grad[:2])
        out = _new_rows(B, 2 * C, N, xr)
        argmax = torch.empty((B, N, C), dtype=torch.uint8, device=x.device) if need_grad else None
        _call("mr_aggregate_fwd", 1, dict(B=B, N=N, M=M, C=C, k=k, dtype=_dtype_code(xr), i64=i64,
                                          argmax=int(argmax is not None)),
              lib.grafp_mr_aggregate_fwd, xr.device, xr.data_ptr(), yr.data_ptr() if yr is not None else None,
              nbr_c.data_ptr(), ctr_c.data_ptr() if ctr_c is not None else None, i64, out.data_ptr(),
              argmax.data_ptr() if argmax is not None else None, B, N, M, C, k, _dtype_code(xr), _stream(xr))
        ctx.save_for_backward(nbr_c, ctr_c, argmax)
        ctx.dims = (B, N, M, C, k, i64, y is not None, _dtype_code(xr))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _native.load()
        nbr_c, ctr_c, argmax = ctx.saved_tensors
        B, N, M, C, k, i64, has_y, dt = ctx.dims
        g = as_rows(grad_out)
        grad_x = _new_rows(B, C, N, g)
        grad_y = _new_rows(B, C, M, g) if has_y else None
        ws, ws_bytes = None, 0
        if ctr_c is None and not has_y and get_option("mr_bwd_form") == 3:
            # only the deterministic gather form reads the reverse-graph workspace
            ws_bytes = lib.grafp_mr_aggregate_bwd_workspace_bytes(B, N, k)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=g.device)
        _call("mr_aggregate_bwd", 1 if (ctr_c is None and not has_y) else 2, dict(B=B, N=N, M=M, C=C, k=k, dtype=dt, i64=i64),
              lib.grafp_mr_aggregate_bwd, g.device, g.data_ptr(), argmax.data_ptr(), nbr_c.data_ptr(),
              ctr_c.data_ptr() if ctr_c is not None else None, i64, grad_x.data_ptr(),
              grad_y.data_ptr() if grad_y is not None else None, B, N, M, C, k, dt,
              ws.data_ptr() if ws is not None else None, ws_bytes, _stream(g))
        return grad_x, grad_y, None, None


def mr_aggregate(x: torch.Tensor, nbr: torch.Tensor, y: Optional[torch.Tensor] = None,
                 ctr: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(B, 2C, N, 1) interleaved [x_c, max_j(x_j - x_i)_c] (reference: torch_vertex.py:21-32).

    nbr / ctr: (B, N, k) neighbour / centre ids; ``ctr=None`` means the centre of row n is n.
    """
    _require_cuda(x, y, nbr, ctr)
    _same_device(x, y, nbr, ctr)
    if ctr is not None or y is not None:  # a user-supplied graph (the k-NN op's graphs come tagged, without ctr)
        _check_index(nbr, (y if y is not None else x).shape[2], "mr_aggregate")
        if ctr is not None:
            _check_index(ctr, x.shape[2], "mr_aggregate")
    return _MRAggregate.apply(x, y, nbr, ctr)


# --------------------------------------------------------------------------------------
# plain gather (batched_index_select)
# --------------------------------------------------------------------------------------

class _Gather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, idx):
        lib = _native.load()
        sr = as_rows(src)
        B, C, M, _ = sr.shape
        idx_c, i64 = _index_arg(idx)
        _, N, k = idx_c.shape
        out = _new_edge_rows(B, C, N, k, sr)
        _call("gather_fwd", 1, dict(B=B, N=N, M=M, C=C, k=k), lib.grafp_gather_fwd, sr.device, sr.data_ptr(),
              idx_c.data_ptr(), i64, out.data_ptr(), B, N, M, C, k, _dtype_code(sr), _stream(sr))
        ctx.save_for_backward(idx_c)
        ctx.dims = (B, N, M, C, k, i64, _dtype_code(sr))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _native.load()
        (idx_c,) = ctx.saved_tensors
        B, N, M, C, k, i64, dt = ctx.dims
        g = as_edge_rows(grad_out)
        grad_src = _new_rows(B, C, M, g)
        _call("gather_bwd", 1, dict(B=B, N=N, M=M, C=C, k=k), lib.grafp_gather_bwd, g.device, g.data_ptr(),
              idx_c.data_ptr(), i64, grad_src.data_ptr(), B, N, M, C, k, dt, _stream(g))
        return grad_src, None


def gather_neighbors(src: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """out[b, c, n, j] = src[b, c, idx[b, n, j]] as a (B, C, N, k) tensor (reference: torch_nn.py:79-98)."""
    _require_cuda(src, idx)
    if idx.dim() != 3 or idx.shape[0] != src.shape[0]:
        raise RuntimeError("grafp_b200.gather_neighbors: idx must be (B, N, k)")
    _same_device(src, idx)
    _check_index(idx, src.shape[2], "gather_neighbors")
    return _Gather.apply(src, idx)


class _NeighborSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, idx):
        lib = _native.load()
        sr = as_rows(src)
        B, C, M, _ = sr.shape
        idx_c, i64 = _index_arg(idx)
        _, N, k = idx_c.shape
        out = _new_rows(B, C, N, sr)
        _call("neighbor_sum_fwd", 1, dict(B=B, N=N, M=M, C=C, k=k), lib.grafp_neighbor_sum_fwd, sr.device, sr.data_ptr(),
              idx_c.data_ptr(), i64, out.data_ptr(), B, N, M, C, k, _dtype_code(sr), _stream(sr))
        ctx.save_for_backward(idx_c)
        ctx.dims = (B, N, M, C, k, i64, _dtype_code(sr))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _native.load()
        (idx_c,) = ctx.saved_tensors
        B, N, M, C, k, i64, dt = ctx.dims
        g = as_rows(grad_out)
        grad_src = _new_rows(B, C, M, g)
        _call("neighbor_sum_bwd", 1, dict(B=B, N=N, M=M, C=C, k=k), lib.grafp_neighbor_sum_bwd, g.device, g.data_ptr(),
              idx_c.data_ptr(), i64, grad_src.data_ptr(), B, N, M, C, k, dt, _stream(g))
        return grad_src, None


def neighbor_sum(src: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """out[b, c, n, 0] = sum_j src[b, c, idx[b, n, j]] as a (B, C, N, 1) tensor: ``batched_index_select`` followed by
    ``torch.sum(x_j, -1, keepdim=True)`` (reference: torch_vertex.py:84-88, GINConv2d) without the (B, C, N, k)
    intermediate."""
    _require_cuda(src, idx)
    if idx.dim() != 3 or idx.shape[0] != src.shape[0]:
        raise RuntimeError("grafp_b200.neighbor_sum: idx must be (B, N, k)")
    _same_device(src, idx)
    return _NeighborSum.apply(src, idx)


# --------------------------------------------------------------------------------------
# EdgeConv features and max over k
# --------------------------------------------------------------------------------------

class _EdgeGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, nbr, ctr):
        lib = _native.load()
        xr = as_rows(x)
        B, C, N, _ = xr.shape
        yr = as_rows(y.to(xr.dtype)) if y is not None else None
        M = yr.shape[2] if yr is not None else N
        nbr_c, i64 = _index_arg(nbr)
        ctr_c = ctr.to(nbr_c.dtype).contiguous() if ctr is not None else None
        k = nbr_c.shape[-1]
        out = _new_edge_rows(B, 2 * C, N, k, xr)
        _call("edge_gather_fwd", 1, dict(B=B, N=N, M=M, C=C, k=k), lib.grafp_edge_gather_fwd, xr.device, xr.data_ptr(),
              yr.data_ptr() if yr is not None else None, nbr_c.data_ptr(),
              ctr_c.data_ptr() if ctr_c is not None else None, i64, out.data_ptr(),
              B, N, M, C, k, _dtype_code(xr), _stream(xr))
        ctx.save_for_backward(nbr_c, ctr_c)
        ctx.dims = (B, N, M, C, k, i64, y is not None, _dtype_code(xr))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _native.load()
        nbr_c, ctr_c = ctx.saved_tensors
        B, N, M, C, k, i64, has_y, dt = ctx.dims
        g = as_edge_rows(grad_out)
        grad_x = _new_rows(B, C, N, g)
        grad_y = _new_rows(B, C, M, g) if has_y else None
        _call("edge_gather_bwd", 2 if ctr_c is None else 1, dict(B=B, N=N, M=M, C=C, k=k), lib.grafp_edge_gather_bwd,
              g.device, g.data_ptr(), nbr_c.data_ptr(), ctr_c.data_ptr() if ctr_c is not None else None,
              i64, grad_x.data_ptr(), grad_y.data_ptr() if grad_y is not None else None,
              B, N, M, C, k, dt, _stream(g))
        return grad_x, grad_y, None, None


def edge_features(x: torch.Tensor, nbr: torch.Tensor, y: Optional[torch.Tensor] = None,
                  ctr: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(B, 2C, N, k) = cat([x_i, x_j - x_i], dim=1) (reference: torch_vertex.py:46-51)."""
    _require_cuda(x, y, nbr, ctr)
    _same_device(x, y, nbr, ctr)
    if ctr is not None or y is not None:
        _check_index(nbr, (y if y is not None else x).shape[2], "edge_features")
        if ctr is not None:
            _check_index(ctr, x.shape[2], "edge_features")
    return _EdgeGather.apply(x, y, nbr, ctr)


class _MaxOverK(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h):
        lib = _native.load()
        hr = as_edge_rows(h)
        B, C, N, k = hr.shape
        out = _new_rows(B, C, N, hr)
        argmax = torch.empty((B, N, C), dtype=torch.uint8, device=h.device) if ctx.needs_input_grad[0] else None
        _call("max_over_k_fwd", 1, dict(B=B, N=N, C=C, k=k), lib.grafp_max_over_k_fwd, hr.device, hr.data_ptr(),
              out.data_ptr(), argmax.data_ptr() if argmax is not None else None, B, N, C, k, _dtype_code(hr), _stream(hr))
        ctx.save_for_backward(argmax)
        ctx.dims = (B, N, C, k, _dtype_code(hr))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _native.load()
        (argmax,) = ctx.saved_tensors
        B, N, C, k, dt = ctx.dims
        g = as_rows(grad_out)
        grad_h = _new_edge_rows(B, C, N, k, g)
        _call("max_over_k_bwd", 1, dict(B=B, N=N, C=C, k=k), lib.grafp_max_over_k_bwd, g.device, g.data_ptr(),
              argmax.data_ptr(), grad_h.data_ptr(), B, N, C, k, dt, _stream(g))
        return grad_h


def max_over_k(h: torch.Tensor) -> torch.Tensor:
    """torch.max(h, -1, keepdim=True).values for h (B, C, N, k) (reference: torch_vertex.py:51,69)."""
    _require_cuda(h)
    if h.dim() != 4:
        raise RuntimeError("grafp_b200.max_over_k: expected a (B, C, N, k) tensor")
    return _MaxOverK.apply(h)


# --------------------------------------------------------------------------------------
# peak point-cloud front end
# --------------------------------------------------------------------------------------

class _PeakExtract(torch.autograd.Function):
    @staticmethod
    def forward(ctx, spec, weight, bias, stride_h):
        lib = _native.load()
        B, H, W = spec.shape
        F_, _, kh, kw = weight.shape
        Ho = (H + 2 * (kh // 2) - kh) // stride_h + 1
        # logical (B, F, N, 1) with node rows (B, N, F) in memory: what GraphEncoder.forward would otherwise copy into
        out = torch.empty((B, F_, Ho * W, 1), dtype=torch.float32, device=spec.device, memory_format=torch.channels_last)
        w, b = weight.detach().contiguous(), bias.detach().contiguous()
        _call("peak_extract_fwd", 1, dict(B=B, N=Ho * W, C=F_), lib.grafp_peak_extract_fwd, spec.device, spec.data_ptr(),
              w.data_ptr(), b.data_ptr(), out.data_ptr(), B, H, W, F_, kh, kw, int(stride_h), _stream(spec))
        ctx.save_for_backward(spec, out)
        ctx.geom = (B, H, W, F_, kh, kw, int(stride_h))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _native.load()
        spec, out = ctx.saved_tensors
        B, H, W, F_, kh, kw, sh = ctx.geom
        g = as_rows(grad_out.to(torch.float32))
        ws_bytes = lib.grafp_peak_extract_workspace_bytes(B, kh, kw)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=spec.device)
        dweight = torch.empty((F_, 3, kh, kw), dtype=torch.float32, device=spec.device)
        dbias = torch.empty(F_, dtype=torch.float32, device=spec.device)
        _call("peak_extract_bwd", 2, dict(B=B, N=out.shape[2], C=F_), lib.grafp_peak_extract_bwd, spec.device, spec.data_ptr(),
              out.data_ptr(), g.data_ptr(), ws.data_ptr(), ws_bytes, dweight.data_ptr(), dbias.data_ptr(), B, H, W, F_, kh, kw, sh,
              _stream(spec))
        return None, dweight, dbias, None


def peak_extract_supported(spec: torch.Tensor, conv: torch.nn.Conv2d) -> bool:
    kh, kw = conv.kernel_size
    return (spec.is_cuda and spec.dtype == torch.float32 and spec.dim() == 3 and conv.weight.dtype == torch.float32
            and conv.in_channels == 3 and conv.out_channels == 8 and kh % 2 == 1 and kw % 2 == 1 and conv.bias is not None
            and conv.stride[1] == 1 and tuple(conv.padding) == (kh // 2, kw // 2) and tuple(conv.dilation) == (1, 1)
            and conv.groups == 1 and conv.padding_mode == "zeros" and not spec.requires_grad
            and 3 * (spec.shape[1] + kh) * (spec.shape[2] + kw) * 4 + spec.shape[1] * spec.shape[2] * 16 < 190 * 1024)


def peak_extract(spec: torch.Tensor, conv: torch.nn.Conv2d) -> torch.Tensor:
    """(B, H, W) log-mel segments -> (B, F, Ho * W, 1) peak point cloud stored as node rows: min-max normalisation,
    position ramps, convolution and ReLU of GPUPeakExtractorv2.forward (peak_extractor.py:56-82) as one kernel."""
    _require_cuda(spec)
    return _PeakExtract.apply(spec.contiguous(), conv.weight, conv.bias, int(conv.stride[0]))


# --------------------------------------------------------------------------------------
# NT-Xent loss
# --------------------------------------------------------------------------------------

class _NTXent(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, inv_tau):
        lib = _native.load()
        n2, d = z.shape
        lse = torch.empty(n2, dtype=torch.float32, device=z.device)
        row_loss = torch.empty(n2, dtype=torch.float32, device=z.device)
        loss = torch.empty((), dtype=torch.float32, device=z.device)
        _call("ntxent_fwd", 2, dict(B=1, N=n2, C=d), lib.grafp_ntxent_fwd, z.device, z.data_ptr(), lse.data_ptr(),
              row_loss.data_ptr(), loss.data_ptr(), n2, d, float(inv_tau), _stream(z))
        ctx.save_for_backward(z, lse)
        ctx.inv_tau = float(inv_tau)
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        lib = _native.load()
        z, lse = ctx.saved_tensors
        n2, d = z.shape
        g = grad_loss.to(torch.float32).contiguous()
        dz = torch.empty_like(z)
        _call("ntxent_bwd", 1, dict(B=1, N=n2, C=d), lib.grafp_ntxent_bwd, z.device, z.data_ptr(), lse.data_ptr(),
              g.data_ptr(), dz.data_ptr(), n2, d, ctx.inv_tau, _stream(z))
        return dz, None


class _NTXentSharded(torch.autograd.Function):
    """NT-Xent over the global batch of a data-parallel run, each rank evaluating only its own anchor rows.

    ``z_local`` (2 B_local, d): this rank's interleaved embeddings.  Forward: all-gather z (every rank holds the same number
    of rows), the rank's rows of the loss against all rows, the scalar summed over the ranks.  Backward: all-gather of the
    per-row log-sum-exps, then d loss / d z for the rank's own rows - complete (the other ranks' anchors enter through
    P_ji), so there is no gradient exchange for z.  The result is scaled by the world size: the data-parallel reduction
    that follows (DistributedDataParallel, training.FlatGradients) AVERAGES the parameter gradients over the ranks, and
    the average of world x (each rank's share) is the exact gradient of the global-batch loss - the same contract as the
    replicated form in simclr/distributed.py."""

    @staticmethod
    def forward(ctx, z_local, inv_tau, group):
        import torch.distributed as dist
        lib = _native.load()
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        nl, d = z_local.shape
        n2 = nl * world
        z_all = torch.empty(n2, d, dtype=torch.float32, device=z_local.device)
        dist.all_gather_into_tensor(z_all, z_local.contiguous(), group=group)
        lo, hi = rank * nl, (rank + 1) * nl
        lse = torch.empty(n2, dtype=torch.float32, device=z_local.device)
        row_loss = torch.empty(n2, dtype=torch.float32, device=z_local.device)
        loss = torch.empty((), dtype=torch.float32, device=z_local.device)
        _call("ntxent_fwd", 2, dict(B=1, N=n2, C=d, rows=nl), lib.grafp_ntxent_rows_fwd, z_local.device, z_all.data_ptr(),
              lse.data_ptr(), row_loss.data_ptr(), loss.data_ptr(), n2, d, lo, hi, float(inv_tau), _stream(z_local))
        dist.all_reduce(loss, group=group)
        lse_all = torch.empty(n2, dtype=torch.float32, device=z_local.device)
        dist.all_gather_into_tensor(lse_all, lse[lo:hi].contiguous(), group=group)
        ctx.save_for_backward(z_all, lse_all)
        ctx.meta = (float(inv_tau), lo, hi, world)
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        lib = _native.load()
        z_all, lse_all = ctx.saved_tensors
        inv_tau, lo, hi, world = ctx.meta
        n2, d = z_all.shape
        g = grad_loss.to(torch.float32).contiguous()
        dz = torch.empty(hi - lo, d, dtype=torch.float32, device=z_all.device)
        _call("ntxent_bwd", 1, dict(B=1, N=n2, C=d, rows=hi - lo), lib.grafp_ntxent_rows_bwd, z_all.device, z_all.data_ptr(),
              lse_all.data_ptr(), g.data_ptr(), dz.data_ptr(), n2, d, lo, hi, inv_tau, float(world), _stream(z_all))
        return dz, None, None


def ntxent_sharded(z_local: torch.Tensor, tau: float, group=None) -> torch.Tensor:
    """Global-batch NT-Xent of a data-parallel run from this rank's (2 B_local, d) interleaved embeddings (same B_local on
    every rank); see :class:`_NTXentSharded`."""
    _require_cuda(z_local)
    return _NTXentSharded.apply(z_local.contiguous(), 1.0 / float(tau), group)


def ntxent_supported(z: torch.Tensor) -> bool:
    return z.is_cuda and z.dtype == torch.float32 and z.dim() == 2 and z.shape[0] % 2 == 0 and z.shape[1] % 4 == 0 \
        and z.shape[1] <= 256


def ntxent(z: torch.Tensor, tau: float) -> torch.Tensor:
    """NT-Xent loss of (2B, d) embeddings whose rows 2m / 2m + 1 are partners (reference: simclr/ntxent.py:17-29),
    forward and backward as fused kernels that never build the (2B, 2B) similarity matrix."""
    _require_cuda(z)
    return _NTXent.apply(z.contiguous(), 1.0 / float(tau))


# --------------------------------------------------------------------------------------
# train-mode BatchNorm fused with the ReLU / residual add that follows it
# --------------------------------------------------------------------------------------

# A/B switches, read once at import: GRAFP_FUSED_BN=0 runs the PyTorch BatchNorm modules, GRAFP_FOLD_BN=0 keeps the
# eval-mode BatchNorm unfolded (=2: folded, but without cuDNN's fused ReLU epilogue)
_FUSED_BN = os.environ.get("GRAFP_FUSED_BN", "1") != "0"
_FOLD_BN = os.environ.get("GRAFP_FOLD_BN", "1")


def _is_rows(t: torch.Tensor) -> bool:
    """(B, C, N, 1) tensor whose memory is (B*N, C) rows."""
    if t.dim() != 4 or t.shape[3] != 1:
        return False
    B, C, N, _ = t.shape
    s = t.stride()
    return (C == 1 or s[1] == 1) and (N == 1 or s[2] == C) and (B == 1 or s[0] == N * C)


def _bn_kernel_ok(t: torch.Tensor, C: int) -> bool:
    """Envelope of the fused BatchNorm kernels: fp32 or bf16 rows, 16 bytes of channels per thread (4 fp32 / 8 bf16),
    C / that a power of two."""
    if t.dtype == torch.float32:
        v = 4
    elif t.dtype == torch.bfloat16:
        v = 8
    else:
        return False
    return C % v == 0 and ((C // v) & (C // v - 1)) == 0


def _bn_fwd_call(lib, h, residual, weight, bias, running_mean, running_var, eps, momentum, relu, conv_bias=None, nbt=None,
                 moments_ws=None):
    """Train-mode BatchNorm forward on rows ``h``.  ``moments_ws``: a workspace that already holds the per-channel raw
    moments of ``h`` (written by :func:`_conv1x1_stats_call`) - the statistics pass is skipped."""
    B, C, N, _ = h.shape
    out = _new_rows(B, C, N, h)
    save_mean = torch.empty(C, dtype=torch.float32, device=h.device)
    save_invstd = torch.empty(C, dtype=torch.float32, device=h.device)
    ws_bytes = lib.grafp_bn_workspace_bytes(C)
    ws = moments_ws if moments_ws is not None else torch.empty(ws_bytes, dtype=torch.uint8, device=h.device)
    name, fn = (("bn_train_fwd", lib.grafp_bn_train_fwd) if moments_ws is None
                else ("bn_apply_fwd", lib.grafp_bn_train_fwd_from_moments))
    _call(name, 2 if moments_ws is None else 1,
          dict(B=B, N=N, C=C, relu=int(relu), res=int(residual is not None), dtype=_dtype_code(h)),
          fn, h.device, h.data_ptr(), residual.data_ptr() if residual is not None else None,
          weight.data_ptr(), bias.data_ptr(),
          running_mean.data_ptr() if running_mean is not None else None,
          running_var.data_ptr() if running_var is not None else None,
          conv_bias.data_ptr() if conv_bias is not None and running_mean is not None else None,
          nbt.data_ptr() if nbt is not None else None,
          out.data_ptr(), save_mean.data_ptr(), save_invstd.data_ptr(), B * N, C, float(eps), float(momentum),
          int(relu), _dtype_code(h), ws.data_ptr(), ws_bytes, _stream(h))
    return out, save_mean, save_invstd


def _bn_bwd_call(lib, g, h, weight, bias, save_mean, save_invstd, relu, want_colsum):
    B, C, N, _ = h.shape
    dh = _new_rows(B, C, N, h)
    dweight = torch.empty(C, dtype=torch.float32, device=h.device)
    dbias = torch.empty(C, dtype=torch.float32, device=h.device)
    colsum = torch.empty(C, dtype=torch.float32, device=h.device) if want_colsum else None
    ws_bytes = lib.grafp_bn_workspace_bytes(C)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=h.device)
    _call("bn_train_bwd", 2, dict(B=B, N=N, C=C, relu=int(relu), dtype=_dtype_code(h)), lib.grafp_bn_train_bwd, h.device,
          g.data_ptr(), h.data_ptr(), weight.data_ptr(), bias.data_ptr(), save_mean.data_ptr(),
          save_invstd.data_ptr(), dh.data_ptr(), dweight.data_ptr(), dbias.data_ptr(),
          colsum.data_ptr() if colsum is not None else None, B * N, C, int(relu), _dtype_code(h),
          ws.data_ptr(), ws_bytes, _stream(h))
    return dh, dweight, dbias, colsum


def conv1x1_stats_ok(x: torch.Tensor, cw: torch.Tensor, conv_args) -> bool:
    """Whether the 1x1 convolution ``cw`` over rows ``x`` takes the tcgen05 GEMM with the BatchNorm statistics in its
    epilogue (``grafp_conv1x1_bn_stats_fwd``): dense or grouped, unit stride, no padding, and - for fp32 - TF32 convolutions allowed
    (``torch.backends.cudnn.allow_tf32``, PyTorch's default: the kernel computes what cuDNN computes then; with TF32
    off the convolution stays cuDNN's fp32 and the BatchNorm takes its own statistics pass).  Option ``conv_gemm``
    (``GRAFP_CONV_GEMM``): 1 = on where it wins (default), 2 = always, 0 = off (A/B)."""
    stride, padding, dilation, groups = conv_args
    if get_option("conv_gemm") == 0:
        return False
    if not (tuple(stride) == (1, 1) and tuple(padding) == (0, 0) and tuple(cw.shape[2:]) == (1, 1)):
        return False
    if x.dtype == torch.float32 and not torch.backends.cudnn.allow_tf32:
        return False
    if x.dtype not in (torch.float32, torch.bfloat16) or not (x.is_cuda and _is_rows(x)):
        return False
    B, Cin, N, _ = x.shape
    # measured (profiles/, scripts/bench_kernels.py gemm): with a k-range of more than 2 KB per row (Cin / groups > 512
    # fp32 / 1024 bf16) the layer is tensor- / L2-bound, cuDNN's larger tiles win by more than the saved statistics pass;
    # conv_gemm = 2 forces
    if get_option("conv_gemm") != 2 and (Cin // groups) * x.element_size() > 2048:
        return False
    return bool(_native.load().grafp_conv1x1_bn_stats_supported(B * N, Cin, cw.shape[0], groups, _dtype_code(x)))


def _conv1x1_stats_call(lib, x, cw_x, groups=1):
    """h = conv1x1(x, cw_x, groups) as rows, plus the BatchNorm workspace holding sum h / sum h^2 per channel."""
    B, Cin, N, _ = x.shape
    Cout = cw_x.shape[0]
    h = _new_rows(B, Cout, N, x)
    w = cw_x.reshape(Cout, Cin // groups)
    if not w.is_contiguous():
        w = w.contiguous()
    ws_bytes = lib.grafp_bn_workspace_bytes(Cout)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    _call("conv1x1_bn_stats_fwd", 1, dict(B=B, N=N, Cin=Cin, Cout=Cout, groups=groups, dtype=_dtype_code(x)),
          lib.grafp_conv1x1_bn_stats_fwd, x.device, x.data_ptr(), w.data_ptr(), h.data_ptr(), B * N, Cin, Cout, groups,
          _dtype_code(x), ws.data_ptr(), ws_bytes, _stream(x))
    return h, ws


class _BatchNormTrain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, residual, weight, bias, running_mean, running_var, eps, momentum, relu, nbt):
        lib = _native.load()
        out, save_mean, save_invstd = _bn_fwd_call(lib, x, residual, weight, bias, running_mean, running_var, eps,
                                                   momentum, relu, None, nbt)
        ctx.save_for_backward(x, weight, bias, save_mean, save_invstd)
        ctx.relu = bool(relu)
        ctx.has_res = residual is not None
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _native.load()
        x, weight, bias, save_mean, save_invstd = ctx.saved_tensors
        g = as_rows(grad_out.to(x.dtype))
        dx, dweight, dbias, _ = _bn_bwd_call(lib, g, x, weight, bias, save_mean, save_invstd, ctx.relu, False)
        return (dx, (grad_out if ctx.has_res else None), dweight.to(weight.dtype), dbias.to(bias.dtype), None, None, None,
                None, None, None)


def _nbt(bn: torch.nn.BatchNorm2d) -> Optional[torch.Tensor]:
    """The module's num_batches_tracked counter when the kernel can bump it in place (a CUDA int64 scalar)."""
    t = bn.num_batches_tracked if bn.track_running_stats else None
    if t is not None and not (t.is_cuda and t.dtype == torch.int64):
        t.add_(1)
        return None
    return t


def batch_norm_act(x: torch.Tensor, bn: torch.nn.BatchNorm2d, relu: bool = False,
                   residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``relu(bn(x))`` / ``bn(x) + residual`` / ``bn(x)`` for node rows (B, C, N, 1).

    In training mode on a CUDA fp32 / bf16 channels-last tensor this is one fused pair of kernels (statistics +
    apply; the backward recomputes the ReLU mask, so no intermediate is kept); the module's running
    statistics and ``num_batches_tracked`` are updated exactly like ``nn.BatchNorm2d`` does.  Anything else
    (eval mode, other dtypes / layouts / channel counts, or GRAFP_FUSED_BN=0 - an A/B switch) runs the module
    and the PyTorch ops unchanged.
    """
    fused = (_FUSED_BN and bn.training and x.is_cuda and _is_rows(x) and bn.affine
             and bn.momentum is not None and not (relu and residual is not None)
             and _bn_kernel_ok(x, x.shape[1]) and x.shape[0] * x.shape[2] > 1
             and (residual is None or (residual.shape == x.shape and residual.dtype == x.dtype and _is_rows(residual)))
             and bn.weight.dtype == torch.float32)
    if not fused:
        if _FUSED_BN and bn.training and x.is_cuda and x.numel() >= (1 << 20):
            _warn_once(f"bn-fallback-{x.dtype}-{x.shape[1]}-{_is_rows(x)}",
                       f"train-mode BatchNorm on a {tuple(x.shape)} {x.dtype} tensor runs the PyTorch modules (the fused kernels "
                       "need fp32 / bf16 node rows with C / (16 bytes of channels) a power of two and an affine BatchNorm)")
        y = bn(x)
        if residual is not None:
            y = y + residual
        return torch.relu(y) if relu else y
    rm = bn.running_mean if bn.track_running_stats else None
    rv = bn.running_var if bn.track_running_stats else None
    return _BatchNormTrain.apply(x, residual, bn.weight, bn.bias, rm, rv, bn.eps, bn.momentum, relu, _nbt(bn))


def grad_with_param_strides(grad: torch.Tensor, param: torch.Tensor) -> torch.Tensor:
    """cuDNN returns the weight gradient of a convolution over channels-last activations with channels-last strides;
    for a (Cout, Cin, 1, 1) weight that is the same memory as the contiguous layout, but DistributedDataParallel
    compares strides literally ("Grad strides do not match bucket view strides") and then copies the gradient into
    its bucket instead of aliasing it.  Re-stride (no copy) when the two layouts are the same bytes."""
    if grad.stride() != param.stride() and grad.shape == param.shape and grad.is_contiguous() and param.is_contiguous():
        return grad.as_strided(param.shape, param.stride())
    return grad


def match_grad_strides(module: torch.nn.Module) -> None:
    """Register :func:`grad_with_param_strides` on every convolution weight of ``module``."""
    for m in module.modules():
        if isinstance(m, torch.nn.Conv2d) and m.weight.requires_grad:
            m.weight.register_hook(lambda g, p=m.weight: grad_with_param_strides(g, p))


def _dense_form_of_grouped(x: torch.Tensor, cw: torch.Tensor, groups: int) -> bool:
    """cuDNN runs the bf16 channels-last grouped 1x1 convolution of BasicConv (groups = 4) through
    conv2d_grouped_direct_kernel at 29 ms per call (B = 512; torch.profiler, profiles/), 80 % of a bf16 training step.
    The same map as a DENSE 1x1 convolution with a block-diagonal weight is a plain tensor-core GEMM: 4x the flops on
    zeros, still HBM-bound at these channel counts, and exact (the added terms are products with 0)."""
    # (fp32: cuDNN's grouped kernels are fine; taking the dense form there just to put the layer on the tcgen05 GEMM with
    #  the statistics epilogue was measured - 106.3 vs 105.6 ms per step - and dropped)
    return groups > 1 and x.dtype == torch.bfloat16 and tuple(cw.shape[2:]) == (1, 1)


def _block_diag_weight(cw: torch.Tensor, groups: int) -> torch.Tensor:
    """(Cout, Cin / groups, 1, 1) grouped weight -> the (Cout, Cin, 1, 1) block-diagonal dense weight."""
    Cout, Cg = cw.shape[0], cw.shape[1]
    return torch.block_diag(*cw.reshape(groups, Cout // groups, Cg)).reshape(Cout, Cg * groups, 1, 1)


def _block_diag_grad(dw: torch.Tensor, groups: int) -> torch.Tensor:
    """Diagonal blocks of the dense weight gradient -> the grouped weight's gradient (Cout, Cin / groups, 1, 1)."""
    Cout, Cin = dw.shape[0], dw.shape[1]
    Og, Cg = Cout // groups, Cin // groups
    d = dw.reshape(groups, Og, groups, Cg)
    idx = torch.arange(groups, device=dw.device)
    return d[idx, :, idx, :].reshape(Cout, Cg, 1, 1)


class _ConvBatchNormTrain(torch.autograd.Function):
    """Conv2d(1x1, bias) -> train-mode BatchNorm [-> ReLU | + residual] as one autograd node: the convolution stays
    cuDNN, the BatchNorm is the fused kernel pair, and the convolution's bias gradient - the per-channel sum of the
    BatchNorm's input gradient - comes out of the BatchNorm backward instead of a separate reduction pass over that
    gradient (aten::convolution_backward is asked for the input and weight gradients only).

    The convolution itself runs WITHOUT its bias: a per-channel constant in front of a train-mode BatchNorm cancels in
    (h - mean(h)), so adding it is a full read + write pass over the activations (cuDNN does not fuse it: ~105
    broadcast-add launches, ~10 ms of the 120 ms step in the ncu launch list) that changes nothing but rounding.  The
    statistics are taken of the bias-free output and the bias is added back where it is visible: the running mean.

    ``x`` may be bf16 (torch.autocast) with fp32 parameters: the convolution then runs on a bf16 copy of the weight
    (what autocast itself does) and the weight gradient is returned in the parameter's dtype."""

    @staticmethod
    def forward(ctx, x, residual, cw, cb, weight, bias, running_mean, running_var, eps, momentum, relu, conv_args, nbt):
        lib = _native.load()
        stride, padding, dilation, groups = conv_args
        moments_ws = None
        with torch.autocast("cuda", enabled=False):
            cw_x = cw if cw.dtype == x.dtype else cw.to(x.dtype)
            if conv1x1_stats_ok(x, cw_x, conv_args):
                # own GEMM (grouped: block-diagonal MMA schedule): the convolution output and its per-channel moments in
                # one pass (conv_gemm.cu)
                h, moments_ws = _conv1x1_stats_call(lib, x, cw_x, groups)
            else:
                dense = _dense_form_of_grouped(x, cw, groups)
                w_eff = _block_diag_weight(cw_x, groups) if dense else cw_x
                h = torch.nn.functional.conv2d(x, w_eff, None, stride, padding, dilation, 1 if dense else groups)   # bias: see the class docstring
        if not _is_rows(h):
            h = as_rows(h)
        if residual is not None and residual.shape != h.shape:
            raise RuntimeError(f"grafp_b200.conv_batch_norm_act: residual {tuple(residual.shape)} does not match the "
                               f"convolution output {tuple(h.shape)}")
        # running_mean <- (1 - m) running_mean + m (mean(h) + cb): the kernel adds the bias the convolution skipped
        cb32 = cb.detach() if cb is not None else None
        if cb32 is not None and cb32.dtype != torch.float32:
            cb32 = cb32.float()
        out, save_mean, save_invstd = _bn_fwd_call(lib, h, residual, weight, bias, running_mean, running_var, eps,
                                                   momentum, relu, cb32, nbt, moments_ws)
        ctx.save_for_backward(x, cw, h, weight, bias, save_mean, save_invstd)
        ctx.relu = bool(relu)
        ctx.has_res = residual is not None
        ctx.has_cb = cb is not None
        ctx.conv_args = conv_args
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _native.load()
        x, cw, h, weight, bias, save_mean, save_invstd = ctx.saved_tensors
        stride, padding, dilation, groups = ctx.conv_args
        g = as_rows(grad_out.to(h.dtype))
        dh, dweight, dbias, dcb = _bn_bwd_call(lib, g, h, weight, bias, save_mean, save_invstd, ctx.relu, ctx.has_cb)
        cw_x = cw if cw.dtype == x.dtype else cw.to(x.dtype)
        if _dense_form_of_grouped(x, cw, groups):
            dx, dcw, _ = torch.ops.aten.convolution_backward(
                dh, x, _block_diag_weight(cw_x, groups), None, list(stride), list(padding), list(dilation), False, [0, 0], 1,
                [ctx.needs_input_grad[0], ctx.needs_input_grad[2], False])
            if dcw is not None:
                dcw = _block_diag_grad(dcw, groups)
        else:
            dx, dcw, _ = torch.ops.aten.convolution_backward(
                dh, x, cw_x, None, list(stride), list(padding), list(dilation), False, [0, 0], groups,
                [ctx.needs_input_grad[0], ctx.needs_input_grad[2], False])
        if dcw is not None and dcw.dtype != cw.dtype:
            dcw = dcw.to(cw.dtype)
        if dcw is not None:
            dcw = grad_with_param_strides(dcw, cw)
        return (dx, (grad_out if ctx.has_res else None), dcw, dcb, dweight.to(weight.dtype), dbias.to(bias.dtype),
                None, None, None, None, None, None, None)


def _fold_eval_ok(x: torch.Tensor, bn: torch.nn.BatchNorm2d) -> bool:
    return (_FOLD_BN != "0" and not bn.training and not torch.is_grad_enabled()
            and x.is_cuda and x.dtype == torch.float32 and bn.track_running_stats and bn.running_var is not None
            and bn.affine)


def _folded_conv_bn_eval(x, cw, cb, bn, relu, residual, conv_args):
    """Inference form of conv -> BatchNorm(eval) [-> ReLU | + residual] (generate.py's path):
    the BatchNorm's affine map is folded into the convolution's weight and bias (w' = w * g / sqrt(var + eps),
    b' = (b - mean) * g / sqrt(var + eps) + beta), so the layer is one cuDNN convolution - with the ReLU in its
    epilogue where cuDNN offers that - instead of three passes over the activations."""
    stride, padding, dilation, groups = conv_args
    scale = bn.weight * torch.rsqrt(bn.running_var + bn.eps)
    w = cw * scale.view(-1, 1, 1, 1)
    b = bn.bias - bn.running_mean * scale if cb is None else (cb - bn.running_mean) * scale + bn.bias
    if relu and residual is None and _FOLD_BN != "2":
        try:
            y = torch.cudnn_convolution_relu(x, w, b, list(stride), list(padding), list(dilation), groups)
            return y if _is_rows(y) else as_rows(y)
        except RuntimeError:
            pass
    y = torch.nn.functional.conv2d(x, w, b, stride, padding, dilation, groups)
    if residual is not None:
        y = y.add_(residual)
    return y.relu_() if relu else y


def conv_batch_norm_act(x: torch.Tensor, conv: torch.nn.Conv2d, bn: torch.nn.BatchNorm2d, relu: bool = False,
                        residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``relu(bn(conv(x)))`` / ``bn(conv(x)) + residual`` / ``bn(conv(x))`` for node rows (B, C, N, 1).

    With a 1x1 convolution in training mode (CUDA, fp32 / bf16, channels-last) the three layers are one autograd
    node (see _ConvBatchNormTrain); otherwise the convolution runs as the module and the rest goes through
    :func:`batch_norm_act`, which applies its own envelope checks.
    """
    if (_fold_eval_ok(x, bn) and conv.padding_mode == "zeros" and isinstance(conv.padding, tuple) and not conv.transposed
            and conv.weight.dtype == torch.float32):
        return _folded_conv_bn_eval(x, conv.weight, conv.bias, bn, relu, residual,
                                    (conv.stride, conv.padding, conv.dilation, conv.groups))
    ok = (conv.kernel_size == (1, 1) and conv.padding_mode == "zeros"
          and isinstance(conv.padding, tuple) and not conv.transposed)
    if not ok:
        return batch_norm_act(conv(x), bn, relu=relu, residual=residual)
    return pointwise_conv_batch_norm_act(x, conv.weight, conv.bias, bn, relu, residual,
                                         (conv.stride, conv.padding, conv.dilation, conv.groups))


def pointwise_conv_batch_norm_act(x, cw, cb, bn, relu=False, residual=None, conv_args=((1, 1), (0, 0), (1, 1), 1)):
    """:func:`conv_batch_norm_act` on explicit 1x1 convolution weights ``cw`` (Cout, Cin / groups, 1, 1) and bias."""
    C = cw.shape[0]
    if _fold_eval_ok(x, bn) and cw.dtype == torch.float32:
        return _folded_conv_bn_eval(x, cw, cb, bn, relu, residual, conv_args)
    stride, padding, dilation, groups = conv_args
    unit = tuple(stride) == (1, 1) and tuple(padding) == (0, 0) and tuple(cw.shape[2:]) == (1, 1)
    fused = (_FUSED_BN and bn.training and x.is_cuda and _is_rows(x)
             and bn.affine and bn.momentum is not None and not (relu and residual is not None)
             and _bn_kernel_ok(x, C) and x.shape[0] * x.shape[2] > 1
             # the residual must have the convolution's output shape (B, C, N, 1); known up front for stride-1 1x1 convs
             and (residual is None or (unit and residual.dtype == x.dtype and _is_rows(residual)
                                       and tuple(residual.shape) == (x.shape[0], C, x.shape[2], 1)))
             and cw.dtype == torch.float32 and bn.weight.dtype == torch.float32
             and torch.is_grad_enabled())
    if not fused:
        return batch_norm_act(torch.nn.functional.conv2d(x, cw, cb, stride, padding, dilation, groups), bn, relu=relu,
                              residual=residual)
    rm = bn.running_mean if bn.track_running_stats else None
    rv = bn.running_var if bn.track_running_stats else None
    return _ConvBatchNormTrain.apply(x, residual, cw, cb, bn.weight, bn.bias, rm, rv, bn.eps, bn.momentum, relu, conv_args,
                                     _nbt(bn))


class _MeanOverNodes(torch.autograd.Function):
    """(B, C, N, 1) node rows -> (B, C, 1, 1), the mean over the nodes.  Only the backward differs from ``torch.mean``:
    it writes the broadcast gradient as node rows.  PyTorch's own backward materialises it NCHW-contiguous, and that
    layout then travels down the residual chain of the last stage - every BatchNorm backward there first copied its
    incoming gradient to rows and every residual accumulation ran as a strided add (~2.5 ms per training step)."""

    @staticmethod
    def forward(ctx, x):
        ctx.shape = tuple(x.shape)
        return torch.mean(x, dim=2, keepdim=True)

    @staticmethod
    def backward(ctx, grad):
        B, C, N, _ = ctx.shape
        dx = torch.empty(ctx.shape, dtype=grad.dtype, device=grad.device, memory_format=torch.channels_last)
        dx.copy_((grad / N).expand(ctx.shape))
        return dx


def mean_over_nodes(x: torch.Tensor) -> torch.Tensor:
    """``torch.mean(x, dim=2, keepdim=True)`` for (B, C, N, 1) node rows, with a row-layout gradient."""
    if x.dim() == 4 and x.shape[3] == 1 and x.is_cuda and torch.is_grad_enabled() and x.requires_grad:
        return _MeanOverNodes.apply(x)
    return torch.mean(x, dim=2, keepdim=True)


class _DownsampleTaps(torch.autograd.Function):
    """(B, C, N, 1) node rows -> (B, 3C, N/2, 1) tap rows (x[2n'-1], x[2n'], x[2n'+1]) of the Downsample block."""

    @staticmethod
    def forward(ctx, x):
        lib = _native.load()
        x = as_rows(x)
        B, C, N, _ = x.shape
        taps = _new_rows(B, 3 * C, N // 2, x)
        _call("downsample_taps_fwd", 1, dict(B=B, N=N, C=C, dtype=_dtype_code(x)), lib.grafp_downsample_taps_fwd, x.device,
              x.data_ptr(), taps.data_ptr(), B, N, C, _dtype_code(x), _stream(x))
        ctx.shape = (B, C, N)
        return taps

    @staticmethod
    def backward(ctx, grad):
        lib = _native.load()
        B, C, N = ctx.shape
        g = as_rows(grad)
        dx = _new_rows(B, C, N, g)
        _call("downsample_taps_bwd", 1, dict(B=B, N=N, C=C, dtype=_dtype_code(g)), lib.grafp_downsample_taps_bwd, g.device,
              g.data_ptr(), dx.data_ptr(), B, N, C, _dtype_code(g), _stream(g))
        return dx


def downsample_rows(x: torch.Tensor, conv: torch.nn.Conv2d, bn: torch.nn.BatchNorm2d) -> Optional[torch.Tensor]:
    """``bn(conv(x))`` for the Downsample block (graph_encoder.py:16-28: Conv2d(3x3, stride 2, padding 1) over a
    (B, C, N, 1) node list), or None when the shape is not that case.

    The image is one pixel wide, so only the middle column of the 3x3 kernel ever meets data: the layer is a 3-tap,
    stride-2 convolution along the node axis, out[n'] = sum_t W[:, :, t, 1] x[2n' + t - 1].  cuDNN's strided 3x3
    data-gradient kernels run it at a few % of the hardware; here the three input rows of every output row are
    laid side by side (one copy, 1.5x the input) and the layer becomes a 1x1 convolution with 3C input channels -
    the same cuDNN GEMM kernels (and TF32 policy) as every other 1x1 convolution of the encoder - followed by the
    fused BatchNorm.  The other six kernel taps multiply zero padding in the reference too: their weight gradients
    are exactly zero in both.
    """
    if not (x.dim() == 4 and x.shape[3] == 1 and x.shape[2] % 2 == 0 and x.shape[2] >= 2 and _is_rows(x)
            and conv.kernel_size == (3, 3) and conv.stride == (2, 2) and conv.padding == (1, 1)
            and conv.dilation == (1, 1) and conv.groups == 1 and conv.padding_mode == "zeros"
            and _FUSED_BN):
        return None
    B, C, N, _ = x.shape
    if x.is_cuda and x.dtype in _DTYPES and (C * x.element_size()) % 16 == 0:
        a = _DownsampleTaps.apply(x)                                     # one shifted copy (downsample.cu)
    else:
        rows = x.permute(0, 2, 3, 1).reshape(B, N // 2, 2 * C)           # [x[2n'], x[2n'+1]] per output row: a view
        prev = torch.nn.functional.pad(rows[:, :-1, C:], (0, 0, 1, 0))   # x[2n'-1], zero row in front of every segment
        taps = torch.cat([prev, rows], dim=2)                            # (B, N/2, 3C)
        a = taps.view(B, N // 2, 1, 3 * C).permute(0, 3, 1, 2)           # logical (B, 3C, N/2, 1), rows in memory
    w = conv.weight[:, :, :, 1]                                      # (Cout, Cin, 3): the middle kernel column
    cw = torch.cat([w[:, :, 0], w[:, :, 1], w[:, :, 2]], dim=1).reshape(conv.out_channels, 3 * C, 1, 1)
    return pointwise_conv_batch_norm_act(a, cw, conv.bias, bn)
