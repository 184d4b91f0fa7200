"""One SimCLR training step as a single CUDA graph ("CUDA-graph capture of the whole encoder step").

The reference's step (train.py:60-80: zero_grad, ``model(x_i, x_j)``, ``ntxent_loss``, ``backward``, ``optimizer.step``)
is ~3 000 kernel launches at batch 512.  With the hot path fused the GPU needs ~105 ms for them and the host ~100 ms to
issue them (torch.profiler, profiles/): every further kernel speed-up would be hidden behind the launch rate.
``GraphedTrainStep`` captures the step once - forward of both views, loss, backward, optimizer update - and replays it;
the C-ABI calls allocate nothing and never synchronise, so they capture as they are (workspaces come from the graph's
private pool, TMA descriptors are kernel parameters).  Numerically it IS the eager step: same kernels, same order.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch


class FlatGradients:
    """All gradients of a module in one flat buffer (``p.grad`` are views into it, autograd accumulates in place), with
    the data-parallel reduction DistributedDataParallel performs - the mean over the ranks - as ONE all-reduce of that
    buffer.  Unlike DistributedDataParallel's reducer it is plain stream-ordered work, so it can be captured in a CUDA
    graph together with the step.  The parameters are broadcast from rank 0 on construction."""

    def __init__(self, module: torch.nn.Module, process_group=None):
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("grafp_b200.FlatGradients: torch.distributed is not initialised")
        self.group = process_group
        self.world = dist.get_world_size(process_group)
        params = [p for p in module.parameters() if p.requires_grad]
        if any(not p.is_contiguous() for p in params) or len({p.dtype for p in params}) != 1:
            raise RuntimeError("grafp_b200.FlatGradients: needs contiguous parameters of one dtype")
        src = dist.get_global_rank(process_group, 0) if process_group is not None else 0
        with torch.no_grad():
            for p in module.parameters():
                dist.broadcast(p, src=src, group=process_group)
        self.flat = torch.zeros(sum(p.numel() for p in params), dtype=params[0].dtype, device=params[0].device)
        offset = 0
        for p in params:
            p.grad = self.flat[offset:offset + p.numel()].view(p.shape)
            offset += p.numel()

    def zero(self) -> None:
        self.flat.zero_()

    def all_reduce_mean(self) -> None:
        import torch.distributed as dist
        dist.all_reduce(self.flat, group=self.group)
        self.flat.div_(self.world)


class GraphedTrainStep:
    """CUDA-graph replay of ``loss = loss_fn(*net(*inputs)); loss.backward(); optimizer.step()``.

    net:        the module to call (a ``SimCLR`` or its ``DistributedDataParallel`` wrapper).
    optimizer:  must have been built with ``capturable=True`` (its step counters then live on the device).
    loss_fn:    maps the module's outputs to a scalar loss.
    example_inputs: device tensors fixing shapes and dtypes; ``__call__`` copies new data into static copies of them.
    autocast_dtype: e.g. ``torch.bfloat16`` to run the forward under ``torch.autocast`` (parameters stay fp32).
    """

    def __init__(self, net: torch.nn.Module, optimizer: torch.optim.Optimizer, loss_fn: Callable[..., torch.Tensor],
                 example_inputs: Sequence[torch.Tensor], autocast_dtype: Optional[torch.dtype] = None, warmup: int = 3,
                 process_group=None, data_parallel: bool = False):
        """data_parallel (one process per GPU, ``torch.distributed`` initialised with NCCL): ``net`` is the PLAIN module,
        not a DistributedDataParallel wrapper (its reducer cannot be captured: it touches the legacy stream).  The
        parameters are broadcast from rank 0 once, every gradient lives in one flat buffer, and the step ends with one
        all-reduce of that buffer - averaged over the ranks like DistributedDataParallel does - captured in the graph
        with everything else.  Buffers (BatchNorm running statistics) stay per replica."""
        if not all(t.is_cuda for t in example_inputs):
            raise RuntimeError("grafp_b200.GraphedTrainStep: CUDA tensors only (there is no CPU path)")
        for group in optimizer.param_groups:
            if not group.get("capturable", False):
                raise RuntimeError("grafp_b200.GraphedTrainStep: build the optimizer with capturable=True")
        self.net, self.optimizer, self.loss_fn = net, optimizer, loss_fn
        self.autocast_dtype = autocast_dtype
        self.static_in = [t.detach().clone() for t in example_inputs]
        dev = example_inputs[0].device
        self.world = 1
        self.flat_grad = None
        if data_parallel:
            if isinstance(net, torch.nn.parallel.DistributedDataParallel):
                raise RuntimeError("grafp_b200.GraphedTrainStep: pass the plain module, not its DistributedDataParallel wrapper")
            flat = FlatGradients(net, process_group)
            self.world = flat.world
            if self.world > 1:
                self.flat_grad = flat
            else:
                for p in net.parameters():
                    p.grad = None
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):  # optimizer state, per-device kernel attributes, cuDNN plans, the NCCL
                self._eager_step()           # communicator: none of that can be set up inside a capture
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        if self.world == 1:
            optimizer.zero_grad(set_to_none=True)  # gradients are then allocated inside the capture, from the graph's pool
        # (thread_local: the NCCL watchdog thread polls events while this thread captures)
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local" if self.world > 1 else "global"):
            self.static_loss = self._forward_backward()
            optimizer.step()

    def _forward_backward(self) -> torch.Tensor:
        if self.flat_grad is not None:
            self.flat_grad.zero()
        with torch.autocast("cuda", dtype=self.autocast_dtype or torch.bfloat16, enabled=self.autocast_dtype is not None):
            outputs = self.net(*self.static_in)
        loss = self.loss_fn(*outputs)
        loss.backward()
        if self.flat_grad is not None:
            self.flat_grad.all_reduce_mean()
        return loss

    def _eager_step(self) -> torch.Tensor:
        if self.flat_grad is None:
            self.optimizer.zero_grad(set_to_none=True)
        loss = self._forward_backward()
        self.optimizer.step()
        return loss

    def __call__(self, *inputs: torch.Tensor) -> torch.Tensor:
        """Copy ``inputs`` (host or device) into the static buffers, replay the step, return the (static) loss tensor."""
        for dst, src in zip(self.static_in, inputs):
            if src is not dst:
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_loss
