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


class GraphedTrainStep:
    """CUDA-graph replay of ``loss = loss_fn(*net(*inputs)); loss.backward(); optimizer.step()``.

    net:        the module to call (a ``SimCLR`` or its ``DistributedDataParallel`` wrapper).
    optimizer:  must have been built with ``capturable=True`` (its step counters then live on the device).
    loss_fn:    maps the module's outputs to a scalar loss.
    example_inputs: device tensors fixing shapes and dtypes; ``__call__`` copies new data into static copies of them.
    autocast_dtype: e.g. ``torch.bfloat16`` to run the forward under ``torch.autocast`` (parameters stay fp32).
    """

    def __init__(self, net: torch.nn.Module, optimizer: torch.optim.Optimizer, loss_fn: Callable[..., torch.Tensor],
                 example_inputs: Sequence[torch.Tensor], autocast_dtype: Optional[torch.dtype] = None, warmup: int = 3):
        if not all(t.is_cuda for t in example_inputs):
            raise RuntimeError("grafp_b200.GraphedTrainStep: CUDA tensors only (there is no CPU path)")
        for group in optimizer.param_groups:
            if not group.get("capturable", False):
                raise RuntimeError("grafp_b200.GraphedTrainStep: build the optimizer with capturable=True")
        self.net, self.optimizer, self.loss_fn = net, optimizer, loss_fn
        self.autocast_dtype = autocast_dtype
        self.static_in = [t.detach().clone() for t in example_inputs]
        dev = example_inputs[0].device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):  # optimizer state, per-device kernel attributes, cuDNN plans: not capturable
                self._eager_step()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)  # gradients are then allocated inside the capture, from the graph's pool
        with torch.cuda.graph(self.graph):
            self.static_loss = self._forward_backward()
            optimizer.step()

    def _forward_backward(self) -> torch.Tensor:
        with torch.autocast("cuda", dtype=self.autocast_dtype or torch.bfloat16, enabled=self.autocast_dtype is not None):
            outputs = self.net(*self.static_in)
        loss = self.loss_fn(*outputs)
        loss.backward()
        return loss

    def _eager_step(self) -> torch.Tensor:
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
