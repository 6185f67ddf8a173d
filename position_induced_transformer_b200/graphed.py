"""CUDA-graph capture of a whole PiT training step.

PiT steps are short (a few milliseconds at Darcy-421, well under one at the 1-D configs) and consist of ~150
kernels, so launch latency and Python overhead are a large share of the wall clock.  The fused position-attention
ops never synchronise or read device data on the host, the mesh constants are cached before capture, and every
launch goes to torch's current stream -- so the complete step (zero grads, forward, loss, backward, gradient
all-reduce, Adam) can be captured once and replayed.  This is plain stream capture (torch.cuda.CUDAGraph), not a
tracing compiler: the replayed kernels are exactly the ones the eager step launched.
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch

from .data_parallel import FlatGradients


class GraphedTrainStep:
    """step(inputs, target) -> loss tensor (device), replaying a captured graph.

    `forward_loss(inputs, target)` must run the model and return the scalar loss using only capturable ops.
    Inputs are copied into static buffers, so callers may pass fresh tensors (device or pinned host) every step.
    """

    def __init__(self, params: Sequence[torch.nn.Parameter], forward_loss: Callable, optimizer,
                 example_inputs: Sequence[torch.Tensor], example_target: torch.Tensor, world_size: int = 1, warmup: int = 3):
        # `optimizer`: a torch optimizer (gradients are packed and summed with one NCCL all-reduce when world_size > 1), or a
        # fused_optimizer.FusedAllReduceAdam, whose step() does the cross-rank SUM and the update in one launch
        self.fused_reduce = hasattr(optimizer, "region_ptrs")
        self.flat = FlatGradients(params, 1 if self.fused_reduce else world_size)
        self.forward_loss = forward_loss
        self.optimizer = optimizer
        self.static_inputs = tuple(x.clone() for x in example_inputs)
        self.static_target = example_target.clone()
        self.loss = None
        self.graph = None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):          # eager warm-up on a side stream: fills the mesh cache, Adam state, cuBLAS handles
            for _ in range(warmup):
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._eager()

    def _eager(self):
        self.flat.release()                     # grads None, as optimizer.zero_grad() leaves them: no fill, no accumulation adds
        loss = self.forward_loss(self.static_inputs, self.static_target)
        loss.backward()
        if self.flat.world_size > 1:            # never with a FusedAllReduceAdam: its step() reduces
            self.flat.gather()                  # one multi-tensor copy into the flat bucket
            self.flat.all_reduce()
        self.optimizer.step()
        return loss.detach()

    def load(self, inputs: Sequence[torch.Tensor], target: torch.Tensor) -> None:
        """Copy a batch into the graph's static input buffers (the source may be reused as soon as this copy has run)."""
        for dst, src in zip(self.static_inputs, inputs):
            dst.copy_(src, non_blocking=True)
        self.static_target.copy_(target, non_blocking=True)

    def replay(self) -> torch.Tensor:
        self.graph.replay()
        return self.loss

    def __call__(self, inputs: Sequence[torch.Tensor], target: torch.Tensor) -> torch.Tensor:
        self.load(inputs, target)
        return self.replay()
