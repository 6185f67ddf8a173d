"""Data parallelism over samples: one process per GPU, one flat gradient all-reduce per step.

The position-attention path has no exchange step (samples are independent; for shared meshes the
attention weights are simply recomputed per rank), so the only collective of a training step is the
gradient reduction.  PiT models are small (8.6 k - 1.27 M parameters), hence a single flat fp32 bucket:
one NCCL all-reduce over NVLink/NVSwitch, latency-bound.  The reduction is a SUM, not a mean, because the
reference loss sums over the batch (utils.py:98): the summed gradient equals the gradient of the global batch.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
import torch.distributed as dist


class FlatGradients:
    """Owns one contiguous buffer; every parameter's ``.grad`` is a view into it.

    Two ways to run a step:
      * ``zero()`` ... backward ... ``all_reduce()``: gradients accumulate into the views (one add kernel per
        parameter, as autograd does for a defined ``.grad``);
      * ``release()`` ... backward ... ``gather()`` + ``all_reduce()``: ``.grad`` is None during backward -- what
        ``optimizer.zero_grad()`` does by default -- so autograd just hands over each gradient (no fill, no adds);
        ``gather`` then packs them into the buffer with one multi-tensor copy and re-attaches the views.  With a
        single rank the pack is not needed at all and the optimizer consumes the gradients where they are.
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], world_size: Optional[int] = None, group=None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        first = self.params[0]
        total = sum(p.numel() for p in self.params)
        self.buffer = torch.zeros(total, dtype=first.dtype, device=first.device)
        offset = 0
        self.views = []
        for p in self.params:
            if p.dtype != first.dtype or p.device != first.device:
                raise ValueError("FlatGradients needs parameters of one dtype on one device")
            self.views.append(self.buffer[offset:offset + p.numel()].view_as(p))
            p.grad = self.views[-1]
            offset += p.numel()
        self.group = group
        if world_size is None:
            world_size = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.world_size = world_size

    def zero(self) -> None:
        """Use instead of ``optimizer.zero_grad()`` (which would detach the views)."""
        self.buffer.zero_()

    def release(self) -> None:
        """Detach the views: the next backward writes fresh gradient tensors instead of accumulating."""
        for p in self.params:
            p.grad = None

    def gather(self) -> None:
        """Pack the gradients produced by backward into the flat buffer and make ``.grad`` the views again."""
        have = [(v, p.grad) for v, p in zip(self.views, self.params) if p.grad is not None and p.grad is not v]
        for v, p in zip(self.views, self.params):
            if p.grad is None:
                v.zero_()
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        for v, p in zip(self.views, self.params):
            p.grad = v

    def all_reduce(self) -> None:
        if self.world_size > 1:
            dist.all_reduce(self.buffer, op=dist.ReduceOp.SUM, group=self.group)


def shard_range(n_items: int, rank: int, world_size: int) -> range:
    """Contiguous shard of ``n_items`` samples owned by ``rank`` (remainder spread over the first ranks)."""
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))
