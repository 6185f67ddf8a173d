"""Data-parallel optimizer step in one launch: gradient SUM over NVLink peer memory fused with Adam.

``FusedAllReduceAdam`` replaces the three things a data-parallel PiT step does after backward -- pack the gradients into a
flat bucket, ``ncclAllReduce`` it, run the optimizer -- by ONE kernel per rank (``pit_allreduce_adam``,
csrc/allreduce_adam.cuh).  Every rank maps every other rank's gradient bucket through ``torch.distributed._symmetric_memory``
(NVSwitch peer access), publishes its own gradients there, waits on system-scope flags and sums the buckets itself while applying
``torch.optim.Adam``'s update (the optimizer of every reference script, train_darcy.py:115) to flat parameter and moment buffers.
The models are small (8.6 k - 1.27 M parameters), so this step is pure latency; NCCL alone costs ~60 us of a 0.9 ms step at 8 GPUs.

With one rank it is simply a single-launch Adam.  torch.distributed is plumbing here: rendezvous of the symmetric buffers only.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Optional

import torch
import torch.distributed as dist

from . import _cabi

MAX_TENSORS, MAX_WORLD = 64, 16


class FusedAllReduceAdam:
    """Adam over a flat parameter buffer with the cross-rank gradient SUM inside the same kernel.

    ``params`` become views into one flat buffer (``p.data`` is re-pointed; values are preserved).  ``step()`` consumes ``p.grad``
    wherever autograd left it (``None`` counts as zero) and is CUDA-graph capturable: the step counter, the bucket parity and the
    learning rate live on the device.  ``lr`` may be changed between steps with ``set_lr`` (a device write, so a scheduler works
    under graph replay as well).
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, group=None):
        self.params = [p for p in params if p.requires_grad]
        if not 1 <= len(self.params) <= MAX_TENSORS:
            raise ValueError(f"FusedAllReduceAdam handles 1..{MAX_TENSORS} parameter tensors, got {len(self.params)}")
        first = self.params[0]
        if not first.is_cuda or any(p.dtype != torch.float32 or p.device != first.device for p in self.params):
            raise ValueError("FusedAllReduceAdam needs float32 CUDA parameters on one device")
        self.device = first.device
        self.total = sum((p.numel() + 3) // 4 * 4 for p in self.params)     # every tensor starts on a 16-byte boundary
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        if self.world > MAX_WORLD:
            raise ValueError(f"at most {MAX_WORLD} ranks")
        self.betas, self.eps = (float(betas[0]), float(betas[1])), float(eps)
        with torch.no_grad():
            self.flat_param = torch.zeros(self.total, dtype=torch.float32, device=self.device)
            off = 0
            for p in self.params:
                view = self.flat_param[off:off + p.numel()].view_as(p)
                view.copy_(p.detach())
                p.data = view
                off += (p.numel() + 3) // 4 * 4
        self.exp_avg = torch.zeros_like(self.flat_param)
        self.exp_avg_sq = torch.zeros_like(self.flat_param)
        self.step_count = torch.zeros((), dtype=torch.int32, device=self.device)
        self.sync = torch.zeros(3, dtype=torch.int32, device=self.device)
        self.lr = torch.full((), float(lr), dtype=torch.float32, device=self.device)
        self.region, self.region_ptrs = None, [None] * MAX_WORLD
        if self.world > 1:
            import torch.distributed._symmetric_memory as symm
            pg = group if group is not None else dist.group.WORLD
            n = int(_cabi.lib.pit_allreduce_adam_region_floats(self.total))
            self.region = symm.empty(n, dtype=torch.float32, device=self.device)
            self.region.zero_()
            self._handle = symm.rendezvous(self.region, pg.group_name)
            ptrs = list(self._handle.buffer_ptrs)
            if len(ptrs) != self.world:
                raise RuntimeError("symmetric memory rendezvous returned an unexpected number of peers")
            self.region_ptrs[:self.world] = ptrs
            torch.cuda.synchronize(self.device)
            dist.barrier(group)                   # every region is zero-filled before anyone's first step

    # -- optimizer surface --------------------------------------------------------------------------------------
    def zero_grad(self, set_to_none: bool = True) -> None:
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def set_lr(self, lr: float) -> None:
        self.lr.fill_(float(lr))

    def step(self) -> None:
        a = _cabi.AllReduceAdam()
        a.world, a.rank, a.n_tensors = self.world, self.rank, len(self.params)
        keep = []
        for k, p in enumerate(self.params):
            g = p.grad
            if g is not None:
                if g.dtype != torch.float32 or not g.is_contiguous():
                    g = g.float().contiguous()
                keep.append(g)
                a.grad[k] = g.data_ptr()
            else:
                a.grad[k] = None
            a.numel[k] = p.numel()
        for r in range(self.world):
            a.region[r] = self.region_ptrs[r]
        a.param, a.exp_avg, a.exp_avg_sq = self.flat_param.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr()
        a.step, a.sync, a.lr = self.step_count.data_ptr(), self.sync.data_ptr(), self.lr.data_ptr()
        a.beta1, a.beta2, a.eps = self.betas[0], self.betas[1], self.eps
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib.pit_allreduce_adam(C.byref(a), torch.cuda.current_stream(self.device).cuda_stream), "pit_allreduce_adam")

    def peer_timeout(self) -> bool:
        """True if a step ever gave up waiting for a peer (2 s): the run is then invalid.  Synchronises."""
        return bool(int(self.sync[2]) != 0)
