"""CPU oracle for PiT position-attention -- TEST INFRASTRUCTURE ONLY.

This file restates, in plain dense fp32 torch ops on the CPU, the algorithm of the
reference's ``posatt*`` classes (``/root/reference/pit.py``).  It is the checker for the
CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product package
(``position_induced_transformer_b200``) never imports anything under ``oracle/``.

Parity pin: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the reference code itself: ``oracle/gen_golden.py``
imports ``/root/reference/pit.py`` in the build container and writes
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this file against them.

Two formulations are provided:

* ``dense_*``   -- the reference's own order of operations (materialise N x M, call
  ``torch.quantile``, ``softmax``, ``einsum``).  Follows pit.py:46-57 (batched),
  133-144 (fixed mesh), 190-200 (periodic 1-D), 247-258 (periodic 2-D).
* ``exact_*``   -- the order-statistic restatement the CUDA kernels implement
  (SURVEY.md section 8a steps 1-5): thresholds from the k-th smallest squared
  distances, per-head mask on the rounded product, softmax with the known maximum.
  ``tests`` assert both formulations give identical kept sets.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

VARIANTS = ("euclid", "periodic1d", "periodic2d")

# 0.25*pi*(1-1e-7): python double, rounded to fp32 when multiplied into an fp32 tensor
# (pit.py:48, 135, 196, 254).
SCALE_CONST = 0.25 * math.pi * (1 - 1e-7)
FLT_MAX = torch.finfo(torch.float32).max


def head_scale(lmda: torch.Tensor) -> torch.Tensor:
    """lmda (H,1,1) -> positive per-head scale s_h, same shape (pit.py:48)."""
    return torch.tan(SCALE_CONST * (1.0 + torch.sin(lmda)))


def period_of(mesh_in: torch.Tensor, variant: str) -> Optional[torch.Tensor]:
    """Domain length l the periodic variants wrap at (pit.py:191-192, 248-250)."""
    if variant == "periodic1d":
        step = torch.abs(mesh_in[1, 0] - mesh_in[0, 0])
        return step * mesh_in.shape[0]
    if variant == "periodic2d":
        res = int(mesh_in.shape[0] ** 0.5)
        step = (torch.max(mesh_in[:, 0]) - torch.min(mesh_in[:, 0])) / (res - 1)
        return step * res
    return None


def sqdist(mesh_out: torch.Tensor, mesh_in: torch.Tensor, variant: str = "euclid") -> torch.Tensor:
    """Pairwise squared distance ([B,]N,M), fp32, reference rounding order.

    euclid: pit.py:47/134; periodic1d: pit.py:193-195; periodic2d: pit.py:251-253.
    """
    delta = mesh_out.unsqueeze(-2) - mesh_in.unsqueeze(-3)
    if variant == "euclid":
        return torch.sum(delta ** 2, dim=-1)
    wrap = period_of(mesh_in, variant)
    delta = torch.abs(delta)
    delta = torch.minimum(delta, wrap - delta)
    if variant == "periodic1d":
        return delta[..., 0] ** 2
    return torch.sum(delta ** 2, dim=-1)


def dense_attention(mesh_out, mesh_in, scale, locality: float, variant: str = "euclid") -> torch.Tensor:
    """Attention weights ([B,]H,N,M) exactly as the reference materialises them.

    ``scale`` is the already mapped per-head scale s_h with shape (H,1,1).
    """
    d2 = sqdist(mesh_out, mesh_in, variant)
    if mesh_out.dim() == 3:  # per-sample meshes: add the head axis (pit.py:48)
        d2 = d2.unsqueeze(1)
    z = d2 * scale
    cut = torch.quantile(z, locality, dim=-1, keepdim=True)
    z = torch.where(z <= cut, z, torch.tensor(FLT_MAX))
    return torch.softmax(-z, dim=-1)


def dense_contract(att: torch.Tensor, values: torch.Tensor) -> torch.Tensor:
    """(…H,N,M) x (B,M,D) -> (B,N,H*D), head-major then feature (pit.py:54-57, 141-144)."""
    eq = "bhnj,bjd->bnhd" if att.dim() == 4 else "hnj,bjd->bnhd"
    out = torch.einsum(eq, att, values)
    return out.reshape(values.shape[0], out.shape[1], -1)


def dense_posatt(mesh_out, mesh_in, values, lmda, locality: float, variant: str = "euclid",
                 self_concat: bool = False) -> torch.Tensor:
    """Full stage: cross (pit.py:63-71 and variants) or self with concat (pit.py:37-44)."""
    out = dense_contract(dense_attention(mesh_out, mesh_in, head_scale(lmda), locality, variant), values)
    return torch.cat((values, out), dim=-1) if self_concat else out


# --------------------------------------------------------------------------------------
# Order-statistic restatement (what the kernels implement)
# --------------------------------------------------------------------------------------

def quantile_ranks(locality: float, m: int) -> Tuple[int, int, float]:
    """k_lo, k_hi, w of torch.quantile's linear interpolation, fp32 rank arithmetic."""
    rank = torch.tensor(locality, dtype=torch.float32) * (m - 1)
    lo = torch.floor(rank)
    return int(lo.item()), int(torch.ceil(rank).item()), float((rank - lo).item())


def exact_rowstat(d2: torch.Tensor, k_lo: int, k_hi: int):
    """v_min, v_lo, v_hi: smallest, k_lo-th and k_hi-th smallest d2 of every row (0-based)."""
    ordered, _ = torch.sort(d2, dim=-1)
    return ordered[..., 0], ordered[..., k_lo], ordered[..., k_hi]


def exact_threshold(v_lo, v_hi, scale, w: float) -> torch.Tensor:
    """Per-head cut T = lerp(fl(v_lo*s), fl(v_hi*s), w); shapes ([B,]N) x (H,1,1) -> ([B,]H,N)."""
    if v_lo.dim() == 2:
        v_lo, v_hi = v_lo.unsqueeze(1), v_hi.unsqueeze(1)
    s = scale.reshape(-1, 1)
    return torch.lerp(v_lo * s, v_hi * s, torch.tensor(w, dtype=torch.float32))


def exact_weights(mesh_out, mesh_in, scale, locality: float, variant: str = "euclid"):
    """Unnormalised weights P ([B,]H,N,M), row sums and the kept mask, never calling quantile."""
    d2 = sqdist(mesh_out, mesh_in, variant)
    m = d2.shape[-1]
    batched = d2.dim() == 3
    v_min = d2.min(dim=-1).values
    d2h = d2.unsqueeze(1) if batched else d2.unsqueeze(0)
    z = d2h * scale
    top = (v_min.unsqueeze(1) if batched else v_min.unsqueeze(0)) * scale.reshape(-1, 1)
    if locality < 1.0:
        k_lo, k_hi, w = quantile_ranks(locality, m)
        _, v_lo, v_hi = exact_rowstat(d2, k_lo, k_hi)
        keep = z <= exact_threshold(v_lo, v_hi, scale, w).unsqueeze(-1)
    else:
        keep = torch.ones_like(z, dtype=torch.bool)
    p = torch.where(keep, torch.exp(top.unsqueeze(-1) - z), torch.zeros(()))
    return p, p.sum(-1), keep


def exact_posatt(mesh_out, mesh_in, values, lmda, locality: float, variant: str = "euclid",
                 self_concat: bool = False) -> torch.Tensor:
    p, rowsum, _ = exact_weights(mesh_out, mesh_in, head_scale(lmda), locality, variant)
    out = dense_contract(p / rowsum.unsqueeze(-1), values)
    return torch.cat((values, out), dim=-1) if self_concat else out
