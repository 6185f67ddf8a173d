"""Synthetic stand-ins for the reference's datasets: shapes, hyper-parameters and batch generators per workload.

Pure torch / numpy -- this module neither imports the PiT modules nor loads libpit_posatt.so, so the CPU reference arm of
bench.py and the oracle-side tests can build the same inputs without touching the product path.  Each spec carries the
literals of the corresponding reference script (the datasets themselves are not distributable: SURVEY.md section 2 #16):

    burgers      train_burgers.py:51-80      pit_periodic1d, 1024 -> 256 -> 1024
    sod          train_sod.py:55-76          pit_fixed (1-D), 2048 -> 256 -> 2048, 3 variables
    darcy421/43  train_darcy.py:62-111       pit_fixed (2-D), n x n -> 16 x 16 -> n x n
    elasticity   train_elasticity.py:56-96   pit (per-sample clouds of 972 points)
    naca         train_naca.py:68-110        pit (120 boundary points -> 56 x 13 latent -> 221 x 51 grid)
    vorticity    train_vorticity.py:76-126   pit_periodic2d, 64 x 64 -> 16 x 16 -> 64 x 64, 20-step unrolled rollout
    cylinder     train_cylinder.py:54-121    pit_fixed on an unstructured 4390 / 896 point mesh, batch 200
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Optional, Tuple

import numpy as np
import torch


@dataclass
class Spec:
    name: str
    family: str                    # "fixed" | "periodic1d" | "periodic2d" | "batched"
    ctor: Tuple                    # (space_dim, in_dim, out_dim, hid_dim, n_head, n_blocks, en_loc, de_loc)
    batch: int                     # the script's batch size
    loss: Tuple[int, int]          # RelLpNorm(out_dim, p)
    make_batch: Callable           # (generator, batch) -> (inputs tuple on CPU, target)
    mesh: Optional[torch.Tensor] = None       # shared query / input mesh (fixed and periodic families)
    mesh_ltt: Optional[torch.Tensor] = None   # shared latent mesh
    en_in: Optional[int] = None    # input width of a script-specific en_layer (train_elasticity.py:39, train_naca.py:45)
    rollout: int = 1               # autoregressive steps per training step (train_vorticity.py:122-126)
    extra: dict = field(default_factory=dict)
    source: str = ""

    @property
    def variant(self) -> str:
        return {"periodic1d": "periodic1d", "periodic2d": "periodic2d"}.get(self.family, "euclid")


def grid_points(n: int, lo: float = 0.0, hi: float = 1.0, endpoint: bool = True) -> torch.Tensor:
    """(n*n, 2) fp32 grid built like train_darcy.py:83-88 (float64 linspace, meshgrid, cast)."""
    ax = np.linspace(lo, hi, n) if endpoint else np.linspace(lo, hi, n + 1)[:-1]
    return torch.tensor(np.vstack([g.ravel() for g in np.meshgrid(ax, ax)]).T, dtype=torch.float)


def burgers() -> Spec:
    mesh = torch.linspace(0, 1, 1025)[:-1].reshape(-1, 1)
    ltt = torch.linspace(0, 1, 257)[:-1].reshape(-1, 1)

    def batch_fn(gen, b):
        return (torch.randn(b, 1024, 1, generator=gen),), torch.randn(b, 1024, 1, generator=gen)

    return Spec("burgers_1024", "periodic1d", (1, 1, 1, 64, 2, 5, 0.02, 0.02), 8, (1, 1), batch_fn, mesh, ltt, source="train_burgers.py:51-80")


def sod() -> Spec:
    mesh = torch.linspace(-5, 5, 2049)[:-1].reshape(-1, 1)
    ltt = torch.linspace(-5, 5, 257)[:-1].reshape(-1, 1)

    def batch_fn(gen, b):
        return (torch.rand(b, 2048, 3, generator=gen) * 0.9 + 0.1,), torch.rand(b, 2048, 3, generator=gen) * 0.9 + 0.1

    return Spec("sod_2048", "fixed", (1, 3, 3, 32, 1, 2, 0.02, 0.02), 8, (3, 2), batch_fn, mesh, ltt, source="train_sod.py:55-76")


def darcy(side: int = 421) -> Spec:
    mesh = grid_points(side).reshape(side, side, 2)
    ltt = grid_points(16).reshape(16, 16, 2)

    def batch_fn(gen, b):
        # piecewise-constant coefficient field {3, 12} from a blurred Gaussian field, then standardised
        # (the real a(x) is a thresholded GRF; train_darcy.py:75-79 normalises it pixel-wise)
        fld = torch.randn(b, 1, side, side, generator=gen)
        k = 9
        fld = torch.nn.functional.avg_pool2d(fld, k, stride=1, padding=k // 2)
        coeff = torch.where(fld > 0, 12.0, 3.0).reshape(b, side, side, 1)
        coeff = (coeff - 7.5) / 4.5
        target = torch.rand(b, side, side, 1, generator=gen) * 0.013 + 1e-4
        return (coeff,), target

    return Spec(f"darcy_{side}x{side}", "fixed", (2, 1, 1, 64, 2, 4, 0.02, 0.02), 8, (1, 2), batch_fn, mesh, ltt, source="train_darcy.py:62-111")


def elasticity(points: int = 972) -> Spec:
    def batch_fn(gen, b):
        # unit-cell point cloud with a central void of random radius, 42 global shape codes broadcast to the points
        ang = torch.rand(b, points, generator=gen) * 2 * np.pi
        hole = 0.2 + 0.2 * torch.rand(b, 1, generator=gen)
        rad = hole + (0.7 - hole) * torch.sqrt(torch.rand(b, points, generator=gen))
        xy = 0.5 + torch.stack((rad * torch.cos(ang), rad * torch.sin(ang)), -1).clamp(-0.5, 0.5)
        codes = torch.rand(b, 1, 42, generator=gen).expand(b, points, 42)
        return (xy, torch.cat((xy, codes), -1), xy), torch.rand(b, points, 1, generator=gen) + 0.5

    return Spec(f"elasticity_{points}", "batched", (2, 44, 1, 256, 2, 4, 0.02, 0.02), 10, (1, 2), batch_fn, en_in=2 * 44,
                extra={"latent": "query"}, source="train_elasticity.py:56-96")


def naca() -> Spec:
    def batch_fn(gen, b):
        # NACA 4-digit-like airfoil polyline (120 points) and a 221 x 51 O-grid growing out of it
        t = torch.linspace(0, 2 * np.pi, 121)[:-1]
        thick = 0.08 + 0.1 * torch.rand(b, 1, generator=gen)
        camber = 0.04 * torch.rand(b, 1, generator=gen)
        xs = 0.5 + 0.5 * torch.cos(t).unsqueeze(0).expand(b, -1)
        ys = thick * torch.sin(t).unsqueeze(0) * torch.sqrt(xs.clamp_min(1e-4)) * (1 - xs) * 3 + camber * torch.sin(np.pi * xs)
        foil = torch.stack((xs, ys), -1)
        t2 = torch.linspace(0, 2 * np.pi, 221)
        x2 = 0.5 + 0.5 * torch.cos(t2).unsqueeze(0).expand(b, -1)
        y2 = thick * torch.sin(t2).unsqueeze(0) * torch.sqrt(x2.clamp_min(1e-4)) * (1 - x2) * 3 + camber * torch.sin(np.pi * x2)
        inner = torch.stack((x2, y2), -1)                                        # (b, 221, 2)
        outer = torch.stack((0.5 + 3 * torch.cos(t2), 3 * torch.sin(t2)), -1)     # (221, 2)
        s = (torch.linspace(0, 1, 51) ** 2).reshape(1, 1, 51, 1)
        grid = inner.unsqueeze(2) * (1 - s) + outer.reshape(1, 221, 1, 2) * s     # (b, 221, 51, 2)
        return (foil, foil.clone(), grid), torch.rand(b, 221, 51, 4, generator=gen) + 0.5

    return Spec("naca_221x51", "batched", (2, 2, 4, 128, 1, 4, 0.02, 0.02), 20, (4, 2), batch_fn, en_in=1 * 2,
                extra={"latent": "strided", "x_down": 4, "y_down": 4}, source="train_naca.py:68-110")


def vorticity(steps: int = 20) -> Spec:
    """Navier-Stokes vorticity on the periodic unit square: 10 past frames in, the next frame out, `steps` autoregressive
    applications per training step with back-propagation through the whole rollout (train_vorticity.py:122-126)."""
    s = 64
    mesh = grid_points(s, endpoint=False).reshape(s, s, 2)
    ltt = grid_points(16, endpoint=False).reshape(16, 16, 2)

    def batch_fn(gen, b):
        # smooth periodic fields: a few random Fourier modes per frame
        kx = torch.arange(1, 4).reshape(1, 1, 3, 1, 1) * 2 * np.pi
        xs = torch.linspace(0, 1, s + 1)[:-1]
        gx, gy = torch.meshgrid(xs, xs, indexing="xy")
        amp = torch.randn(b, 10 + steps, 3, 1, 1, generator=gen) / 3
        ph = torch.rand(b, 10 + steps, 3, 1, 1, generator=gen) * 2 * np.pi
        frames = (amp * torch.sin(kx * gx + ph) * torch.cos(kx * gy + 0.5 * ph)).sum(2)    # (b, 10+steps, s, s)
        frames = frames.permute(0, 2, 3, 1).contiguous()
        return (frames[..., :10].contiguous(),), frames[..., 10:].contiguous()

    return Spec(f"vorticity_64x64_T{steps}", "periodic2d", (2, 10, 1, 256, 2, 4, 0.02, 0.02), 20, (1, 2), batch_fn, mesh, ltt, rollout=steps,
                extra={"instance_norm": True}, source="train_vorticity.py:76-126")


def cylinder() -> Spec:
    """Wake behind a cylinder on an unstructured mesh of 4390 vertices, 896 latent vertices, batch 200, residual output
    (train_cylinder.py:40-52, 82-121).  The vertex files are not distributable: a seeded jittered lattice around a
    circular hole stands in for them."""
    g = torch.Generator().manual_seed(4390)

    def cloud(n):
        pts = torch.empty(0, 2)
        while pts.shape[0] < n:
            p = torch.rand(2 * n, 2, generator=g) * torch.tensor([2.2, 0.41])
            keep = ((p - torch.tensor([0.2, 0.2])) ** 2).sum(-1) > 0.05 ** 2
            pts = torch.cat((pts, p[keep]))
        return pts[:n].contiguous()

    mesh, ltt = cloud(4390), cloud(896)

    def batch_fn(gen, b):
        x = torch.randn(b, 4390, 3, generator=gen)
        return (x,), x + 0.1 * torch.randn(b, 4390, 3, generator=gen)

    return Spec("cylinder_4390", "fixed", (2, 3, 3, 256, 1, 4, 0.01, 0.01), 200, (3, 2), batch_fn, mesh, ltt, extra={"residual": True},
                source="train_cylinder.py:54-121")


SPECS = {
    "burgers": burgers,
    "sod": sod,
    "darcy421": lambda: darcy(421),
    "darcy43": lambda: darcy(43),
    "elasticity": elasticity,
    "naca": naca,
    "vorticity": vorticity,
    "cylinder": cylinder,
}

# the five configurations BASELINE.json names, in its order
BASELINE_WORKLOADS = ("burgers", "sod", "darcy421", "elasticity", "naca")
