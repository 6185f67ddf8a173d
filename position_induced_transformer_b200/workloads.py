"""The five BASELINE.json workloads as (model, synthetic batch) factories.

Each model class adds the ``forward`` that the corresponding reference script wraps around
``encoder / processor / decoder`` (train_burgers.py:40-49, train_sod.py, train_darcy.py:46-59,
train_elasticity.py:41-54, train_naca.py:47-65); each ``make_*`` function returns the model with
the script's hyper-parameters and a deterministic synthetic batch of the dataset's shape
(the datasets themselves are not distributable: SURVEY.md section 2 #16).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Tuple

import numpy as np
import torch

from . import pit as P
from .utils import RelLpNorm


class SharedMeshPiT:
    """Mixin: the batch shares one mesh, coordinates are prepended to the features (Burgers/Sod/Darcy)."""

    def forward(self, mesh_in, func_in, mesh_out):
        lead = mesh_out.shape[:-1]
        mesh_in = mesh_in.reshape(-1, self.space_dim)
        mesh_out = mesh_out.reshape(-1, self.space_dim)
        func_in = func_in.reshape(func_in.shape[0], -1, self.in_dim)
        feats = torch.cat((mesh_in.unsqueeze(0).expand(func_in.shape[0], -1, -1), func_in), dim=-1)
        latent = self.encoder(mesh_in, feats, self.mesh_ltt)
        latent = self.processor(latent, self.mesh_ltt)
        return self.decoder(self.mesh_ltt, latent, mesh_out).reshape(func_in.shape[0], *lead, self.out_dim)


class BurgersPiT(SharedMeshPiT, P.pit_periodic1d):
    pass


class SodPiT(SharedMeshPiT, P.pit_fixed):
    pass


class DarcyPiT(SharedMeshPiT, P.pit_fixed):
    pass


class VorticityPiT(SharedMeshPiT, P.pit_periodic2d):
    pass


class ElasticityPiT(P.pit):
    """Per-sample point clouds; the latent mesh is the query mesh itself (train_elasticity.py:39-54)."""

    def __init__(self, *args, **kw):
        super().__init__(*args, **kw)
        self.en_layer = P.kaiming_mlp(self.n_head * self.in_dim, self.hid_dim, self.hid_dim)

    def forward(self, mesh_in, func_in, mesh_out):
        mesh_ltt = mesh_out
        latent = self.encoder(mesh_in, func_in, mesh_ltt)
        latent = self.processor(latent, mesh_ltt)
        return self.decoder(mesh_ltt, latent, mesh_out)


class NacaPiT(P.pit):
    """Airfoil: inputs live on the 1-D boundary polyline, the latent mesh is a strided sub-grid of the
    structured query mesh (train_naca.py:26-65)."""

    def __init__(self, *args, x_downsample=4, y_downsample=4, **kw):
        super().__init__(*args, **kw)
        self.x_down, self.y_down = x_downsample, y_downsample
        self.en_layer = P.kaiming_mlp(self.n_head * self.in_dim, self.hid_dim, self.hid_dim)

    def forward(self, mesh_in, func_in, mesh_out):
        lead = mesh_out.shape[:-1]
        b = mesh_out.shape[0]
        mesh_ltt = mesh_out[:, ::self.x_down, ::self.y_down, :].reshape(b, -1, self.space_dim)
        mesh_out = mesh_out.reshape(b, -1, self.space_dim)
        latent = self.encoder(mesh_in, func_in, mesh_ltt)
        latent = self.processor(latent, mesh_ltt)
        return self.decoder(mesh_ltt, latent, mesh_out).reshape(*lead, self.out_dim)


@dataclass
class Workload:
    name: str
    model: torch.nn.Module
    batch_size: int
    make_batch: Callable[[torch.Generator, int], Tuple[tuple, torch.Tensor]]  # (generator, batch) -> (model inputs on CPU, target)
    loss: Callable
    meshes: tuple = ()          # tensors that stay resident on the device across steps (not per-step input)
    note: str = ""

    def to(self, device):
        self.model.to(device)
        if getattr(self.model, "mesh_ltt", None) is not None:
            self.model.mesh_ltt = self.model.mesh_ltt.to(device)
        self.meshes = tuple(m.to(device) for m in self.meshes)
        return self


def grid_points(n: int, lo: float = 0.0, hi: float = 1.0) -> torch.Tensor:
    """(n*n, 2) fp32 grid built like train_darcy.py:83-88 (float64 linspace, meshgrid, cast)."""
    ax = np.linspace(lo, hi, n)
    return torch.tensor(np.vstack([g.ravel() for g in np.meshgrid(ax, ax)]).T, dtype=torch.float)


def make_burgers(batch: int = 8) -> Workload:
    torch.manual_seed(0)
    mesh = torch.linspace(0, 1, 1025)[:-1].reshape(-1, 1)
    ltt = torch.linspace(0, 1, 257)[:-1].reshape(-1, 1)
    model = BurgersPiT(1, 1, 1, 64, 2, 5, ltt, 0.02, 0.02)

    def batch_fn(gen, b):
        return (torch.randn(b, 1024, 1, generator=gen),), torch.randn(b, 1024, 1, generator=gen)

    return Workload("burgers_1024", model, batch, batch_fn, RelLpNorm(1, 1), (mesh,), "train_burgers.py:51-80")


def make_sod(batch: int = 8) -> Workload:
    torch.manual_seed(0)
    mesh = torch.linspace(-5, 5, 2049)[:-1].reshape(-1, 1)
    ltt = torch.linspace(-5, 5, 257)[:-1].reshape(-1, 1)
    model = SodPiT(1, 3, 3, 32, 1, 2, ltt, 0.02, 0.02)

    def batch_fn(gen, b):
        return (torch.rand(b, 2048, 3, generator=gen) * 0.9 + 0.1,), torch.rand(b, 2048, 3, generator=gen) * 0.9 + 0.1

    return Workload("sod_2048", model, batch, batch_fn, RelLpNorm(3, 2), (mesh,), "train_sod.py:55-76")


def make_darcy(side: int = 421, batch: int = 8) -> Workload:
    torch.manual_seed(0)
    mesh = grid_points(side).reshape(side, side, 2)
    ltt = grid_points(16).reshape(16, 16, 2)
    model = DarcyPiT(2, 1, 1, 64, 2, 4, ltt, 0.02, 0.02)

    def batch_fn(gen, b):
        # piecewise-constant coefficient field {3, 12} from a blurred Gaussian field, then standardised
        # (the real a(x) is a thresholded GRF; train_darcy.py:75-79 normalises it pixel-wise)
        field = torch.randn(b, 1, side, side, generator=gen)
        k = 9
        field = torch.nn.functional.avg_pool2d(field, k, stride=1, padding=k // 2)
        coeff = torch.where(field > 0, 12.0, 3.0).reshape(b, side, side, 1)
        coeff = (coeff - 7.5) / 4.5
        target = torch.rand(b, side, side, 1, generator=gen) * 0.013 + 1e-4
        return (coeff,), target

    return Workload(f"darcy_{side}x{side}", model, batch, batch_fn, RelLpNorm(1, 2), (mesh,), "train_darcy.py:62-111")


def make_elasticity(batch: int = 10, points: int = 972) -> Workload:
    torch.manual_seed(0)
    model = ElasticityPiT(2, 44, 1, 256, 2, 4, None, 0.02, 0.02)

    def batch_fn(gen, b):
        # unit-cell point cloud with a central void of random radius, 42 global shape codes broadcast to the points
        ang = torch.rand(b, points, generator=gen) * 2 * np.pi
        hole = 0.2 + 0.2 * torch.rand(b, 1, generator=gen)
        rad = hole + (0.7 - hole) * torch.sqrt(torch.rand(b, points, generator=gen))
        xy = 0.5 + torch.stack((rad * torch.cos(ang), rad * torch.sin(ang)), -1).clamp(-0.5, 0.5)
        codes = torch.rand(b, 1, 42, generator=gen).expand(b, points, 42)
        return (xy, torch.cat((xy, codes), -1), xy), torch.rand(b, points, 1, generator=gen) + 0.5

    return Workload(f"elasticity_{points}", model, batch, batch_fn, RelLpNorm(1, 2), (), "train_elasticity.py:56-96")


def make_naca(batch: int = 20) -> Workload:
    torch.manual_seed(0)
    model = NacaPiT(2, 2, 4, 128, 1, 4, None, 0.02, 0.02)

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

    return Workload("naca_221x51", model, batch, batch_fn, RelLpNorm(4, 2), (), "train_naca.py:68-110")


WORKLOADS = {
    "burgers": make_burgers,
    "sod": make_sod,
    "darcy421": lambda batch=8: make_darcy(421, batch),
    "darcy43": lambda batch=8: make_darcy(43, batch),
    "elasticity": make_elasticity,
    "naca": make_naca,
}


def run_model(w: Workload, inputs):
    """Call the workload's model the way its script does."""
    if w.meshes:                       # shared mesh: model(mesh, x, mesh)
        return w.model(w.meshes[0], inputs[0], w.meshes[0])
    return w.model(*inputs)
