"""The reference's experiment configurations as (model, synthetic batch) factories.

Each model class adds the ``forward`` that the corresponding reference script wraps around
``encoder / processor / decoder`` (train_burgers.py:40-49, train_sod.py, train_darcy.py:46-59,
train_elasticity.py:41-54, train_naca.py:47-65, train_vorticity.py:44-62, train_cylinder.py:40-52); the
hyper-parameters, meshes and synthetic batch generators come from ``workload_specs`` (pure data, shared with the
CPU reference arm of bench.py, which must not load the CUDA library).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Tuple

import numpy as np
import torch

from . import pit as P
from . import workload_specs as WS
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


class Periodic2dPiT(SharedMeshPiT, P.pit_periodic2d):
    pass


class VorticityPiT(P.pit_periodic2d):
    """Periodic 2-D vorticity model with an instance norm after the encoder and after the processor
    (train_vorticity.py:43, 56-59)."""

    def __init__(self, *args, **kw):
        super().__init__(*args, **kw)
        self.norm = torch.nn.InstanceNorm1d(self.hid_dim)

    def forward(self, mesh_in, func_in, mesh_out):
        lead = mesh_out.shape[:-1]
        mesh_in = mesh_in.reshape(-1, self.space_dim)
        mesh_out = mesh_out.reshape(-1, self.space_dim)
        func_in = func_in.reshape(func_in.shape[0], -1, self.in_dim)
        feats = torch.cat((mesh_in.unsqueeze(0).expand(func_in.shape[0], -1, -1), func_in), dim=-1)
        latent = self.encoder(mesh_in, feats, self.mesh_ltt)
        latent = self.norm(latent.permute(0, 2, 1)).permute(0, 2, 1)
        latent = self.processor(latent, self.mesh_ltt)
        latent = self.norm(latent.permute(0, 2, 1)).permute(0, 2, 1)
        return self.decoder(self.mesh_ltt, latent, mesh_out).reshape(func_in.shape[0], *lead, self.out_dim)


class CylinderPiT(SharedMeshPiT, P.pit_fixed):
    """Unstructured wake mesh; the model predicts the increment (train_cylinder.py:40-52: `return func_out + x`)."""

    def forward(self, mesh_in, func_in, mesh_out):
        return SharedMeshPiT.forward(self, mesh_in, func_in, mesh_out) + func_in


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
    rollout: int = 1            # autoregressive model applications per training step (train_vorticity.py:122-126)
    spec: WS.Spec = None

    def to(self, device):
        self.model.to(device)
        if getattr(self.model, "mesh_ltt", None) is not None:
            self.model.mesh_ltt = self.model.mesh_ltt.to(device)
        self.meshes = tuple(m.to(device) for m in self.meshes)
        return self


_MODEL_CLASS = {"burgers": BurgersPiT, "sod": SodPiT, "darcy421": DarcyPiT, "darcy43": DarcyPiT, "elasticity": ElasticityPiT,
                "naca": NacaPiT, "vorticity": VorticityPiT, "cylinder": CylinderPiT}


def make(name: str, batch: int = 0, spec: WS.Spec = None) -> Workload:
    """Model with the script's hyper-parameters (seeded like pit.py:3) and the synthetic batch generator of `name`
    (`spec` overrides the stock one, e.g. a shorter vorticity rollout)."""
    spec = spec or WS.SPECS[name]()
    torch.manual_seed(0)
    sd, in_dim, out_dim, hid, heads, blocks, en_loc, de_loc = spec.ctor
    kw = {}
    if name == "naca":
        kw = {"x_downsample": spec.extra["x_down"], "y_downsample": spec.extra["y_down"]}
    model = _MODEL_CLASS[name](sd, in_dim, out_dim, hid, heads, blocks, spec.mesh_ltt, en_loc, de_loc, **kw)
    meshes = () if spec.mesh is None else (spec.mesh,)
    return Workload(spec.name, model, batch or spec.batch, spec.make_batch, RelLpNorm(*spec.loss), meshes, spec.source, spec.rollout, spec)


def make_burgers(batch: int = 8) -> Workload:
    return make("burgers", batch)


def make_sod(batch: int = 8) -> Workload:
    return make("sod", batch)


def make_darcy(side: int = 421, batch: int = 8) -> Workload:
    if side in (421, 43):
        return make(f"darcy{side}", batch)
    spec = WS.darcy(side)
    torch.manual_seed(0)
    model = DarcyPiT(*spec.ctor[:6], spec.mesh_ltt, *spec.ctor[6:])
    return Workload(spec.name, model, batch, spec.make_batch, RelLpNorm(*spec.loss), (spec.mesh,), spec.source, 1, spec)


def make_elasticity(batch: int = 10) -> Workload:
    return make("elasticity", batch)


def make_naca(batch: int = 20) -> Workload:
    return make("naca", batch)


WORKLOADS = {name: (lambda batch=0, _n=name: make(_n, batch)) for name in WS.SPECS}


def run_model(w: Workload, inputs):
    """Call the workload's model the way its script does."""
    if w.meshes:                       # shared mesh: model(mesh, x, mesh)
        return w.model(w.meshes[0], inputs[0], w.meshes[0])
    return w.model(*inputs)


def step_loss(w: Workload, inputs, target):
    """Loss of one training step as the script forms it: one model application, or -- vorticity -- an unrolled rollout in
    which every prediction is appended to the input window and the per-step losses are summed (train_vorticity.py:122-126)."""
    if w.rollout == 1:
        return w.loss(target, run_model(w, inputs))
    x, loss = inputs[0], 0.0
    for t in range(w.rollout):
        out = run_model(w, (x,))
        loss = loss + w.loss(out, target[..., t:t + 1])
        x = torch.cat((x[..., 1:], out), dim=-1)
    return loss
