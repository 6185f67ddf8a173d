"""Whole-model CPU oracle for PiT -- TEST INFRASTRUCTURE ONLY (see posatt_oracle.py header).

Functional restatement of ``pit.encoder / processor / decoder`` (pit.py:108-127) and of the
``forward`` methods the experiment scripts wrap around them, driven by a plain
``state_dict`` whose keys are the reference's (``down.lmda``, ``en_layer.mlp1.weight`` ...).
Every position-attention stage goes through the dense formulation of
``posatt_oracle`` (quantile + softmax + einsum on materialised N x M tensors), i.e. the
reference's own CPU cost profile -- which is why ``bench.py`` may time it as ``cpu_baseline``.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from . import posatt_oracle as po

Params = Dict[str, torch.Tensor]


def two_layer(p: Params, name: str, x: torch.Tensor) -> torch.Tensor:
    """kaiming_mlp.forward (pit.py:21-26): Linear -> exact GELU -> Linear."""
    x = F.linear(x, p[f"{name}.mlp1.weight"], p[f"{name}.mlp1.bias"])
    return F.linear(F.gelu(x), p[f"{name}.mlp2.weight"], p[f"{name}.mlp2.bias"])


def encode(p: Params, variant, mesh_in, func_in, mesh_ltt, en_loc):
    """pit.encoder (pit.py:108-112)."""
    z = po.dense_posatt(mesh_ltt, mesh_in, func_in, p["down.lmda"], en_loc, variant)
    return F.gelu(two_layer(p, "en_layer", z))


def process(p: Params, variant, func_ltt, mesh_ltt, n_blocks):
    """pit.processor (pit.py:114-122): self stage with concat, locality fixed at 1.0 (pit.py:102)."""
    for i in range(n_blocks):
        z = po.dense_posatt(mesh_ltt, mesh_ltt, func_ltt, p[f"conv.{i}.lmda"], 1.0, variant, self_concat=True)
        func_ltt = F.gelu(two_layer(p, f"mlp.{i}", z))
    return func_ltt


def decode(p: Params, variant, mesh_ltt, func_ltt, mesh_out, de_loc):
    """pit.decoder (pit.py:124-127)."""
    z = po.dense_posatt(mesh_out, mesh_ltt, func_ltt, p["up.lmda"], de_loc, variant)
    return two_layer(p, "de", z)


def n_blocks_of(p: Params) -> int:
    return 1 + max(int(k.split(".")[1]) for k in p if k.startswith("conv."))


def _instance_norm(x: torch.Tensor) -> torch.Tensor:
    """nn.InstanceNorm1d(hid) applied to (B, L, hid) through the permutes of train_vorticity.py:56, 59."""
    return F.instance_norm(x.permute(0, 2, 1)).permute(0, 2, 1)


def forward_shared_mesh(p: Params, variant, mesh_in, func_in, mesh_ltt, mesh_out, en_loc, de_loc, instance_norm=False):
    """Burgers / Sod / Darcy style forward (train_burgers.py:40-49, train_darcy.py:46-59):
    the batch shares one mesh; coordinates are prepended to the input features.  `instance_norm`: the vorticity script's
    normalisation of the latent features after the encoder and after the processor (train_vorticity.py:44-62)."""
    sd = mesh_ltt.shape[-1]
    lead = mesh_out.shape[:-1]
    mesh_in, mesh_out = mesh_in.reshape(-1, sd), mesh_out.reshape(-1, sd)
    func_in = func_in.reshape(func_in.shape[0], mesh_in.shape[0], -1)
    feats = torch.cat((mesh_in.unsqueeze(0).expand(func_in.shape[0], -1, -1), func_in), -1)
    norm = _instance_norm if instance_norm else (lambda x: x)
    h = norm(encode(p, variant, mesh_in, feats, mesh_ltt, en_loc))
    h = norm(process(p, variant, h, mesh_ltt, n_blocks_of(p)))
    out = decode(p, variant, mesh_ltt, h, mesh_out, de_loc)
    return out.reshape(func_in.shape[0], *lead, -1)


def forward_point_cloud(p: Params, mesh_in, func_in, mesh_ltt, mesh_out, en_loc, de_loc):
    """Elasticity / NACA style forward (train_elasticity.py:41-54, train_naca.py:47-61):
    every sample carries its own meshes (B,L,sd); features are used as given."""
    h = encode(p, "euclid", mesh_in, func_in, mesh_ltt, en_loc)
    h = process(p, "euclid", h, mesh_ltt, n_blocks_of(p))
    return decode(p, "euclid", mesh_ltt, h, mesh_out, de_loc)


def rel_lp_loss(true: torch.Tensor, pred: torch.Tensor, out_dim: int, p: int) -> torch.Tensor:
    """RelLpNorm (utils.py:80-98): per-sample relative L_p error, mean over variables, SUM over batch."""
    t = true.reshape(true.shape[0], -1, out_dim)
    q = pred.reshape(pred.shape[0], -1, out_dim)
    ratio = torch.norm(t - q, p=p, dim=1) / torch.norm(t, p=p, dim=1)
    return ratio.mean(-1).sum()


# --------------------------------------------------------------------------------------
# Workload-level helpers (driven by a `workload_specs.Spec`-like object; nothing here touches the CUDA library)
# --------------------------------------------------------------------------------------

def init_params(spec, seed: int = 0) -> Params:
    """Random parameters with the reference's state_dict keys and shapes for a spec (Kaiming-normal Linear weights, lmda ~ U[0,1)
    as pit.py:35; NOT the reference's RNG stream -- used where only shapes and magnitudes matter, e.g. CPU timing)."""
    g = torch.Generator().manual_seed(seed)
    sd, in_dim, out_dim, hid, heads, blocks, _, _ = spec.ctor
    feat = in_dim if spec.family == "batched" else in_dim + sd      # shared-mesh scripts prepend the coordinates
    p: Params = {}

    def linear(name, fan_in, fan_out):
        p[f"{name}.weight"] = torch.randn(fan_out, fan_in, generator=g) * (2.0 / fan_in) ** 0.5
        p[f"{name}.bias"] = (torch.rand(fan_out, generator=g) * 2 - 1) / fan_in ** 0.5

    def mlp(name, a, b, c):
        linear(f"{name}.mlp1", a, b)
        linear(f"{name}.mlp2", b, c)

    p["down.lmda"] = torch.rand(heads, 1, 1, generator=g)
    mlp("en_layer", spec.en_in if spec.en_in else heads * feat, hid, hid)
    for i in range(blocks):
        p[f"conv.{i}.lmda"] = torch.rand(heads, 1, 1, generator=g)
        mlp(f"mlp.{i}", (1 + heads) * hid, hid, hid)
    p["up.lmda"] = torch.rand(heads, 1, 1, generator=g)
    mlp("de", heads * hid, hid, out_dim)
    return p


def forward_spec(p: Params, spec, inputs):
    """One model application for a spec: the `forward` its script defines."""
    _, _, _, _, _, _, en_loc, de_loc = spec.ctor
    if spec.family == "batched":
        mesh_in, func_in, mesh_out = inputs
        if spec.extra.get("latent") == "strided":        # train_naca.py:62-65
            b = mesh_out.shape[0]
            lead = mesh_out.shape[:-1]
            ltt = mesh_out[:, ::spec.extra["x_down"], ::spec.extra["y_down"], :].reshape(b, -1, 2)
            return forward_point_cloud(p, mesh_in, func_in, ltt, mesh_out.reshape(b, -1, 2), en_loc, de_loc).reshape(*lead, -1)
        return forward_point_cloud(p, mesh_in, func_in, mesh_out, mesh_out, en_loc, de_loc)
    mesh, ltt = spec.mesh, spec.mesh_ltt.reshape(-1, spec.mesh_ltt.shape[-1])
    out = forward_shared_mesh(p, spec.variant, mesh, inputs[0], ltt, mesh, en_loc, de_loc, bool(spec.extra.get("instance_norm")))
    return out + inputs[0] if spec.extra.get("residual") else out     # train_cylinder.py:52


def step_loss(p: Params, spec, inputs, target) -> torch.Tensor:
    """Training loss of one step, with the autoregressive rollout of train_vorticity.py:122-126 where the spec asks for it."""
    out_dim, ord_ = spec.loss
    if spec.rollout == 1:
        return rel_lp_loss(target, forward_spec(p, spec, inputs), out_dim, ord_)
    x, loss = inputs[0], 0.0
    for t in range(spec.rollout):
        out = forward_spec(p, spec, (x,))
        loss = loss + rel_lp_loss(out, target[..., t:t + 1], out_dim, ord_)
        x = torch.cat((x[..., 1:], out), -1)
    return loss
