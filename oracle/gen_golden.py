"""Generate tests/golden/*.npz by running the UNMODIFIED reference (`/root/reference/pit.py`) on CPU.

Run in the build container only (the GPU box has no /root/reference):

    python -m oracle.gen_golden

The reference is loaded under a private module name; nothing from it is copied into the
repo -- only its numerical inputs/outputs are stored.  Each op-level file holds the meshes,
value features, lmda, an upstream gradient, and the reference's forward output, input
gradient, lmda gradient, bit-packed kept mask (att > 0) and attention row checksums.
Each model-level file holds a reference state_dict, inputs, output, loss and gradients.
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

REF = "/root/reference/pit.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_pit", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)          # reseeds torch/numpy RNG and sets matmul precision 'high'
    torch.set_float32_matmul_precision("highest")
    return mod


def grid2d(s, lo=0.0, hi=1.0):
    ax = np.linspace(lo, hi, s)
    pts = np.vstack([g.ravel() for g in np.meshgrid(ax, ax)]).T      # same construction as train_darcy.py:83-88
    return torch.tensor(pts, dtype=torch.float)


def grid2d_periodic(s):
    ax = np.linspace(0, 1, s + 1)[:-1]
    pts = np.vstack([g.ravel() for g in np.meshgrid(ax, ax)]).T
    return torch.tensor(pts, dtype=torch.float)


def op_case(ref, name, kind, cls_name, mesh_out, mesh_in, values, n_head, locality, gen):
    """kind: 'cross' or 'self'.  Runs reference forward/backward and stores everything."""
    layer = getattr(ref, cls_name)(n_head, values.shape[-1], locality)
    with torch.no_grad():
        layer.lmda.copy_(torch.rand(n_head, 1, 1, generator=gen) * 4 - 2)     # lmda in [-2, 2)
    values = values.clone().requires_grad_(True)
    if kind == "self":
        out = layer(mesh_out, values)
        att = layer.dist2att(mesh_out, mesh_out, layer.lmda, layer.locality)
    else:
        out = layer(mesh_out, mesh_in, values)
        att = layer.dist2att(mesh_out, mesh_in, layer.lmda, layer.locality)
    upstream = torch.randn(out.shape, generator=gen)
    out.backward(upstream)
    att = att.detach()
    scale = torch.tan(0.25 * np.pi * (1 - 1e-7) * (1.0 + torch.sin(layer.lmda.detach())))
    np.savez_compressed(
        os.path.join(OUT, f"op_{name}.npz"),
        kind=kind, cls=cls_name, locality=np.float64(locality), n_head=n_head,
        mesh_out=mesh_out.numpy(), mesh_in=mesh_in.numpy(), values=values.detach().numpy(),
        lmda=layer.lmda.detach().numpy(), scale=scale.numpy(), upstream=upstream.numpy(),
        out=out.detach().numpy(), d_values=values.grad.numpy(), d_lmda=layer.lmda.grad.numpy(),
        kept_bits=np.packbits((att > 0).numpy().reshape(-1)), att_shape=np.array(att.shape),
        kept_per_row=(att > 0).sum(-1).numpy().astype(np.int32),
        att_max=att.max(-1).values.numpy(),
    )
    print(f"op_{name}: out {tuple(out.shape)} kept/row {int((att > 0).sum(-1).min())}..{int((att > 0).sum(-1).max())}")


def model_case(ref, name, cls_name, ctor, forward, inputs, target, loss_p, en_hidden=None):
    torch.manual_seed(0)
    model = getattr(ref, cls_name)(**ctor)
    if en_hidden is not None:                      # train_elasticity.py:39 / train_naca.py:45 override
        model.en_layer = ref.kaiming_mlp(*en_hidden)
    gen = torch.Generator().manual_seed(99)
    with torch.no_grad():                          # move lmda away from the U[0,1) init to exercise the scale map
        for k, v in model.named_parameters():
            if k.endswith("lmda"):
                v.copy_(torch.rand(v.shape, generator=gen) * 3 - 1.5)
    out = forward(model, *inputs)
    t = target.reshape(target.shape[0], -1, ctor["out_dim"])
    q = out.reshape(out.shape[0], -1, ctor["out_dim"])
    loss = (torch.norm(t - q, p=loss_p, dim=1) / torch.norm(t, p=loss_p, dim=1)).mean(-1).sum()
    loss.backward()
    blob = {f"param/{k}": v.detach().numpy() for k, v in model.state_dict().items()}
    blob.update({f"grad/{k}": v.grad.numpy() for k, v in model.named_parameters()})
    blob.update({f"input/{i}": x.numpy() for i, x in enumerate(inputs)})
    ctor_np = {f"ctor/{k}": (v.numpy() if torch.is_tensor(v) else np.array(-1 if v is None else v)) for k, v in ctor.items()}
    np.savez_compressed(os.path.join(OUT, f"model_{name}.npz"), cls=cls_name, target=target.numpy(),
                        out=out.detach().numpy(), loss=loss.detach().numpy(), loss_p=loss_p, **blob, **ctor_np)
    print(f"model_{name}: out {tuple(out.shape)} loss {float(loss):.6f}")


def shared_mesh_forward(model, mesh_in, func_in, mesh_out):
    """What train_burgers.py:40-49 / train_darcy.py:46-59 do around encoder/processor/decoder."""
    sd = model.space_dim
    lead = mesh_out.shape[:-1]
    mi, mo = mesh_in.reshape(-1, sd), mesh_out.reshape(-1, sd)
    f = func_in.reshape(func_in.shape[0], -1, model.in_dim)
    f = torch.cat((torch.tile(mi.unsqueeze(0), [f.shape[0], 1, 1]), f), -1)
    h = model.encoder(mi, f, model.mesh_ltt)
    h = model.processor(h, model.mesh_ltt)
    return model.decoder(model.mesh_ltt, h, mo).reshape(f.shape[0], *lead, model.out_dim)


def vorticity_forward(model, mesh_in, func_in, mesh_out):
    """train_vorticity.py:44-62: the shared-mesh wrapper with a parameter-free InstanceNorm1d over the latent points after the
    encoder and after the processor."""
    sd = model.space_dim
    lead = mesh_out.shape[:-1]
    norm = torch.nn.InstanceNorm1d(model.hid_dim)
    mi, mo = mesh_in.reshape(-1, sd), mesh_out.reshape(-1, sd)
    f = func_in.reshape(func_in.shape[0], -1, model.in_dim)
    f = torch.cat((torch.tile(mi.unsqueeze(0), [f.shape[0], 1, 1]), f), -1)
    h = model.encoder(mi, f, model.mesh_ltt)
    h = norm(h.permute(0, 2, 1)).permute(0, 2, 1)
    h = model.processor(h, model.mesh_ltt)
    h = norm(h.permute(0, 2, 1)).permute(0, 2, 1)
    return model.decoder(model.mesh_ltt, h, mo).reshape(f.shape[0], *lead, model.out_dim)


def cloud_forward(model, mesh_in, func_in, mesh_ltt, mesh_out):
    h = model.encoder(mesh_in, func_in, mesh_ltt)
    h = model.processor(h, mesh_ltt)
    return model.decoder(mesh_ltt, h, mesh_out)


def main():
    if not os.path.exists(REF):
        sys.exit("reference not present; golden vectors can only be generated in the build container")
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference()
    g = torch.Generator().manual_seed(20261017)
    rn = lambda *s: torch.randn(*s, generator=g)
    ru = lambda *s: torch.rand(*s, generator=g)

    # ---- op level: fixed-mesh Euclidean (Darcy-43 grids: tie-heavy, non-dyadic) ----
    m43, m16 = grid2d(43), grid2d(16)
    op_case(ref, "fixed_darcy_enc", "cross", "posatt_cross_fixed", m16, m43, rn(2, 1849, 3), 2, 0.02, g)
    op_case(ref, "fixed_darcy_dec", "cross", "posatt_cross_fixed", m43, m16, rn(2, 256, 8), 2, 0.02, g)
    op_case(ref, "fixed_darcy_self", "self", "posatt_fixed", m16, m16, rn(2, 256, 8), 2, 1.0, g)
    # 1-D fixed mesh on [-5,5) (Sod)
    s_in = torch.linspace(-5, 5, 2049)[:-1].reshape(-1, 1)
    s_lt = torch.linspace(-5, 5, 257)[:-1].reshape(-1, 1)
    op_case(ref, "fixed_sod_enc", "cross", "posatt_cross_fixed", s_lt, s_in, rn(2, 2048, 4), 1, 0.02, g)
    op_case(ref, "fixed_sod_dec", "cross", "posatt_cross_fixed", s_in, s_lt, rn(2, 256, 8), 1, 0.02, g)
    # unstructured fixed mesh (cylinder-like): no ties
    cu_in, cu_lt = ru(700, 2), ru(96, 2)
    op_case(ref, "fixed_cloud_enc", "cross", "posatt_cross_fixed", cu_lt, cu_in, rn(3, 700, 5), 1, 0.01, g)
    op_case(ref, "fixed_cloud_self", "self", "posatt_fixed", cu_lt, cu_lt, rn(3, 96, 16), 3, 1.0, g)
    # ---- periodic 1-D (Burgers; dyadic coordinates => exact ties) ----
    b_in = torch.linspace(0, 1, 1025)[:-1].reshape(-1, 1)
    b_lt = torch.linspace(0, 1, 257)[:-1].reshape(-1, 1)
    op_case(ref, "per1d_enc", "cross", "posatt_cross_periodic1d", b_lt, b_in, rn(2, 1024, 2), 2, 0.02, g)
    op_case(ref, "per1d_dec", "cross", "posatt_cross_periodic1d", b_in, b_lt, rn(2, 256, 8), 2, 0.02, g)
    op_case(ref, "per1d_self", "self", "posatt_periodic1d", b_lt, b_lt, rn(2, 256, 8), 2, 1.0, g)
    # ---- periodic 2-D (vorticity) ----
    p32, p8 = grid2d_periodic(32), grid2d_periodic(8)
    op_case(ref, "per2d_enc", "cross", "posatt_cross_periodic2d", p8, p32, rn(2, 1024, 4), 2, 0.05, g)
    op_case(ref, "per2d_dec", "cross", "posatt_cross_periodic2d", p32, p8, rn(2, 64, 8), 2, 0.1, g)
    op_case(ref, "per2d_self", "self", "posatt_periodic2d", p8, p8, rn(2, 64, 8), 2, 1.0, g)
    # ---- per-sample meshes (elasticity / NACA) ----
    c_out, c_in = ru(3, 97, 2), ru(3, 120, 2)
    op_case(ref, "batched_cross", "cross", "posatt_cross", c_out, c_in, rn(3, 120, 5), 2, 0.05, g)
    op_case(ref, "batched_self", "self", "posatt", c_out, c_out, rn(3, 97, 12), 2, 1.0, g)
    op_case(ref, "batched_self_local", "self", "posatt", c_out, c_out, rn(3, 97, 12), 2, 0.1, g)
    # structured per-sample meshes (ties inside every sample)
    st = grid2d(12).unsqueeze(0).repeat(2, 1, 1) * torch.tensor([[[1.0, 1.0]], [[0.5, 2.0]]])
    op_case(ref, "batched_grid_cross", "cross", "posatt_cross", st, st[:, ::3].contiguous(), rn(2, 48, 6), 2, 0.1, g)

    # ---- model level ----
    torch.manual_seed(0)
    model_case(ref, "burgers", "pit_periodic1d",
               dict(space_dim=1, in_dim=1, out_dim=1, hid_dim=32, n_head=2, n_blocks=2, mesh_ltt=b_lt, en_loc=0.02, de_loc=0.02),
               shared_mesh_forward, (b_in, rn(2, 1024, 1), b_in), rn(2, 1024, 1), 1)
    model_case(ref, "sod", "pit_fixed",
               dict(space_dim=1, in_dim=3, out_dim=3, hid_dim=32, n_head=1, n_blocks=2, mesh_ltt=s_lt, en_loc=0.02, de_loc=0.02),
               shared_mesh_forward, (s_in, ru(2, 2048, 3) + 0.1, s_in), ru(2, 2048, 3) + 0.1, 2)
    model_case(ref, "darcy43", "pit_fixed",
               dict(space_dim=2, in_dim=1, out_dim=1, hid_dim=32, n_head=2, n_blocks=2, mesh_ltt=m16.reshape(16, 16, 2), en_loc=0.02, de_loc=0.02),
               shared_mesh_forward, (m43.reshape(43, 43, 2), rn(2, 43, 43, 1), m43.reshape(43, 43, 2)), ru(2, 43, 43, 1) + 0.5, 2)
    cl = ru(2, 150, 2)
    model_case(ref, "elasticity", "pit",
               dict(space_dim=2, in_dim=6, out_dim=1, hid_dim=32, n_head=2, n_blocks=2, mesh_ltt=None, en_loc=0.05, de_loc=0.05),
               cloud_forward, (cl, torch.cat((cl, ru(2, 1, 4).expand(2, 150, 4)), -1), cl.clone(), cl), ru(2, 150, 1) + 0.5, 2,
               en_hidden=(12, 32, 32))
    na_out = ru(2, 40, 9, 2)
    model_case(ref, "naca", "pit",
               dict(space_dim=2, in_dim=2, out_dim=4, hid_dim=32, n_head=1, n_blocks=2, mesh_ltt=None, en_loc=0.1, de_loc=0.05),
               cloud_forward, (ru(2, 30, 2), ru(2, 30, 2), na_out[:, ::4, ::4].reshape(2, -1, 2), na_out.reshape(2, -1, 2)),
               ru(2, 360, 4) + 0.5, 2, en_hidden=(2, 32, 32))
    model_case(ref, "vorticity", "pit_periodic2d",
               dict(space_dim=2, in_dim=1, out_dim=1, hid_dim=16, n_head=2, n_blocks=1, mesh_ltt=p8, en_loc=0.05, de_loc=0.1),
               shared_mesh_forward, (p32, rn(2, 1024, 1), p32), rn(2, 1024, 1), 2)
    # the script's own wrapper (instance norms around the processor), 10 input frames as train_vorticity.py:76
    model_case(ref, "vorticity_norm", "pit_periodic2d",
               dict(space_dim=2, in_dim=10, out_dim=1, hid_dim=16, n_head=2, n_blocks=1, mesh_ltt=p8, en_loc=0.05, de_loc=0.1),
               vorticity_forward, (p32, rn(2, 1024, 10), p32), rn(2, 1024, 1), 2)


if __name__ == "__main__":
    main()
