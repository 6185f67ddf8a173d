"""dX: gradients of position-attention with respect to the mesh coordinates, against the reference's dense algorithm under
autograd (oracle/posatt_oracle.py restates pit.py:46-57 with ordinary differentiable torch ops)."""
import numpy as np
import pytest
import torch

from conftest import rel_linf
from oracle import posatt_oracle as po

pytestmark = pytest.mark.gpu


def _periodic_grid(n):
    ax = np.linspace(0, 1, n + 1)[:-1]
    return torch.tensor(np.vstack([m.ravel() for m in np.meshgrid(ax, ax)]).T, dtype=torch.float)


def _case(kind, g):
    if kind == "fixed_cross_local":
        return "euclid", torch.rand(70, 2, generator=g), torch.rand(90, 2, generator=g), 0.1, False
    if kind == "fixed_self_global":
        m = torch.rand(48, 2, generator=g)
        return "euclid", m, m, 1.0, True
    if kind == "fixed_1d":
        return "euclid", torch.rand(40, 1, generator=g) * 10 - 5, torch.rand(64, 1, generator=g) * 10 - 5, 0.2, False
    if kind == "batched_cross_local":
        return "euclid", torch.rand(3, 50, 2, generator=g), torch.rand(3, 60, 2, generator=g), 0.15, False
    if kind == "periodic1d":
        jitter = lambda n: (torch.linspace(0, 1, n + 1)[:-1] + 0.2 / n * torch.rand(n, generator=g)).reshape(-1, 1)
        return "periodic1d", jitter(32), jitter(64), 0.2, False
    if kind == "periodic2d":
        return "periodic2d", _periodic_grid(6) + 0.01 * torch.rand(36, 2, generator=g), _periodic_grid(8) + 0.01 * torch.rand(64, 2, generator=g), 0.3, False
    raise KeyError(kind)


@pytest.mark.parametrize("kind", ["fixed_cross_local", "fixed_self_global", "fixed_1d", "batched_cross_local", "periodic1d", "periodic2d"])
def test_coordinate_gradients_match_reference_autograd(kind, cuda_device, host_scale_map):
    from position_induced_transformer_b200.pit import head_scale
    from position_induced_transformer_b200.posatt import position_attention
    g = torch.Generator().manual_seed(len(kind))
    variant, mesh_out, mesh_in, q, self_stage = _case(kind, g)
    B, H, D = (mesh_in.shape[0] if mesh_in.dim() == 3 else 2), 2, 5
    values = torch.randn(B, mesh_in.shape[-2], D, generator=g)
    lmda = torch.rand(H, 1, 1, generator=g) * 2 - 1
    # reference algorithm on the CPU, meshes as leaves
    mo_c = mesh_out.clone().requires_grad_(True)
    mi_c = mo_c if self_stage else mesh_in.clone().requires_grad_(True)
    want = po.dense_posatt(mo_c, mi_c, values, lmda, q, variant, self_concat=self_stage)
    up = torch.randn(want.shape, generator=g)
    want.backward(up)
    # fused op
    mo_g = mesh_out.to(cuda_device).requires_grad_(True)
    mi_g = mo_g if self_stage else mesh_in.to(cuda_device).requires_grad_(True)
    v_g = values.to(cuda_device).requires_grad_(True)
    got = position_attention(mo_g, mi_g, v_g, head_scale(lmda.to(cuda_device)), q, variant=variant, self_concat=self_stage)
    got.backward(up.to(cuda_device))
    assert rel_linf(got.detach().cpu(), want.detach()) <= 1e-5
    assert rel_linf(mo_g.grad.cpu(), mo_c.grad) <= 1e-4
    if not self_stage:
        assert rel_linf(mi_g.grad.cpu(), mi_c.grad) <= 1e-4


def test_learnable_latent_mesh_trains_through_the_model(cuda_device):
    """A latent mesh that requires grad routes pit.processor / pit.decoder around the fused kernels and receives a gradient."""
    import position_induced_transformer_b200.pit as pit_mod
    g = torch.Generator().manual_seed(0)
    ltt = torch.rand(32, 2, generator=g).to(cuda_device).requires_grad_(True)
    model = pit_mod.pit_fixed(2, 3, 1, 32, 2, 1, ltt, 0.2, 0.2).to(cuda_device)
    mesh = torch.rand(80, 2, generator=g).to(cuda_device)
    feats = torch.randn(2, 80, 5, generator=g).to(cuda_device)      # in_dim 3 + the two coordinates (train_darcy.py:55)
    out = model.decoder(model.mesh_ltt, model.processor(model.encoder(mesh, feats, model.mesh_ltt), model.mesh_ltt), mesh)
    out.square().sum().backward()
    assert ltt.grad is not None and bool(torch.isfinite(ltt.grad).all()) and float(ltt.grad.abs().max()) > 0
