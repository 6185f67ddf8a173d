"""End-to-end parity at BASELINE.json's full sizes: every workload's model (random init, synthetic batch) through the
CUDA path against the CPU oracle -- output, loss and every parameter gradient."""
import pytest
import torch

from conftest import rel_linf
from oracle import pit_oracle

pytestmark = pytest.mark.gpu

CASES = [("burgers", 2), ("sod", 2), ("darcy43", 2), ("darcy421", 1), ("darcy421", 8), ("elasticity", 2), ("naca", 2)]  # darcy421 x 8 = the bench configuration


def _oracle_forward(w, params, ins):
    name = type(w.model).__name__
    if w.meshes:
        variant = {"BurgersPiT": "periodic1d", "VorticityPiT": "periodic2d"}.get(name, "euclid")
        mesh = w.meshes[0].cpu()
        return pit_oracle.forward_shared_mesh(params, variant, mesh, ins[0], w.model.mesh_ltt.cpu(), mesh, w.model.en_local, w.model.de_local)
    mesh_in, func_in, mesh_out = ins
    if name == "NacaPiT":
        b = mesh_out.shape[0]
        lead = mesh_out.shape[:-1]
        ltt = mesh_out[:, ::w.model.x_down, ::w.model.y_down, :].reshape(b, -1, 2)
        out = pit_oracle.forward_point_cloud(params, mesh_in, func_in, ltt, mesh_out.reshape(b, -1, 2), w.model.en_local, w.model.de_local)
        return out.reshape(*lead, -1)
    return pit_oracle.forward_point_cloud(params, mesh_in, func_in, mesh_out, mesh_out, w.model.en_local, w.model.de_local)


@pytest.mark.parametrize("name,batch", CASES)
def test_workload_matches_oracle_at_full_size(name, batch, cuda_device, host_scale_map):
    from position_induced_transformer_b200 import workloads
    w = workloads.WORKLOADS[name](batch)
    gen = torch.Generator().manual_seed(17)
    ins, target = w.make_batch(gen, batch)
    params = {k: v.detach().clone().requires_grad_(True) for k, v in w.model.state_dict().items()}
    want = _oracle_forward(w, params, ins)
    loss_cpu = pit_oracle.rel_lp_loss(target, want, w.model.out_dim, w.loss._ord)
    loss_cpu.backward()

    w.to(cuda_device)
    prev = torch.get_float32_matmul_precision()
    torch.set_float32_matmul_precision("highest")
    try:
        got = workloads.run_model(w, tuple(x.to(cuda_device) for x in ins))
        loss = w.loss(target.to(cuda_device), got)
        loss.backward()
    finally:
        torch.set_float32_matmul_precision(prev)
    assert got.shape == want.shape
    assert rel_linf(got.detach().cpu(), want.detach()) <= 5e-5
    assert abs(float(loss.detach()) - float(loss_cpu.detach())) <= 5e-5 * abs(float(loss_cpu.detach()))
    for k, p in w.model.named_parameters():
        # d(lmda) is a sum of cancelling per-row terms (|result| << sum |terms|), so fp32 summation order (of the upstream
        # MLP gradients too) and the 3xTF32 products show up at the 1e-3..1e-2 level of the (tiny) result -- NACA's encoder
        # gradient is 9e-8, measured against the 1e-6 floor; everything else agrees to 1e-3 of its max-norm
        tol = 1e-2 if k.endswith("lmda") else 1e-3
        assert rel_linf(p.grad.cpu(), params[k].grad, floor=1e-6) <= tol, k
