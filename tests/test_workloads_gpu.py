"""End-to-end parity at BASELINE.json's full sizes: every workload's model (random init, synthetic batch) through the
CUDA path against the CPU oracle -- output, loss and every parameter gradient."""
import pytest
import torch

from conftest import rel_linf
from oracle import pit_oracle

pytestmark = pytest.mark.gpu

# darcy421 x 8 = the bench configuration; vorticity with a 2-step rollout (the script unrolls 20) and cylinder at batch 8 (the
# script uses 200) keep the dense CPU oracle within seconds
CASES = [("burgers", 2), ("sod", 2), ("darcy43", 2), ("darcy421", 1), ("darcy421", 8), ("elasticity", 2), ("naca", 2),
         ("vorticity", 2), ("cylinder", 8)]


@pytest.mark.parametrize("name,batch", CASES)
def test_workload_matches_oracle_at_full_size(name, batch, cuda_device, host_scale_map):
    from position_induced_transformer_b200 import workload_specs, workloads
    spec = workload_specs.vorticity(2) if name == "vorticity" else None
    w = workloads.make(name, batch, spec)
    gen = torch.Generator().manual_seed(17)
    ins, target = w.make_batch(gen, batch)
    params = {k: v.detach().clone().requires_grad_(True) for k, v in w.model.state_dict().items()}
    want = pit_oracle.forward_spec(params, w.spec, ins)
    loss_cpu = pit_oracle.step_loss(params, w.spec, ins, target)
    loss_cpu.backward()

    w.to(cuda_device)
    prev = torch.get_float32_matmul_precision()
    torch.set_float32_matmul_precision("highest")
    try:
        dev_ins = tuple(x.to(cuda_device) for x in ins)
        got = workloads.run_model(w, dev_ins)
        loss = workloads.step_loss(w, dev_ins, target.to(cuda_device))
        loss.backward()
    finally:
        torch.set_float32_matmul_precision(prev)
    assert got.shape == want.shape
    errs = {k: rel_linf(p.grad.cpu(), params[k].grad, floor=1e-6) for k, p in w.model.named_parameters()}
    worst_l = max((v, k) for k, v in errs.items() if k.endswith("lmda"))
    worst_o = max((v, k) for k, v in errs.items() if not k.endswith("lmda"))
    print(f"\n{name} B={batch}: out {rel_linf(got.detach().cpu(), want.detach()):.2e} loss "
          f"{abs(float(loss.detach()) - float(loss_cpu.detach())) / abs(float(loss_cpu.detach())):.2e} grads {worst_o[0]:.2e} ({worst_o[1]}) "
          f"lmda grads {worst_l[0]:.2e} ({worst_l[1]})")
    # Measured (B200, this commit): shared-mesh workloads out <= 3e-6, loss <= 2.5e-6, gradients <= 6e-5 of their max-norm;
    # per-sample meshes (elasticity, NACA: six 3xTF32 tcgen05 stages with 728..972-long reductions accumulated in TMEM,
    # which truncates) out 4.3e-5, gradients 8.5e-5 (1.2e-4 on the vorticity rollout).  The bounds below are ~3-4x those figures.
    per_sample = name in ("elasticity", "naca")
    assert rel_linf(got.detach().cpu(), want.detach()) <= (1.5e-4 if per_sample else 1e-5)
    assert abs(float(loss.detach()) - float(loss_cpu.detach())) <= 1e-5 * abs(float(loss_cpu.detach()))
    for k, p in w.model.named_parameters():
        # d(lmda) is a sum of cancelling per-row terms (|result| << sum |terms|), so fp32 summation order (of the upstream
        # MLP gradients too) and the 3xTF32 products show up at the 1e-3..1e-2 level of the (tiny) result -- NACA's encoder
        # gradient is 9e-8, measured against the 1e-6 floor (8.9e-3 there; <= 4.8e-3 elsewhere)
        tol = 3e-2 if k.endswith("lmda") else 5e-4
        assert errs[k] <= tol, k
