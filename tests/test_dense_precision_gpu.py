"""Operand-precision modes of the global (dense, tcgen05) stages against the CPU oracle: fp32 parity (3xTF32), TF32 and BF16
operands under the stated bounds of SURVEY 8c."""
import pytest
import torch

from conftest import rel_linf
from oracle import posatt_oracle as po

pytestmark = pytest.mark.gpu

# forward rel-Linf, value gradient, lmda gradient (each of its max-norm).  The scale gradient always multiplies 3xTF32 (it is a
# difference of large sums), so its error in the reduced modes is what the perturbed forward output feeds back through autograd.
BOUNDS = {"fp32": (1e-5, 1e-4, 1e-4), "tf32": (1e-3, 1e-3, 1e-3), "bf16": (5e-3, 5e-3, 5e-3)}


@pytest.fixture
def dense_precision():
    from position_induced_transformer_b200 import posatt
    yield posatt.set_dense_precision
    posatt.set_dense_precision("fp32")


@pytest.mark.parametrize("batched,B,N,D,H", [(True, 3, 300, 64, 2), (False, 4, 512, 128, 1)])
@pytest.mark.parametrize("mode", ["fp32", "tf32", "bf16"])
def test_dense_stage_precision_modes(batched, B, N, D, H, mode, cuda_device, host_scale_map, dense_precision):
    from position_induced_transformer_b200.pit import head_scale
    from position_induced_transformer_b200.posatt import get_dense_precision, position_attention
    g = torch.Generator().manual_seed(N + D)
    mesh = torch.rand((B, N, 2) if batched else (N, 2), generator=g)
    values = torch.randn(B, N, D, generator=g)
    lmda = torch.rand(H, 1, 1, generator=g) * 1.5 - 1.0
    vc, lc = values.clone().requires_grad_(True), lmda.clone().requires_grad_(True)
    want = po.dense_posatt(mesh, mesh, vc, lc, 1.0, "euclid", self_concat=True)
    up = torch.randn(want.shape, generator=g)
    want.backward(up)
    dense_precision(mode)
    assert get_dense_precision() == mode
    vg = values.to(cuda_device).requires_grad_(True)
    lg = lmda.to(cuda_device).requires_grad_(True)
    mg = mesh.to(cuda_device)
    got = position_attention(mg, mg, vg, head_scale(lg), 1.0, variant="euclid", self_concat=True)
    got.backward(up.to(cuda_device))
    fwd, grad, grad_l = BOUNDS[mode]
    e_out, e_dv, e_dl = rel_linf(got.detach().cpu(), want.detach()), rel_linf(vg.grad.cpu(), vc.grad), rel_linf(lg.grad.cpu(), lc.grad)
    print(f"\ndense {mode} batched={batched} N={N} D={D}: out {e_out:.2e} dU {e_dv:.2e} dlmda {e_dl:.2e}")
    assert e_out <= fwd and e_dv <= grad and e_dl <= grad_l
    if mode != "fp32":
        assert e_out > 2e-6            # the mode really changed the arithmetic (3xTF32 sits at 2e-7, TF32 operands at 1.3e-5)
