"""The one-launch narrow-input MLP (the encoder lift, pit.py:110-111) against plain fp32 torch ops."""
import pytest
import torch

from conftest import rel_linf

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,k,d", [(2048, 6, 64), (2048, 4, 32), (100, 32, 64), (33, 1, 32)])
@pytest.mark.parametrize("act_out", [True, False])
def test_mlp_fused_matches_torch(rows, k, d, act_out, cuda_device):
    from position_induced_transformer_b200.posatt import mlp_fused, mlp_fused_supported
    g = torch.Generator().manual_seed(rows + k)
    x = torch.randn(4, rows // 4 if rows % 4 == 0 else rows, k, generator=g)[: (4 if rows % 4 == 0 else 1)]
    w1, b1 = torch.randn(d, k, generator=g) / k ** 0.5, torch.randn(d, generator=g) * 0.1
    w2, b2 = torch.randn(d, d, generator=g) / d ** 0.5, torch.randn(d, generator=g) * 0.1
    leaves_c = [t.clone().requires_grad_(True) for t in (x, w1, b1, w2, b2)]
    y = torch.nn.functional.linear(torch.nn.functional.gelu(torch.nn.functional.linear(leaves_c[0], leaves_c[1], leaves_c[2])), leaves_c[3], leaves_c[4])
    want = torch.nn.functional.gelu(y) if act_out else y
    up = torch.randn(want.shape, generator=g)
    want.backward(up)
    leaves_g = [t.to(cuda_device).requires_grad_(True) for t in (x, w1, b1, w2, b2)]
    assert mlp_fused_supported(leaves_g[0], leaves_g[1], leaves_g[3])
    prev = torch.get_float32_matmul_precision()
    torch.set_float32_matmul_precision("highest")
    try:
        got = mlp_fused(*leaves_g, act_out)
        got.backward(up.to(cuda_device))
    finally:
        torch.set_float32_matmul_precision(prev)
    assert rel_linf(got.detach().cpu(), want.detach()) <= 1e-5
    for a, b, name in zip(leaves_g, leaves_c, ("x", "w1", "b1", "w2", "b2")):
        assert rel_linf(a.grad.cpu(), b.grad) <= 1e-4, name


def test_encoder_lift_takes_the_fused_path(cuda_device):
    import position_induced_transformer_b200.pit as pit_mod
    from position_induced_transformer_b200 import _cabi
    mlp = pit_mod.kaiming_mlp(6, 64, 64).to(cuda_device)
    x = torch.randn(8, 256, 6, device=cuda_device)
    before = _cabi.launch_count()
    out = mlp.forward_gelu(x)
    assert _cabi.launch_count() - before == 1                      # the whole lift is one launch
    ref = torch.nn.functional.gelu(mlp.mlp2(torch.nn.functional.gelu(mlp.mlp1(x))))
    assert rel_linf(out, ref) <= 2e-3                              # torch ran TF32 Linears here ('high', pit.py:2)
