"""Fused bias/GELU epilogues of kaiming_mlp (pit.py:21-26, 111, 121) against the torch ops they replace."""
import pytest
import torch

from conftest import rel_linf

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,apply_gelu", [((8, 256, 64), True), ((8, 256, 64), False), ((3, 1000, 32), True),
                                              ((10, 972, 256), True), ((1, 7, 128), False), ((2, 5, 4), True)])
def test_bias_act_matches_torch(shape, apply_gelu, cuda_device):
    from position_induced_transformer_b200.posatt import bias_act, bias_act_supported
    g = torch.Generator().manual_seed(sum(shape))
    z = (torch.randn(shape, generator=g) * 2).to(cuda_device)
    b = torch.randn(shape[-1], generator=g).to(cuda_device)
    up = torch.randn(shape, generator=g).to(cuda_device)
    za, ba = z.clone().requires_grad_(True), b.clone().requires_grad_(True)
    zb, bb = z.clone().requires_grad_(True), b.clone().requires_grad_(True)
    assert bias_act_supported(za, ba)
    want = za + ba
    want = torch.nn.functional.gelu(want) if apply_gelu else want
    got = bias_act(zb, bb, apply_gelu)
    want.backward(up)
    got.backward(up)
    assert rel_linf(got.detach(), want.detach()) <= 1e-6
    assert rel_linf(zb.grad, za.grad) <= 1e-6
    assert rel_linf(bb.grad, ba.grad) <= 1e-5          # column sums over up to 10k rows: summation order differs


def test_unsupported_width_falls_back_to_torch_ops(cuda_device):
    """Widths the epilogue kernel does not cover (not a multiple of 4 dividing 1024) run as plain torch ops inside kaiming_mlp."""
    import position_induced_transformer_b200.pit as pit_mod
    from position_induced_transformer_b200.posatt import bias_act_supported
    torch.manual_seed(0)
    mlp = pit_mod.kaiming_mlp(6, 24, 10).to(cuda_device)
    x = torch.randn(4, 50, 6, device=cuda_device)
    assert not bias_act_supported(torch.empty(4, 50, 24, device=cuda_device), mlp.mlp1.bias)
    want = mlp.mlp2(torch.nn.functional.gelu(mlp.mlp1(x)))
    assert rel_linf(mlp(x), want) <= 1e-6
    assert rel_linf(mlp.forward_gelu(x), torch.nn.functional.gelu(want)) <= 1e-6


def test_kaiming_mlp_fused_equals_plain(cuda_device):
    """kaiming_mlp with the fused epilogues equals Linear -> GELU -> Linear (-> GELU), values and every gradient."""
    import position_induced_transformer_b200.pit as pit_mod
    prev = torch.get_float32_matmul_precision()
    torch.set_float32_matmul_precision("highest")
    try:
        torch.manual_seed(0)
        mlp = pit_mod.kaiming_mlp(192, 64, 64).to(cuda_device)
        x = torch.randn(8, 256, 192, device=cuda_device)
        up = torch.randn(8, 256, 64, device=cuda_device)
        res = {}
        for fused in (True, False):
            pit_mod.use_fused_mlp_epilogue(fused)
            xi = x.clone().requires_grad_(True)
            mlp.zero_grad()
            out = mlp.forward_gelu(xi)
            out.backward(up)
            res[fused] = (out.detach(), xi.grad, [p.grad.clone() for p in mlp.parameters()])
        pit_mod.use_fused_mlp_epilogue(True)
    finally:
        torch.set_float32_matmul_precision(prev)
    assert rel_linf(res[True][0], res[False][0]) <= 1e-6
    assert rel_linf(res[True][1], res[False][1]) <= 1e-5
    for a, b in zip(res[True][2], res[False][2]):
        assert rel_linf(a, b) <= 1e-5


@pytest.mark.parametrize("shape,p", [((8, 1849, 1), 2), ((8, 1024, 1), 1), ((3, 2048, 3), 2), ((2, 500, 4), 1), ((20, 11271, 4), 2),
                                     ((1, 7, 2), 2)])
def test_rel_lp_loss_matches_torch(shape, p, cuda_device):
    """Fused RelLpNorm (utils.py:60-98) against the torch expression it replaces: value and gradient."""
    from position_induced_transformer_b200 import utils as u
    g = torch.Generator().manual_seed(sum(shape) + p)
    true = torch.randn(shape, generator=g).to(cuda_device)
    pred = (true.cpu() + 0.3 * torch.randn(shape, generator=g)).to(cuda_device)
    loss_fn = u.RelLpNorm(shape[-1], p)
    a = pred.clone().requires_grad_(True)
    b = pred.clone().requires_grad_(True)
    got = loss_fn(true, a)
    want = (torch.norm(true - b, p=p, dim=1) / torch.norm(true, p=p, dim=1)).mean(dim=-1).sum()
    (3.0 * got).backward()
    (3.0 * want).backward()
    assert got.shape == want.shape
    assert abs(float(got) - float(want)) <= 2e-6 * abs(float(want))
    assert rel_linf(a.grad, b.grad) <= 1e-5
