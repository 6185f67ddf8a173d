"""Host-side logic that needs no GPU: the CPU branches of the drop-in module equal the reference expressions."""
import pytest
import torch

import position_induced_transformer_b200.pit as pit_mod
from position_induced_transformer_b200.utils import RelLpNorm, RelMaxNorm


def test_kaiming_mlp_cpu_branch_is_the_reference_expression():
    torch.manual_seed(0)
    mlp = pit_mod.kaiming_mlp(6, 16, 8)
    x = torch.randn(3, 11, 6)
    want = mlp.mlp2(torch.nn.functional.gelu(mlp.mlp1(x)))          # pit.py:25-26
    assert torch.equal(mlp(x), want)
    assert torch.equal(mlp.forward_gelu(x), torch.nn.functional.gelu(want))
    assert list(mlp.state_dict()) == ["mlp1.weight", "mlp1.bias", "mlp2.weight", "mlp2.bias"]


def test_block_activation_helper_accepts_any_module():
    """pit.encoder / pit.processor apply gelu(module(x)); scripts may install their own en_layer (train_naca.py:45)."""
    torch.manual_seed(0)
    custom = torch.nn.Linear(4, 4)
    x = torch.randn(2, 5, 4)
    assert torch.equal(pit_mod.pit._mlp_gelu(custom, x), torch.nn.functional.gelu(custom(x)))
    stock = pit_mod.kaiming_mlp(4, 8, 4)
    assert torch.equal(pit_mod.pit._mlp_gelu(stock, x), torch.nn.functional.gelu(stock(x)))


def test_rel_lp_norm_cpu_branch_is_the_reference_expression():
    g = torch.Generator().manual_seed(1)
    true, pred = torch.randn(4, 30, 2, generator=g), torch.randn(4, 30, 2, generator=g)
    for p in (1, 2):
        want = (torch.norm(true - pred, p=p, dim=1) / torch.norm(true, p=p, dim=1)).mean(dim=-1).sum()   # utils.py:60-98
        assert torch.equal(RelLpNorm(2, p)(true, pred), want)
    want = ((true - pred).abs().amax(dim=1) / true.abs().amax(dim=1)).mean(dim=-1).sum()
    assert torch.equal(RelMaxNorm(2)(true, pred), want)


def test_head_scale_cpu_branch_is_the_reference_expression():
    from math import pi
    lmda = torch.rand(3, 1, 1, generator=torch.Generator().manual_seed(2))
    assert torch.equal(pit_mod.head_scale(lmda), torch.tan(0.25 * pi * (1 - 1e-7) * (1.0 + torch.sin(lmda))))   # pit.py:48


def test_mesh_cache_skips_tensors_without_version_counter():
    """Inference tensors have no version counter: they must bypass the cache instead of raising."""
    from position_induced_transformer_b200.posatt import _MeshCache
    cache = _MeshCache()
    ok = torch.zeros(4, 2)
    with torch.inference_mode():
        inf = torch.zeros(4, 2)
    assert cache.key(ok, ok, "euclid", 0.5) is not None
    assert cache.key(inf, ok, "euclid", 0.5) is None and cache.key(ok, inf, "euclid", 0.5) is None


def test_mesh_cache_is_lru():
    from position_induced_transformer_b200.posatt import _MeshCache, _MeshEntry
    cache = _MeshCache(capacity=2)
    for k in ("a", "b"):
        cache.put(k, _MeshEntry(None, None, None, None, None))
    assert cache.get("a") is not None          # "a" becomes the most recently used
    cache.put("c", _MeshEntry(None, None, None, None, None))
    assert cache.get("b") is None and cache.get("a") is not None and cache.get("c") is not None


def test_meshes_requiring_grad_are_refused():
    from position_induced_transformer_b200.posatt import _meshes_are_constants
    m = torch.zeros(4, 2, requires_grad=True)
    with pytest.raises(RuntimeError, match="mesh coordinates"):
        _meshes_are_constants(m, torch.zeros(4, 2))
    with torch.no_grad():
        _meshes_are_constants(m, m)


def test_tall_linear_weight_gradient_by_slices_matches_plain_autograd():
    """The split-K weight gradient of very tall Linears (pit._TallLinear) is the ordinary gradient, remainder rows included."""
    import position_induced_transformer_b200.pit as pit_mod
    g = torch.Generator().manual_seed(2)
    x = torch.randn(3, 1001, 16, generator=g, requires_grad=True)
    w = torch.randn(5, 16, generator=g, requires_grad=True)
    up = torch.randn(3, 1001, 5, generator=g)
    splits = pit_mod._TALL_SPLITS
    pit_mod._TALL_SPLITS = 4                      # 3003 rows = 4 slices of 750 + 3 remainder rows
    try:
        pit_mod._TallLinear.apply(x, w).backward(up)
    finally:
        pit_mod._TALL_SPLITS = splits
    x2, w2 = x.detach().clone().requires_grad_(True), w.detach().clone().requires_grad_(True)
    torch.nn.functional.linear(x2, w2).backward(up)
    assert torch.allclose(x.grad, x2.grad) and torch.allclose(w.grad, w2.grad, rtol=1e-5, atol=1e-5)
