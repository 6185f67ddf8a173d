"""Parity of the CUDA path (through the C ABI) against the golden vectors of the unmodified reference
and against the CPU oracle.  Tolerances: kept sets identical; forward
rel-L_inf <= 1e-5; gradients <= 1e-4 of their max-norm."""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden, rel_linf, t, variant_of
from oracle import posatt_oracle as po

pytestmark = pytest.mark.gpu

FWD_TOL = 1e-5
GRAD_TOL = 1e-4


def _pa():
    from position_induced_transformer_b200.posatt import position_attention
    return position_attention


@pytest.mark.parametrize("name", golden_names("op_"))
def test_rowstat_is_bit_exact(name, cuda_device):
    from position_induced_transformer_b200 import _cabi, posatt
    g = load_golden("op_" + name)
    variant = variant_of(str(g["cls"]))
    mo, mi = t(g["mesh_out"]), t(g["mesh_in"])
    q = float(g["locality"])
    d2 = po.sqdist(mo, mi, variant)
    k_lo, k_hi, w = _cabi.quantile_ranks(q, mi.shape[-2]) if q < 1.0 else (0, 0, 0.0)
    ref = po.exact_rowstat(d2, k_lo, k_hi)
    vals = t(g["values"], cuda_device)
    st = posatt._Stage(mo.to(cuda_device), mi.to(cuda_device), vals, int(g["n_head"]), variant)
    period = posatt.wrap_period(mi.to(cuda_device), variant)
    got = posatt.row_statistics(st, mo.to(cuda_device), mi.to(cuda_device), period, q)
    for a, b in zip(got[:3], ref):
        assert torch.equal(a.cpu(), b)
    if q < 1.0:
        assert (k_lo, k_hi, w) == po.quantile_ranks(q, mi.shape[-2])


@pytest.mark.parametrize("case", ["random", "few_columns", "duplicates", "lattice_ties", "rank_31", "rank_32", "batched"])
def test_rowstat_small_rank_path_is_bit_exact(case, cuda_device):
    """The candidate-compaction path of the warp kernel (k_hi < 32) and its fall-backs (too few lanes with keys, more than
    four candidates in a lane) against a full sort."""
    from position_induced_transformer_b200 import _cabi, posatt
    g = torch.Generator().manual_seed(11)
    n, m, q = 300, 700, 0.02
    mo, mi = torch.rand(n, 2, generator=g), torch.rand(m, 2, generator=g)
    if case == "few_columns":
        m, q = 20, 0.4                                   # fewer valid keys than lanes: no bound, full path
        mi = torch.rand(m, 2, generator=g)
    elif case == "duplicates":
        mi = mi[torch.randint(0, 40, (m,), generator=g)]  # 40 distinct points: every distance 17-fold tied
    elif case == "lattice_ties":
        ax = torch.linspace(0, 1, 26)
        mi = torch.stack(torch.meshgrid(ax, ax, indexing="ij"), -1).reshape(-1, 2)
        mo = mi[::3].clone()                              # rows sit on lattice points: 4- and 8-fold ties
        m = mi.shape[0]
    elif case == "rank_31":
        q = 30.5 / (m - 1)
    elif case == "rank_32":
        q = 31.5 / (m - 1)                               # k_hi = 32: just outside the small-rank path
    elif case == "batched":
        mo, mi = torch.rand(3, n, 2, generator=g), torch.rand(3, m, 2, generator=g)
    d2 = po.sqdist(mo, mi, "euclid")
    k_lo, k_hi, w = _cabi.quantile_ranks(q, m)
    ref = po.exact_rowstat(d2, k_lo, k_hi)
    vals = torch.zeros((mi.shape[0] if mi.dim() == 3 else 1), m, 4, device=cuda_device)
    st = posatt._Stage(mo.to(cuda_device), mi.to(cuda_device), vals, 1, "euclid")
    got = posatt.row_statistics(st, mo.to(cuda_device), mi.to(cuda_device), None, q)
    for a, b in zip(got[:3], ref):
        assert torch.equal(a.cpu(), b)


@pytest.mark.parametrize("name", golden_names("op_"))
def test_forward_backward_match_reference(name, cuda_device):
    g = load_golden("op_" + name)
    variant = variant_of(str(g["cls"]))
    concat = str(g["kind"]) == "self"
    mo, mi = t(g["mesh_out"], cuda_device), t(g["mesh_in"], cuda_device)
    vals = t(g["values"], cuda_device).requires_grad_(True)
    # the per-head scale is fed exactly as the reference computed it on CPU, so the parity of the kernels does
    # not depend on the last ulp of the device's tan/sin
    scale = t(g["scale"], cuda_device).requires_grad_(True)
    out = _pa()(mo, mi, vals, scale, float(g["locality"]), variant, concat)
    assert out.shape == g["out"].shape
    assert rel_linf(out.detach().cpu(), t(g["out"])) <= FWD_TOL
    out.backward(t(g["upstream"], cuda_device))
    assert rel_linf(vals.grad.cpu(), t(g["d_values"])) <= GRAD_TOL
    # d_scale reference: oracle autograd with the same scale as a leaf
    s_cpu = t(g["scale"]).requires_grad_(True)
    v_cpu = t(g["values"])
    o_cpu = po.dense_contract(po.dense_attention(t(g["mesh_out"]), t(g["mesh_in"]), s_cpu, float(g["locality"]), variant), v_cpu)
    if concat:
        o_cpu = torch.cat((v_cpu, o_cpu), -1)
    o_cpu.backward(t(g["upstream"]))
    assert rel_linf(scale.grad.cpu().reshape(-1), s_cpu.grad.reshape(-1)) <= GRAD_TOL


@pytest.mark.parametrize("name", [n for n in golden_names("op_")])
def test_kept_set_equals_reference(name, cuda_device, host_scale_map):
    """Attention exported through the fused kernel (identity values) has the reference's support."""
    import position_induced_transformer_b200.pit as pit_mod
    g = load_golden("op_" + name)
    layer = getattr(pit_mod, str(g["cls"]))(int(g["n_head"]), 1, float(g["locality"])).to(cuda_device)
    mo, mi = t(g["mesh_out"], cuda_device), t(g["mesh_in"], cuda_device)
    with torch.no_grad():
        layer.lmda.copy_(t(g["lmda"], cuda_device))
        att = layer.dist2att(mo, mi, layer.lmda, layer.locality).cpu()
    shape = tuple(g["att_shape"])
    assert tuple(att.shape) == shape
    ref_pos = np.unpackbits(g["kept_bits"])[: int(np.prod(shape))].reshape(shape).astype(bool)
    # with the host scale map the H scales are bit-identical to the ones the reference run used, so the comparison is unconditional
    assert torch.equal(pit_mod.head_scale(layer.lmda.detach()).cpu(), t(g["scale"]))
    ours = (att > 0).numpy()
    if float(g["locality"]) < 1.0:
        assert (ours == ref_pos).all()
    else:  # global stage: only underflow may differ, and only where the weight is negligible
        assert (att.numpy()[ours != ref_pos] < 1e-30).all()
    assert torch.allclose(att.sum(-1), torch.ones(shape[:-1]), atol=1e-5)


@pytest.mark.parametrize("name", golden_names("op_"))
def test_module_forward(name, cuda_device, host_scale_map):
    """Whole layer (lmda -> scale -> fused op, autograd back to lmda) against the reference output."""
    import position_induced_transformer_b200.pit as pit_mod
    g = load_golden("op_" + name)
    layer = getattr(pit_mod, str(g["cls"]))(int(g["n_head"]), g["values"].shape[-1], float(g["locality"])).to(cuda_device)
    with torch.no_grad():
        layer.lmda.copy_(t(g["lmda"], cuda_device))
    vals = t(g["values"], cuda_device).requires_grad_(True)
    mo, mi = t(g["mesh_out"], cuda_device), t(g["mesh_in"], cuda_device)
    out = layer(mo, vals) if str(g["kind"]) == "self" else layer(mo, mi, vals)
    out.backward(t(g["upstream"], cuda_device))
    assert rel_linf(out.detach().cpu(), t(g["out"])) <= FWD_TOL
    assert rel_linf(vals.grad.cpu(), t(g["d_values"])) <= GRAD_TOL
    assert rel_linf(layer.lmda.grad.cpu(), t(g["d_lmda"])) <= GRAD_TOL


def test_softmax_rows_reproduce_constants(cuda_device):
    """Size-independent property at the Darcy-421 decoder's full size: weights of a row sum to one, so a
    constant value field is reproduced, and the op is linear in the values."""
    ax = np.linspace(0, 1, 421)
    mesh = torch.tensor(np.vstack([m.ravel() for m in np.meshgrid(ax, ax)]).T, dtype=torch.float, device=cuda_device)
    ax = np.linspace(0, 1, 16)
    ltt = torch.tensor(np.vstack([m.ravel() for m in np.meshgrid(ax, ax)]).T, dtype=torch.float, device=cuda_device)
    gen = torch.Generator(device="cpu").manual_seed(5)
    scale = torch.tensor([1.7, 4.2], device=cuda_device)
    ones = torch.ones(2, 256, 8, device=cuda_device)
    out = _pa()(mesh, ltt, ones, scale, 0.02, "euclid")
    assert out.shape == (2, 177241, 16)
    assert float((out - 1).abs().max()) <= 2e-6
    u1 = torch.randn(2, 256, 8, generator=gen).to(cuda_device)
    u2 = torch.randn(2, 256, 8, generator=gen).to(cuda_device)
    lin = _pa()(mesh, ltt, 2.0 * u1 - 0.5 * u2, scale, 0.02, "euclid")
    ref = 2.0 * _pa()(mesh, ltt, u1, scale, 0.02, "euclid") - 0.5 * _pa()(mesh, ltt, u2, scale, 0.02, "euclid")
    assert float((lin - ref).abs().max()) <= 1e-5
    # full-size parity on a random subset of rows (rows are independent)
    rows = torch.randperm(177241, generator=gen)[:384]
    want = po.dense_contract(po.dense_attention(mesh.cpu()[rows], ltt.cpu(), scale.cpu().reshape(-1, 1, 1), 0.02), u1.cpu())
    got = _pa()(mesh, ltt, u1, scale, 0.02, "euclid").cpu()[:, rows]
    assert rel_linf(got, want) <= FWD_TOL


@pytest.fixture(params=[True, False], ids=["tile_plan", "scan"])
def column_plan(request):
    """Run a wide (encoder) stage with the cached column plan and with the per-launch row walk."""
    from position_induced_transformer_b200 import posatt
    posatt.use_tail_plan(request.param)
    posatt.mesh_cache.clear()
    yield request.param
    posatt.use_tail_plan(True)


@pytest.mark.parametrize("variant,q", [("euclid", 0.05), ("euclid", 0.004), ("periodic2d", 0.02)])
def test_wide_encoder_irregular_mesh(variant, q, column_plan, cuda_device):
    """Few rows, thousands of columns, narrow values: the local-encoder kernels on a mesh without any coherence."""
    gen = torch.Generator().manual_seed(31)
    if variant == "periodic2d":
        ax = np.linspace(0, 1, 81)[:-1]
        mi = torch.tensor(np.vstack([m.ravel() for m in np.meshgrid(ax, ax)]).T, dtype=torch.float)
        ax = np.linspace(0, 1, 7)[:-1]
        mo = torch.tensor(np.vstack([m.ravel() for m in np.meshgrid(ax, ax)]).T, dtype=torch.float)
    else:
        mo, mi = torch.rand(50, 2, generator=gen), torch.rand(5000, 2, generator=gen)
    vals = torch.randn(3, mi.shape[0], 5, generator=gen)
    scale = torch.tensor([0.8, 4.0])
    s_cpu = scale.clone().reshape(-1, 1, 1).requires_grad_(True)
    want = po.dense_contract(po.dense_attention(mo, mi, s_cpu, q, variant), vals)
    up = torch.randn(want.shape, generator=gen)
    want.backward(up)
    s_gpu = scale.to(cuda_device).requires_grad_(True)
    got = _pa()(mo.to(cuda_device), mi.to(cuda_device), vals.to(cuda_device), s_gpu, q, variant)
    got.backward(up.to(cuda_device))
    assert rel_linf(got.detach().cpu(), want.detach()) <= FWD_TOL
    assert rel_linf(s_gpu.grad.cpu().reshape(-1), s_cpu.grad.reshape(-1)) <= GRAD_TOL


def test_darcy421_encoder_full_size(column_plan, cuda_device):
    """Encoder stage at BASELINE size (256 x 177241, split along the columns) against the oracle."""
    ax = np.linspace(0, 1, 421)
    mesh = torch.tensor(np.vstack([m.ravel() for m in np.meshgrid(ax, ax)]).T, dtype=torch.float)
    ax = np.linspace(0, 1, 16)
    ltt = torch.tensor(np.vstack([m.ravel() for m in np.meshgrid(ax, ax)]).T, dtype=torch.float)
    gen = torch.Generator().manual_seed(11)
    vals = torch.randn(2, 177241, 3, generator=gen)
    scale = torch.tensor([1.3, 6.0])
    rows = torch.arange(0, 256, 8)
    s_cpu = scale.clone().reshape(-1, 1, 1).requires_grad_(True)
    want = po.dense_contract(po.dense_attention(ltt[rows], mesh, s_cpu, 0.02), vals)
    up = torch.randn(2, 256, 6, generator=gen)
    want.backward(up[:, rows])
    s_gpu = scale.to(cuda_device).requires_grad_(True)
    got = _pa()(ltt.to(cuda_device), mesh.to(cuda_device), vals.to(cuda_device), s_gpu, 0.02, "euclid")
    assert rel_linf(got.detach().cpu()[:, rows], want.detach()) <= FWD_TOL
    # scale gradient restricted to the same rows: zero the upstream elsewhere
    up_masked = torch.zeros_like(up)
    up_masked[:, rows] = up[:, rows]
    got.backward(up_masked.to(cuda_device))
    assert rel_linf(s_gpu.grad.cpu().reshape(-1), s_cpu.grad.reshape(-1)) <= GRAD_TOL


@pytest.mark.parametrize("shape", [dict(B=1, N=1, M=1, D=1, H=1), dict(B=2, N=3, M=33, D=5, H=3), dict(B=1, N=40, M=1025, D=4, H=1),
                                   dict(B=3, N=70, M=2500, D=7, H=2), dict(B=2, N=9, M=129, D=160, H=2)])
@pytest.mark.parametrize("batched", [False, True])
def test_ragged_shapes_against_oracle(shape, batched, cuda_device):
    gen = torch.Generator().manual_seed(sum(shape.values()))
    B, N, M, D, H = (shape[k] for k in "BNMDH")
    mo = torch.rand((B, N, 2) if batched else (N, 2), generator=gen)
    mi = torch.rand((B, M, 2) if batched else (M, 2), generator=gen)
    vals = torch.randn(B, M, D, generator=gen)
    scale = torch.rand(H, generator=gen) * 5 + 0.5
    for q in (0.3, 1.0):
        s_cpu = scale.clone().reshape(-1, 1, 1).requires_grad_(True)
        v_cpu = vals.clone().requires_grad_(True)
        want = po.dense_contract(po.dense_attention(mo, mi, s_cpu, q), v_cpu)
        up = torch.randn(want.shape, generator=gen)
        want.backward(up)
        s_gpu = scale.to(cuda_device).requires_grad_(True)
        v_gpu = vals.to(cuda_device).requires_grad_(True)
        got = _pa()(mo.to(cuda_device), mi.to(cuda_device), v_gpu, s_gpu, q, "euclid")
        got.backward(up.to(cuda_device))
        assert rel_linf(got.detach().cpu(), want.detach()) <= FWD_TOL
        assert rel_linf(v_gpu.grad.cpu(), v_cpu.grad) <= GRAD_TOL
        # a single-column row has an exactly zero scale gradient: compare against the gradient's natural size
        assert rel_linf(s_gpu.grad.cpu().reshape(-1), s_cpu.grad.reshape(-1), floor=1e-3) <= GRAD_TOL


def test_errors_are_loud(cuda_device):
    pa = _pa()
    mesh = torch.rand(8, 2, device=cuda_device)
    with pytest.raises(RuntimeError):
        pa(mesh, mesh, torch.rand(1, 9, 4, device=cuda_device), torch.ones(1, device=cuda_device), 0.5)   # M mismatch
    with pytest.raises(RuntimeError):
        pa(torch.rand(8, 3, device=cuda_device), torch.rand(8, 3, device=cuda_device), torch.rand(1, 8, 4, device=cuda_device),
           torch.ones(1, device=cuda_device), 0.5)                                                           # space_dim 3
    with pytest.raises(RuntimeError):
        pa(mesh, mesh, torch.rand(1, 8, 4, device=cuda_device).double(), torch.ones(1, device=cuda_device), 0.5)


def test_rowstat_cache_hits_and_invalidates(cuda_device):
    from position_induced_transformer_b200 import posatt
    pa = _pa()
    gen = torch.Generator().manual_seed(21)
    mo = torch.rand(50, 2, generator=gen).to(cuda_device)
    mi = torch.rand(300, 2, generator=gen).to(cuda_device)
    vals = torch.randn(2, 300, 4, generator=gen).to(cuda_device)
    scale = torch.tensor([2.0], device=cuda_device)
    posatt.rowstat_cache.clear()
    h0, m0 = posatt.rowstat_cache.hits, posatt.rowstat_cache.misses
    a = pa(mo, mi, vals, scale, 0.1)
    b = pa(mo.reshape(-1, 2), mi, vals, scale, 0.1)          # a fresh view of the same storage hits
    assert posatt.rowstat_cache.misses == m0 + 1 and posatt.rowstat_cache.hits == h0 + 1
    assert torch.equal(a, b)
    mi[0, 0] += 0.25                                          # in-place write bumps the version: miss, new statistics
    c = pa(mo, mi, vals, scale, 0.1)
    assert posatt.rowstat_cache.misses == m0 + 2
    want = po.dense_contract(po.dense_attention(mo.cpu(), mi.cpu(), scale.cpu().reshape(-1, 1, 1), 0.1), vals.cpu())
    assert rel_linf(c.cpu(), want) <= FWD_TOL


def test_head_scale_matches_torch_ops_bit_for_bit(cuda_device):
    """The one-launch scale map equals the reference's chain of torch CUDA ops (pit.py:48) exactly, and so does its derivative
    up to rounding."""
    from math import pi
    from position_induced_transformer_b200.posatt import head_scale_cuda
    g = torch.Generator().manual_seed(5)
    lmda = (torch.rand(100000, 1, 1, generator=g) * 8 - 4).to(cuda_device)
    lmda[:1000] = torch.rand(1000, 1, 1, generator=g).to(cuda_device)      # the initialisation range of pit.py:35
    a = lmda.clone().requires_grad_(True)
    b = lmda.clone().requires_grad_(True)
    want = torch.tan(0.25 * pi * (1 - 1e-7) * (1.0 + torch.sin(a)))
    got = head_scale_cuda(b)
    assert got.shape == want.shape
    assert torch.equal(got, want)
    up = torch.randn(want.shape, generator=g).to(cuda_device)
    want.backward(up)
    got.backward(up)
    assert b.grad.shape == a.grad.shape
    err = ((b.grad - a.grad).abs() / a.grad.abs().clamp_min(1e-30)).max()
    assert float(err) <= 1e-5
