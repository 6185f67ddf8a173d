"""Fused decoder tail (attention + two-layer MLP in one kernel) against the CPU oracle's decode()."""
import pytest
import torch

from conftest import rel_linf
from oracle import pit_oracle, posatt_oracle as po

pytestmark = pytest.mark.gpu


def _case(variant, sd, N, M, B, D, H, Cw, O, q, seed):
    g = torch.Generator().manual_seed(seed)
    if variant == "periodic1d":
        mesh_out = torch.linspace(0, 1, N + 1)[:-1].reshape(-1, 1)
        mesh_in = torch.linspace(0, 1, M + 1)[:-1].reshape(-1, 1)
    elif variant == "periodic2d":
        import numpy as np
        def grid(n):
            ax = np.linspace(0, 1, n + 1)[:-1]
            return torch.tensor(np.vstack([m.ravel() for m in np.meshgrid(ax, ax)]).T, dtype=torch.float)
        mesh_out, mesh_in = grid(int(N ** 0.5)), grid(int(M ** 0.5))
    else:
        mesh_out, mesh_in = torch.rand(N, sd, generator=g), torch.rand(M, sd, generator=g)
    p = {
        "up.lmda": torch.rand(H, 1, 1, generator=g) * 2 - 0.5,
        "de.mlp1.weight": torch.randn(Cw, H * D, generator=g) / (H * D) ** 0.5,
        "de.mlp1.bias": torch.randn(Cw, generator=g) * 0.1,
        "de.mlp2.weight": torch.randn(O, Cw, generator=g) / Cw ** 0.5,
        "de.mlp2.bias": torch.randn(O, generator=g) * 0.1,
    }
    feats = torch.randn(B, M, D, generator=g)
    return mesh_out, mesh_in, feats, p, g


@pytest.mark.parametrize("variant,sd,N,M,B,D,H,Cw,O,q", [
    ("euclid", 2, 700, 256, 8, 64, 2, 64, 1, 0.02),      # Darcy-like
    ("euclid", 1, 500, 256, 8, 32, 1, 32, 3, 0.02),      # Sod-like
    ("periodic1d", 1, 512, 128, 4, 64, 2, 64, 1, 0.05),  # Burgers-like
    ("periodic2d", 2, 1024, 64, 2, 32, 2, 256, 2, 0.1),  # wide hidden layer: one sample spans two register groups
    ("euclid", 2, 300, 100, 3, 16, 2, 128, 4, 1.0),      # unmasked
    ("euclid", 2, 97, 40, 1, 8, 1, 512, 1, 0.2),
    ("euclid", 2, 600, 700, 2, 16, 2, 64, 1, 0.01),      # tensor-core tail with 32 columns per lane (M > 256)
    ("periodic2d", 2, 1024, 64, 2, 32, 2, 64, 2, 0.1),   # tensor-core tail, periodic distance, out_dim 2 (run-time out_dim path)
    ("euclid", 1, 33, 256, 8, 32, 2, 32, 4, 0.03),       # tensor-core tail: fewer rows than one round, hidden width 32, out_dim 4
    ("euclid", 2, 5, 6, 2, 8, 1, 32, 1, 0.5),            # tensor-core tail at its smallest: one tile, one chunk, six columns
    ("euclid", 2, 1000, 9, 1, 4, 2, 64, 1, 0.3),         # one sample, nine latent points
])
@pytest.mark.parametrize("tile_plan", [True, False])   # cached tile plan (hidden width 32 / 64) or a latent-mesh scan per launch
def test_decoder_tail_matches_oracle(variant, sd, N, M, B, D, H, Cw, O, q, tile_plan, cuda_device, host_scale_map, request):
    import position_induced_transformer_b200.pit as pit_mod
    from position_induced_transformer_b200 import posatt as posatt_mod
    from position_induced_transformer_b200.posatt import decoder_tail, decoder_tail_supported
    posatt_mod.use_tail_plan(tile_plan)
    request.addfinalizer(lambda: posatt_mod.use_tail_plan(True))
    mesh_out, mesh_in, feats, p, g = _case(variant, sd, N, M, B, D, H, Cw, O, q, seed=N + M + Cw)
    # oracle
    pc = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    fc = feats.clone().requires_grad_(True)
    want = pit_oracle.decode(pc, variant, mesh_in, fc, mesh_out, q)
    up = torch.randn(want.shape, generator=g)
    want.backward(up)
    # fused
    dev = cuda_device
    pg = {k: v.clone().to(dev).requires_grad_(True) for k, v in p.items()}
    fg = feats.clone().to(dev).requires_grad_(True)
    assert decoder_tail_supported(mesh_in.to(dev), fg, H, Cw, O)
    prev = torch.get_float32_matmul_precision()
    torch.set_float32_matmul_precision("highest")
    try:
        got = decoder_tail(mesh_out.to(dev), mesh_in.to(dev), fg, pit_mod.head_scale(pg["up.lmda"]), q,
                           pg["de.mlp1.weight"], pg["de.mlp1.bias"], pg["de.mlp2.weight"], pg["de.mlp2.bias"], variant=variant)
        got.backward(up.to(dev))
    finally:
        torch.set_float32_matmul_precision(prev)
    assert got.shape == want.shape
    assert rel_linf(got.detach().cpu(), want.detach()) <= 1e-5
    assert rel_linf(fg.grad.cpu(), fc.grad) <= 1e-4
    for k in p:
        assert rel_linf(pg[k].grad.cpu(), pc[k].grad, floor=1e-6) <= 1e-4, k


def _grid(n):
    import numpy as np
    ax = np.linspace(0, 1, n)
    return torch.tensor(np.vstack([g.ravel() for g in np.meshgrid(ax, ax)]).T, dtype=torch.float)


@pytest.mark.parametrize("case", ["darcy130", "random", "unmasked", "periodic1d"])
def test_tail_plan_structure(case, cuda_device):
    """The plan is a permutation of the rows cut into 32-row tiles whose candidate lists are sorted, duplicate-free and
    contain every column any head can keep (checked against the oracle's kept mask for several lmda)."""
    from position_induced_transformer_b200 import posatt as pa
    g = torch.Generator().manual_seed(5)
    variant, q = "euclid", 0.02
    if case == "darcy130":
        mesh_out, mesh_in = _grid(130), _grid(16)
    elif case == "random":
        mesh_out, mesh_in = torch.rand(1000, 2, generator=g), torch.rand(300, 2, generator=g)
    elif case == "unmasked":
        mesh_out, mesh_in, q = torch.rand(70, 2, generator=g), torch.rand(40, 2, generator=g), 1.0
    else:
        variant, q = "periodic1d", 0.05
        mesh_out, mesh_in = torch.linspace(0, 1, 513)[:-1].reshape(-1, 1), torch.linspace(0, 1, 129)[:-1].reshape(-1, 1)
    N, M = mesh_out.shape[0], mesh_in.shape[0]
    dev = cuda_device
    values = torch.zeros(2, M, 64, device=dev)
    mo, mi, st, period, stats, _ = pa.prepare_meshes(mesh_out.to(dev), mesh_in.to(dev), values, 2, variant, q)
    plan = pa.build_tail_plan(st, mo, mi, period, stats)
    torch.cuda.synchronize()
    n_tiles = (N + 31) // 32
    assert plan.n_tiles == n_tiles
    rows = plan.rec[:, 3].contiguous().view(torch.int32).cpu()
    assert sorted(rows[rows >= 0].tolist()) == list(range(N)) and int((rows < 0).sum()) == n_tiles * 32 - N
    off = plan.tile_off.cpu().tolist()
    cnt = plan.tile_cnt.cpu().tolist()
    cand = plan.cand.cpu().tolist()
    assert off[0] == 0 and all(b - a == (n + 7) // 8 * 8 for a, b, n in zip(off, off[1:], cnt))
    # records carry the row's statistics, the distance matrix the reference's squared distances bit for bit
    rec = plan.rec.cpu()
    valid = rows >= 0
    for i in range(3 if q < 1.0 else 1):
        assert torch.equal(rec[valid][:, i], stats[i].cpu()[rows[valid].long()])
    d2_ref = po.sqdist(mesh_out, mesh_in, variant)
    d2_plan = plan.d2.cpu()
    # superset of the kept sets
    keep_any = torch.zeros(N, M, dtype=torch.bool)
    for lm in (-1.3, 0.0, 0.4, 1.1, 2.5):
        scale = po.head_scale(torch.full((1, 1, 1), lm))
        _, _, keep = po.exact_weights(mesh_out, mesh_in, scale, q, variant)
        keep_any |= keep[0]
    sizes = []
    for t in range(n_tiles):
        c = cand[off[t]:off[t] + cnt[t]]
        assert c == sorted(set(c)) and all(0 <= j < M for j in c)
        sizes.append(len(c))
        tile_rows = rows[t * 32:(t + 1) * 32]
        need = keep_any[tile_rows[tile_rows >= 0].long()].any(0).nonzero().flatten().tolist()
        assert set(need) <= set(c), (t, need, c)
        live = tile_rows >= 0
        want = d2_ref[tile_rows[live].long()][:, torch.tensor(c, dtype=torch.long)].T
        assert torch.equal(d2_plan[off[t]:off[t] + cnt[t]][:, live], want), t
    if case == "darcy130":     # sorted rows share their candidate set: most tiles fit one k-step of 8, nearly all one block of 16
        assert sum(s <= 16 for s in sizes) >= 0.9 * n_tiles
    if case == "unmasked":
        assert all(s == M for s in sizes)


def test_pit_decoder_uses_the_fused_tail(cuda_device, host_scale_map):
    """pit.decoder takes the fused path for the stock layers and equals the two-module composition."""
    import position_induced_transformer_b200.pit as pit_mod
    from position_induced_transformer_b200 import _cabi, workloads
    torch.manual_seed(0)
    w = workloads.make_darcy(43, batch=2).to(cuda_device)
    gen = torch.Generator().manual_seed(3)
    latent = torch.randn(2, 256, 64, generator=gen).to(cuda_device)
    mesh = w.meshes[0].reshape(-1, 2)
    prev = torch.get_float32_matmul_precision()
    torch.set_float32_matmul_precision("highest")
    try:
        n0 = _cabi.launch_count()
        fused = w.model.decoder(w.model.mesh_ltt, latent, mesh)
        fused_launches = _cabi.launch_count() - n0
        pit_mod.use_fused_decoder_tail(False)
        plain = w.model.decoder(w.model.mesh_ltt, latent, mesh)
        pit_mod.use_fused_decoder_tail(True)
    finally:
        torch.set_float32_matmul_precision(prev)
    assert 1 <= fused_launches <= 6          # scale map + tail kernel (+ row statistics and the three plan-construction launches if this mesh pair was not cached yet)
    assert rel_linf(fused, plain) <= 1e-5
