"""Pin the oracle (oracle/*.py) to the golden vectors produced by the unmodified reference
(oracle/gen_golden.py -> tests/golden/*.npz).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden, rel_linf, t, variant_of
from oracle import pit_oracle, posatt_oracle as po


@pytest.mark.parametrize("name", golden_names("op_"))
def test_dense_oracle_matches_reference_bitwise(name):
    g = load_golden("op_" + name)
    variant = variant_of(str(g["cls"]))
    mo, mi = t(g["mesh_out"]), t(g["mesh_in"])
    vals = t(g["values"]).requires_grad_(True)
    lmda = t(g["lmda"]).requires_grad_(True)
    out = po.dense_posatt(mo, mi, vals, lmda, float(g["locality"]), variant, self_concat=str(g["kind"]) == "self")
    out.backward(t(g["upstream"]))
    # same ops in the same order on the same CPU: bit-identical
    assert torch.equal(out.detach(), t(g["out"]))
    assert torch.equal(vals.grad, t(g["d_values"]))
    assert torch.equal(lmda.grad, t(g["d_lmda"]))
    assert torch.equal(po.head_scale(lmda.detach()), t(g["scale"]))


@pytest.mark.parametrize("name", golden_names("op_"))
def test_exact_restatement_keeps_the_reference_set(name):
    g = load_golden("op_" + name)
    variant = variant_of(str(g["cls"]))
    q = float(g["locality"])
    p, rowsum, keep = po.exact_weights(t(g["mesh_out"]), t(g["mesh_in"]), t(g["scale"]), q, variant)
    shape = tuple(g["att_shape"])
    ref_pos = np.unpackbits(g["kept_bits"])[: int(np.prod(shape))].reshape(shape).astype(bool)
    ours = keep.numpy()
    assert ours.shape == shape
    # every weight the reference left positive is kept; anything extra we keep has underflowed there
    assert not (ref_pos & ~ours).any()
    extra = ours & ~ref_pos
    att = (p / rowsum.unsqueeze(-1)).numpy()
    assert (att[extra] < 1e-37).all()
    if q < 1.0:
        # masked stages never underflow in these fixtures: sets are identical
        assert (g["kept_per_row"] == ours.sum(-1)).all()
        assert (ref_pos == ours).all()
    vals, lmda = t(g["values"]), t(g["lmda"])
    out = po.exact_posatt(t(g["mesh_out"]), t(g["mesh_in"]), vals, lmda, q, variant, self_concat=str(g["kind"]) == "self")
    assert rel_linf(out, t(g["out"])) <= 2e-6


@pytest.mark.parametrize("m,q", [(51, 0.02), (101, 0.02), (120, 0.02), (256, 0.02), (972, 0.02), (1024, 0.02),
                                 (1849, 0.02), (2048, 0.02), (177241, 0.02), (4390, 0.01), (256, 0.5), (7, 0.99)])
def test_quantile_rank_arithmetic(m, q):
    k_lo, k_hi, w = po.quantile_ranks(q, m)
    x = torch.rand(3, m, generator=torch.Generator().manual_seed(m))
    srt = torch.sort(x, -1).values
    ours = torch.lerp(srt[:, k_lo], srt[:, k_hi], torch.tensor(w))
    assert torch.equal(ours, torch.quantile(x, q, dim=-1))


def _params(g):
    return {k[6:]: t(v) for k, v in g.items() if k.startswith("param/")}


@pytest.mark.parametrize("name", golden_names("model_"))
def test_model_oracle_matches_reference(name):
    g = load_golden("model_" + name)
    params = {k: v.requires_grad_(True) for k, v in _params(g).items()}
    ins = [t(g[f"input/{i}"]) for i in range(sum(k.startswith("input/") for k in g))]
    en_loc, de_loc = float(g["ctor/en_loc"]), float(g["ctor/de_loc"])
    cls = str(g["cls"])
    if cls == "pit":
        out = pit_oracle.forward_point_cloud(params, ins[0], ins[1], ins[2], ins[3], en_loc, de_loc)
    else:
        sd = int(g["ctor/space_dim"])
        out = pit_oracle.forward_shared_mesh(params, variant_of(cls), ins[0], ins[1], t(g["ctor/mesh_ltt"]).reshape(-1, sd),
                                             ins[2], en_loc, de_loc, instance_norm=name.endswith("_norm"))
    assert torch.equal(out.detach(), t(g["out"]))
    loss = pit_oracle.rel_lp_loss(t(g["target"]), out, int(g["ctor/out_dim"]), int(g["loss_p"]))
    assert torch.equal(loss.detach(), t(g["loss"]))
    loss.backward()
    for k, v in params.items():
        assert torch.equal(v.grad, t(g["grad/" + k])), k
