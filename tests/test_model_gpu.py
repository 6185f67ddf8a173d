"""Whole-model parity on the GPU: modules of position_induced_transformer_b200.pit loaded with a reference
state_dict reproduce the reference's output, loss and parameter gradients (golden vectors made by the
unmodified reference on CPU in fp32, so the Linears are run with matmul precision 'highest' here)."""
import pytest
import torch

from conftest import golden_names, load_golden, rel_linf, t

pytestmark = pytest.mark.gpu

MODEL_CLASS = {"burgers": "BurgersPiT", "sod": "SodPiT", "darcy43": "DarcyPiT", "elasticity": "ElasticityPiT",
               "naca": "NacaPiT", "vorticity": "Periodic2dPiT", "vorticity_norm": "VorticityPiT"}


@pytest.mark.parametrize("name", golden_names("model_"))
def test_model_matches_reference(name, cuda_device, host_scale_map):
    from position_induced_transformer_b200 import workloads
    from position_induced_transformer_b200.utils import RelLpNorm
    g = load_golden("model_" + name)
    ctor = {k[5:]: g[k] for k in g if k.startswith("ctor/")}
    mesh = None if ctor["mesh_ltt"].ndim == 0 else t(ctor["mesh_ltt"], cuda_device)
    args = [int(ctor[k]) for k in ("space_dim", "in_dim", "out_dim", "hid_dim", "n_head", "n_blocks")]
    model = getattr(workloads, MODEL_CLASS[name])(*args, mesh, float(ctor["en_loc"]), float(ctor["de_loc"])).to(cuda_device)
    model.load_state_dict({k[6:]: t(v) for k, v in g.items() if k.startswith("param/")})
    ins = [t(g[f"input/{i}"], cuda_device) for i in range(sum(k.startswith("input/") for k in g))]
    prev = torch.get_float32_matmul_precision()
    torch.set_float32_matmul_precision("highest")
    try:
        if name in ("elasticity", "naca"):
            # golden inputs: mesh_in, func_in, mesh_ltt, mesh_out (latent mesh given explicitly)
            latent = model.encoder(ins[0], ins[1], ins[2])
            latent = model.processor(latent, ins[2])
            out = model.decoder(ins[2], latent, ins[3])
        else:
            out = model(ins[0], ins[1], ins[2])
        loss = RelLpNorm(int(ctor["out_dim"]), int(g["loss_p"]))(t(g["target"], cuda_device), out)
        loss.backward()
    finally:
        torch.set_float32_matmul_precision(prev)
    assert out.shape == g["out"].shape
    errs = {k: rel_linf(p.grad.cpu(), t(g["grad/" + k])) for k, p in model.named_parameters()}
    print(f"\nmodel_{name}: out {rel_linf(out.detach().cpu(), t(g['out'])):.2e} loss {abs(float(loss) - float(g['loss'])) / abs(float(g['loss'])):.2e} "
          f"grads {max(errs.values()):.2e} ({max(errs, key=errs.get)})")
    assert rel_linf(out.detach().cpu(), t(g["out"])) <= 1e-5          # the north-star bound for the fp32 / 3xTF32 path
    assert abs(float(loss) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    for k in errs:
        assert errs[k] <= 5e-4, k


def test_train_step_reduces_loss(cuda_device):
    """A few Adam steps on the Burgers workload through the fused path: the loss goes down."""
    from position_induced_transformer_b200 import workloads
    w = workloads.make_burgers(batch=4).to(cuda_device)
    gen = torch.Generator().manual_seed(0)
    ins, target = w.make_batch(gen, 4)
    ins = tuple(x.to(cuda_device) for x in ins)
    target = target.to(cuda_device)
    opt = torch.optim.Adam(w.model.parameters(), lr=1e-3)
    losses = []
    for _ in range(8):
        opt.zero_grad()
        loss = w.loss(target, workloads.run_model(w, ins))
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0]


def test_graphed_step_matches_eager(cuda_device):
    """Replaying the captured CUDA graph gives the same losses and parameters as the eager step."""
    import copy
    from position_induced_transformer_b200 import workloads
    from position_induced_transformer_b200.data_parallel import FlatGradients
    from position_induced_transformer_b200.graphed import GraphedTrainStep
    gen = torch.Generator().manual_seed(5)
    w = workloads.make_darcy(43, batch=2).to(cuda_device)
    batches = [w.make_batch(gen, 2) for _ in range(3)]
    batches = [(tuple(x.to(cuda_device) for x in ins), tgt.to(cuda_device)) for ins, tgt in batches]
    ref_model = copy.deepcopy(w.model)
    ref_model.mesh_ltt = w.model.mesh_ltt

    def loss_of(model):
        return lambda ins, tgt: w.loss(tgt, model(w.meshes[0], ins[0], w.meshes[0]))

    # plain SGD: Adam's sign-like first steps would amplify the 1e-7 summation-order noise of the atomics
    opt_g = torch.optim.SGD(w.model.parameters(), lr=1e-5)
    # capture needs an eager warm-up (cuBLAS handles, allocator pools): two steps on the first batch ...
    step = GraphedTrainStep(list(w.model.parameters()), loss_of(w.model), opt_g, batches[0][0], batches[0][1], warmup=2)
    opt_e = torch.optim.SGD(ref_model.parameters(), lr=1e-5)
    flat = FlatGradients(ref_model.parameters(), 1)

    def eager_step(ins, tgt):
        flat.zero()
        loss = loss_of(ref_model)(ins, tgt)
        loss.backward()
        opt_e.step()
        return loss

    for _ in range(2):                       # ... which the eager twin takes as well
        eager_step(*batches[0])
    for ins, tgt in batches:
        loss_e = eager_step(ins, tgt)
        loss_g = step(ins, tgt)
        assert abs(float(loss_g) - float(loss_e)) <= 1e-4 * abs(float(loss_e))
    for a, b in zip(w.model.parameters(), ref_model.parameters()):
        assert torch.allclose(a, b, rtol=1e-3, atol=1e-5)


def test_torch_compile_wraps_the_model(cuda_device):
    """The scripts call torch.compile(model) (train_darcy.py:112): the fused ops are opaque to Dynamo and results match eager."""
    from position_induced_transformer_b200 import workloads
    gen = torch.Generator().manual_seed(9)
    w = workloads.make_darcy(43, batch=2).to(cuda_device)
    (coeff,), target = w.make_batch(gen, 2)
    coeff, target = coeff.to(cuda_device), target.to(cuda_device)
    mesh = w.meshes[0]
    eager = w.model(mesh, coeff, mesh)
    compiled = torch.compile(w.model)
    out = compiled(mesh, coeff, mesh)
    loss = w.loss(target, out)
    loss.backward()
    assert rel_linf(out.detach(), eager.detach()) <= 1e-3       # Inductor may reassociate the TF32/fp32 MLP arithmetic
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in w.model.parameters())
