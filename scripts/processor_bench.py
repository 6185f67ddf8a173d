"""Time the fused processor (forward / backward C-ABI calls) with CUDA events, against the per-block path.

    python scripts/processor_bench.py [workload=darcy421] [batch=8] [precision=high]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from position_induced_transformer_b200 import posatt, workloads  # noqa: E402
import position_induced_transformer_b200.pit as pit_mod  # noqa: E402


def run(model, latent, fused, iters=20):
    pit_mod.use_fused_processor(fused)
    up = torch.ones_like(latent)
    params = [p for n, p in model.named_parameters() if n.startswith(("conv.", "mlp."))]
    latent = latent.detach()

    def step():
        for p in params:
            p.grad = None
        x = latent.clone().requires_grad_(True)      # a leaf born on the capturing stream
        out = model.processor(x, model.mesh_ltt)
        out.backward(up)
        return out, None

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out, grads = step()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.replay()
    s.record()
    for _ in range(iters):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3, out.detach().clone()


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "darcy421"
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    torch.set_float32_matmul_precision(sys.argv[3] if len(sys.argv) > 3 else "high")
    dev = torch.device("cuda:0")
    w = workloads.make(name, batch).to(dev)
    model = w.model
    latent = torch.randn(batch, model.mesh_ltt.shape[0], model.hid_dim, device=dev, requires_grad=True)
    timer = posatt.KernelTimer()
    posatt.set_kernel_timer(timer)
    for _ in range(5):
        out = model.processor(latent, model.mesh_ltt)
        out.backward(torch.ones_like(out))
    posatt.set_kernel_timer(None)
    for key, v in timer.summary().items():
        print(key[0], f"N={key[5]} D={key[7]} H={key[4]} blocks={key[9]}", f"{v['ms_avg'] * 1e3:.1f} us  x{v['calls']} (eager, event-timed)")
    t_fused, a = run(model, latent, True)
    t_block, b = run(model, latent, False)
    print(f"{name} B={batch}: processor fwd+bwd in a replayed graph: fused {t_fused:.1f} us, per-block path {t_block:.1f} us; "
          f"rel-Linf between them {float((a - b).abs().max() / b.abs().max()):.2e}")


if __name__ == "__main__":
    with torch.cuda.stream(torch.cuda.Stream()):     # nothing touches the legacy stream: gradients can be captured later
        main()
