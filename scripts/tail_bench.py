"""Time the fused decoder tail (forward / backward C-ABI calls) at the Darcy-421 shape with CUDA events.

    python scripts/tail_bench.py [n_side=421] [batch=8]
PIT_TAIL_MMA=0 selects the SIMT gather kernels for comparison.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from position_induced_transformer_b200 import posatt, workloads  # noqa: E402
import position_induced_transformer_b200.pit as pit_mod  # noqa: E402


def main():
    side = int(sys.argv[1]) if len(sys.argv) > 1 else 421
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    dev = torch.device("cuda:0")
    w = workloads.make_darcy(side, batch=batch).to(dev)
    model = w.model
    mesh = w.meshes[0].reshape(-1, 2)
    latent = torch.randn(batch, model.mesh_ltt.shape[0], model.hid_dim, device=dev, requires_grad=True)
    timer = posatt.KernelTimer()
    for it in range(8):
        if it == 3:
            posatt.set_kernel_timer(timer)
        out = model.decoder(model.mesh_ltt, latent, mesh)
        out.backward(torch.ones_like(out))
    posatt.set_kernel_timer(None)
    for key, v in timer.summary().items():
        print(key[0], f"N={key[5]} M={key[6]} C={key[7]}", f"{v['ms_avg'] * 1e3:.1f} us  x{v['calls']}")
    pit_mod.use_fused_decoder_tail(False)
    ref = model.decoder(model.mesh_ltt, latent, mesh)
    print("fused vs two-module rel-Linf:", float((out - ref).abs().max() / ref.abs().max()))


if __name__ == "__main__":
    main()
