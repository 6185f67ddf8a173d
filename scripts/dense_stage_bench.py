"""Time the global self stage at the latent-grid shape: CUDA-graph replay of forward+backward, so that launch overhead of
the eager path does not blur the number (PIT_DENSE_TCGEN05=0 selects the SIMT kernels for comparison).   python scripts/dense_ab.py [N=256] [B=8] [D=64] [H=2]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from position_induced_transformer_b200 import posatt  # noqa: E402


def main():
    n, b, d, h = (int(x) for x in (sys.argv[1:5] + ["256", "8", "64", "2"][len(sys.argv) - 1:]))
    dev = torch.device("cuda:0")
    side = int(n ** 0.5)
    ax = torch.linspace(0, 1, side)
    mesh = torch.stack(torch.meshgrid(ax, ax, indexing="ij"), -1).reshape(-1, 2).to(dev)
    u = torch.randn(b, mesh.shape[0], d, device=dev, requires_grad=True)
    scale = (torch.rand(h, device=dev) * 5 + 1).requires_grad_(True)
    up = torch.randn(b, mesh.shape[0], (1 + h) * d, device=dev)

    def fb():
        out = posatt.position_attention(mesh, mesh, u, scale, 1.0, "euclid", True)
        out.backward(up)
        return out

    for _ in range(3):
        fb()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            fb()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    print(f"N=M={mesh.shape[0]} B={b} D={d} H={h}: fwd+bwd {a.elapsed_time(e) / 200 * 1e3:.1f} us per stage")


if __name__ == "__main__":
    main()
