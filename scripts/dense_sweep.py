"""Throughput of the global (dense) position-attention stage on the tcgen05 kernel: forward and backward.

    python scripts/dense_sweep.py [--quick]

Prints one JSON line per shape: algorithmic TFLOP/s (2*H*N*M*B*D per contraction; forward = 1 contraction,
backward = 2: dU and the scale gradient) and issued TF32 TFLOP/s (x3, 3xTF32 split; the scale gradient issues two
products per tile).  Timing: CUDA events, 3 warm-ups, median of 10, inputs far larger than L2 are not needed here
because the value block is re-read from L2 by design.
"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from position_induced_transformer_b200.posatt import position_attention, KernelTimer, set_dense_precision, set_kernel_timer  # noqa: E402


def bench(B, N, D, H, batched, reps=10, precision="fp32"):
    set_dense_precision(precision)
    terms = 3 if precision == "fp32" else 1
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    mesh = torch.rand((B, N, 2) if batched else (N, 2), generator=g).to(dev)
    vals = torch.randn(B, N, D, generator=g).to(dev).requires_grad_(True)
    scale = (torch.rand(H, generator=g) * 20 + 5).to(dev).requires_grad_(True)
    up = torch.randn(B, N, (1 + H) * D, generator=g).to(dev)
    for _ in range(3):
        out = position_attention(mesh, mesh, vals, scale, 1.0, "euclid", True)
        out.backward(up)
    torch.cuda.synchronize()
    timer = KernelTimer()
    set_kernel_timer(timer)
    for _ in range(reps):
        out = position_attention(mesh, mesh, vals, scale, 1.0, "euclid", True)
        out.backward(up)
        vals.grad = None
        scale.grad = None
    set_kernel_timer(None)
    summ = timer.summary()
    fwd = [v for k, v in summ.items() if k[0] == "fwd"][0]["ms_avg"]
    bwd = [v for k, v in summ.items() if k[0] == "bwd"][0]["ms_avg"]
    flops = 2.0 * H * N * N * B * D
    set_dense_precision("fp32")
    return {"B": B, "N": N, "D": D, "H": H, "mesh_batched": batched, "operands": precision, "fwd_ms": fwd, "bwd_ms": bwd,
            "fwd_TFLOPs_algorithmic": flops / fwd / 1e9, "fwd_TFLOPs_issued_tf32": terms * flops / fwd / 1e9,
            "bwd_TFLOPs_algorithmic": 2 * flops / bwd / 1e9, "bwd_TFLOPs_issued_tf32": 3 * terms * flops / bwd / 1e9}


def matmul_peak(dtype, tf32):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(8192, 8192, device="cuda", dtype=dtype)
    b = torch.randn(8192, 8192, device="cuda", dtype=dtype)
    for _ in range(3):
        a @ b
    best = 1e9
    for _ in range(10):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        a @ b
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return 2 * 8192 ** 3 / best / 1e9


def main():
    quick = "--quick" in sys.argv
    print(json.dumps({"cublas_tf32_8192_TFLOPs": matmul_peak(torch.float32, True), "cublas_bf16_8192_TFLOPs": matmul_peak(torch.bfloat16, False),
                      "cublas_fp32_8192_TFLOPs": matmul_peak(torch.float32, False)}), flush=True)
    shapes = [
        (8, 256, 64, 2, False),      # Darcy / Burgers processor
        (10, 972, 256, 2, True),     # elasticity processor
        (20, 728, 128, 1, True),     # NACA processor
        (200, 896, 256, 1, False),   # cylinder processor
    ]
    if not quick:
        shapes += [(8, 1024, 64, 2, False), (16, 1024, 256, 1, False), (16, 4096, 256, 1, False), (8, 4096, 128, 2, True)]
    import os
    if os.environ.get("DENSE_SHAPE"):          # profiling aid: one shape, one precision (DENSE_PREC)
        print(json.dumps(bench(*shapes[int(os.environ["DENSE_SHAPE"])], reps=3, precision=os.environ.get("DENSE_PREC", "fp32"))), flush=True)
        return
    for s in shapes:
        for precision in ("fp32", "tf32", "bf16"):
            if s[1] <= 256 and precision != "fp32":
                continue        # small shared latents run the fused processor kernel in the models, not this one
            print(json.dumps(bench(*s, precision=precision)), flush=True)


if __name__ == "__main__":
    main()
