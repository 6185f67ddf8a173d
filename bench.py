#!/usr/bin/env python
"""PiT train-step benchmark (BASELINE.json metric: PiT train samples/s, fwd+bwd, at 1/2/4/8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload darcy421] [--impl ours|reference]

One "step" is one training step of the workload's PiT model on one synthetic batch per GPU:
zero_grad, forward, RelLp loss, backward, gradient all-reduce(SUM) over NCCL when N > 1, Adam update.
Rank 0 prints ONE JSON line.  See DESIGN.md section "Measurement" for every field.

  value   samples/s, whole job, inputs resident in HBM, timed on the device with CUDA events (max over ranks)
  e2e     same step through the public module API with HOST (pinned) inputs: H2D copy of the batch and a D2H
          read of the loss inside the timed region, every step
  roofline  the dominant position-attention kernel, timed live with CUDA events on its stream
  cpu_baseline  the CPU oracle (restatement of the reference's dense algorithm) on this box's host cores
--impl reference times that CPU oracle as the reference arm (the reference is a Python/PyTorch CPU/eager code
path; /root/reference itself does not exist on the GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "pit_train_samples_per_s"
UNIT = "samples/s"
STEP_DESC = "zero_grad+forward+RelLp loss+backward+grad allreduce(SUM)+Adam"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default="darcy421")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (0 = the reference script's batch size)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample-batch", type=int, default=0, help="samples per CPU-baseline step (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a captured CUDA graph")
    ap.add_argument("--precision", default="high", choices=["high", "highest"],
                    help="torch matmul precision for the MLP Linears (reference pit.py:2 sets 'high')")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# algorithmic bytes of one position-attention launch
# ----------------------------------------------------------------------------------------------
def algorithmic_bytes(key) -> int:
    """Q = 4*[B*M*D + B*N*H*D (+2*B*N*D concat) + sd*(N+M)*(B or 1)] for a forward stage; the
    backward re-reads U, reads dO (and O) and writes dU.  For the fused decoder tail the figure is that of the unfused
    stage it replaces (attention output of H*D floats per point), so `traffic` far below it is the effect of the fusion."""
    tag, _variant, batched, B, H, N, M, D, sd, concat = key
    meshes = sd * (N + M) * (B if batched else 1)
    values, outs = B * M * D, B * N * H * D
    if tag in ("fwd", "tail_fwd"):
        words = values + outs + meshes + (2 * B * N * D if concat else 0)
    elif tag == "bwd":             # reads U and dO once (fused scale + value gradients), writes dU
        words = 2 * values + outs + meshes + (B * N * D if concat else 0)
    elif tag == "tail_bwd":        # unfused: read dO and O (recomputed here), re-read U, write dU
        words = 2 * values + 2 * outs + meshes
    else:                          # rowstat: coordinates in, three floats per row out
        words = meshes + 3 * N * (B if batched else 1)
    return 4 * words


def measured_traffic(key):
    """DRAM bytes per launch of a position-attention call from the committed ncu captures (profiles/traffic.json)."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    table = json.load(open(path))
    tag, _variant, _batched, B, H, N, M, D, _sd, _concat = key
    return table.get(f"{tag}:B{B}:H{H}:N{N}:M{M}:D{D}")


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
# CPU oracle arm (reference's dense algorithm on host cores)
# ----------------------------------------------------------------------------------------------
def cpu_oracle_steps(workload_name: str, sample_batch: int, steps: int, warmup: int):
    """Train steps of the CPU oracle on a bounded sample; returns (samples/s, seconds per step, cores)."""
    from oracle import pit_oracle
    from position_induced_transformer_b200 import workloads
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    w = workloads.WORKLOADS[workload_name](sample_batch)
    params = {k: v.detach().clone().requires_grad_(True) for k, v in w.model.state_dict().items()}
    opt = torch.optim.Adam(list(params.values()), lr=1e-3)
    gen = torch.Generator().manual_seed(1234)
    ins, target = w.make_batch(gen, sample_batch)
    out_dim = w.model.out_dim
    loss_p = w.loss._ord
    mesh_ltt = w.model.mesh_ltt

    def step():
        opt.zero_grad()
        if w.meshes:
            variant = {"BurgersPiT": "periodic1d", "VorticityPiT": "periodic2d"}.get(type(w.model).__name__, "euclid")
            out = pit_oracle.forward_shared_mesh(params, variant, w.meshes[0], ins[0], mesh_ltt, w.meshes[0], w.model.en_local, w.model.de_local)
        else:
            mesh_in, func_in, mesh_out = ins
            if type(w.model).__name__ == "NacaPiT":
                b = mesh_out.shape[0]
                ltt = mesh_out[:, ::w.model.x_down, ::w.model.y_down, :].reshape(b, -1, 2)
                mesh_out_flat = mesh_out.reshape(b, -1, 2)
            else:
                ltt, mesh_out_flat = mesh_out, mesh_out
            out = pit_oracle.forward_point_cloud(params, mesh_in, func_in, ltt, mesh_out_flat, w.model.en_local, w.model.de_local)
        loss = pit_oracle.rel_lp_loss(target, out, out_dim, loss_p)
        loss.backward()
        opt.step()
        return float(loss)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return sample_batch / dt, dt, cores


def auto_cpu_sample(workload_name: str, batch: int) -> int:
    # the dense CPU path takes ~0.7 s per sample at Darcy-421 / NACA on 8 cores: keep the sample small there
    return {"darcy421": 2, "naca": 2, "elasticity": 2}.get(workload_name, batch)


def run_reference(args, rank, world):
    if rank != 0:
        return
    batch = args.batch or default_batch(args.workload)
    sample = args.cpu_sample_batch or auto_cpu_sample(args.workload, batch)
    steps, warmup = max(1, min(args.steps, 10)), max(1, min(args.warmup, 3))
    t0 = time.perf_counter()
    value, dt, cores = cpu_oracle_steps(args.workload, sample, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": bench_config(args, batch),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{steps} train steps of {sample} samples each (of the {batch}-sample batch) through oracle/pit_oracle.py, "
                                   f"{warmup} warm-up; torch {torch.__version__} CPU, {cores} threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
        "note": "reference arm = CPU restatement of the reference's dense PyTorch path (the Python reference cannot travel to the GPU box)",
    }
    print(json.dumps(line), flush=True)


def default_batch(name: str) -> int:
    return {"elasticity": 10, "naca": 20}.get(name, 8)


def bench_config(args, batch):
    return {"workload": args.workload, "per_gpu_batch": batch, "global_batch": batch * args.gpus, "step": STEP_DESC,
            "launch": "eager" if getattr(args, "no_graph", False) else "cuda_graph_replay",
            "parallelism": f"dp{args.gpus}", "mlp_matmul_precision": args.precision,
            "l2": "per-step working set (activations of the decoder stage) exceeds the 126 MB L2 and input batches rotate over 4 buffers; no explicit flush"}


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from position_induced_transformer_b200 import _cabi, posatt, workloads
    from position_induced_transformer_b200.data_parallel import FlatGradients
    from position_induced_transformer_b200.graphed import GraphedTrainStep

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU path)"
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    torch.set_float32_matmul_precision(args.precision)
    batch = args.batch or default_batch(args.workload)
    w = workloads.WORKLOADS[args.workload](batch).to(dev)
    model = w.model
    use_graph = not args.no_graph
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=use_graph, fused=True)   # torch's single-kernel Adam
    gen = torch.Generator().manual_seed(1234 + rank)
    n_buf = 4
    host = [w.make_batch(gen, batch) for _ in range(n_buf)]
    host = [(tuple(x.pin_memory() for x in ins), tgt.pin_memory()) for ins, tgt in host]
    resident = [(tuple(x.to(dev) for x in ins), tgt.to(dev)) for ins, tgt in host]
    h2d_bytes = sum(x.numel() * x.element_size() for x in host[0][0]) + host[0][1].numel() * 4

    def forward_loss(ins, target):
        return w.loss(target, workloads.run_model(w, ins))

    if use_graph:
        # the whole step (zero grads, forward, loss, backward, all-reduce, Adam) is captured once and replayed
        step = GraphedTrainStep(list(model.parameters()), forward_loss, opt, resident[0][0], resident[0][1], world)
    else:
        flat = FlatGradients(model.parameters(), world)

        def step(ins, target):
            flat.release()
            loss = forward_loss(ins, target)
            loss.backward()
            if world > 1:
                flat.gather()
                flat.all_reduce()
            opt.step()
            return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident measurement ----
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()          # sampled from the warm-up on (20 ms period): the timed region itself can be < 100 ms
    for i in range(args.warmup):
        step(*resident[i % n_buf])
    barrier()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for i in range(args.steps):
        step(*resident[i % n_buf])
    end.record()
    barrier()
    clock_info = clocks.stop() if rank == 0 else None
    ms = start.elapsed_time(end)

    # ---- per-kernel timing: the same step launched eagerly with CUDA events around every C-ABI call ----
    eager = step._eager if use_graph else (lambda: step(*resident[0]))
    eager()
    barrier()
    launches0 = _cabi.launch_count()
    timer = posatt.KernelTimer()
    posatt.set_kernel_timer(timer)
    k_steps = min(args.steps, 5)
    k_start, k_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k_start.record()
    for _ in range(k_steps):
        eager()
    k_end.record()
    barrier()
    posatt.set_kernel_timer(None)
    eager_ms = k_start.elapsed_time(k_end) / k_steps
    launches = (_cabi.launch_count() - launches0) // k_steps * args.steps   # the graph replays exactly these launches
    kernels = timer.summary()

    # ---- end to end: host inputs, H2D every step, loss read back every step ----
    # The H2D copy of batch i+1 runs on a copy stream while step i computes: it lands in one of two staging tensors, which
    # step i+1 copies into the graph's static inputs as its first action (the static inputs are read by forward AND
    # backward, so they cannot be overwritten mid-step).  A staging slot is free again as soon as that device-to-device copy
    # has run -- not when the whole step has finished.  Every step still waits for its own inputs and reads its own loss back,
    # all inside the timed region.
    copy_stream = torch.cuda.Stream()
    stage = [(tuple(torch.empty_like(x, device=dev) for x in host[0][0]), torch.empty_like(host[0][1], device=dev)) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def prefetch(i):
        ins, tgt = host[i % n_buf]
        slot = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])                 # the step that last read this slot has taken its copy
            for dst, src in zip(stage[slot][0], ins):
                dst.copy_(src, non_blocking=True)
            stage[slot][1].copy_(tgt, non_blocking=True)
            ready[slot].record(copy_stream)

    def e2e_loop(n):
        main = torch.cuda.current_stream()
        prefetch(0)
        for i in range(n):
            slot = i % 2
            main.wait_event(ready[slot])
            if use_graph:
                step.load(stage[slot][0], stage[slot][1])
                consumed[slot].record(main)
                loss = step.replay()
            else:
                loss = step(stage[slot][0], stage[slot][1])         # eager: the staging tensors are the step's inputs until it ends
                consumed[slot].record(main)
            if i + 1 < n:
                prefetch(i + 1)
            float(loss)                                            # D2H read of the step's result

    e2e_loop(2)
    barrier()
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record()
    e2e_loop(args.steps)
    e_end.record()
    barrier()
    e2e_ms = e_start.elapsed_time(e_end)

    # ---- forward only (inference): the same batches through model + loss under no_grad, replayed as a graph ----
    fwd_ms = None
    if use_graph:
        static_in = tuple(x.clone() for x in resident[0][0])
        static_tgt = resident[0][1].clone()
        with torch.no_grad():
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    forward_loss(static_in, static_tgt)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            fgraph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(fgraph):
                floss = forward_loss(static_in, static_tgt)

            def fwd_step(ins, target):
                for dst, src in zip(static_in, ins):
                    dst.copy_(src, non_blocking=True)
                static_tgt.copy_(target, non_blocking=True)
                fgraph.replay()
                return floss

            for i in range(min(args.warmup, 5)):
                fwd_step(*resident[i % n_buf])
            barrier()
            f_start, f_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f_start.record()
            for i in range(args.steps):
                fwd_step(*resident[i % n_buf])
            f_end.record()
            barrier()
            fwd_ms = f_start.elapsed_time(f_end)

    if world > 1:
        tms = torch.tensor([ms, e2e_ms, fwd_ms or 0.0], device=dev, dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(tms[0]), float(tms[1])
        fwd_ms = float(tms[2]) if fwd_ms is not None else None
    if rank != 0:
        return

    total = batch * world * args.steps
    value = total / (ms * 1e-3)
    peak, peak_src = load_peaks()
    roofline = None
    if kernels:
        top = max((k for k in kernels if k[0] != "rowstat"), key=lambda k: kernels[k]["ms_total"], default=None)
        if top is not None:
            q = algorithmic_bytes(top)
            ach = q / (kernels[top]["ms_avg"] * 1e-3) / 1e9
            share = kernels[top]["ms_avg"] * kernels[top]["calls"] / k_steps / (ms / args.steps)
            roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": measured_traffic(top),
                        "kernel": {"call": top[0], "variant": top[1], "mesh_batched": bool(top[2]), "B": top[3], "H": top[4],
                                   "N": top[5], "M": top[6], "D": top[7]},
                        "algorithmic_bytes_per_launch": q, "avg_launch_ms": kernels[top]["ms_avg"],
                        "share_of_step": share, "peak_source": peak_src,
                        "note": ("fused decoder tail: `achieved` counts the bytes the unfused attention stage has to move (SURVEY 8d); the kernel "
                                 "itself keeps them on chip (see `traffic`) and is bound by instruction issue, profiles/r1_v9_ncu_tail_mma.md")
                        if top[0].startswith("tail") else None}
    kernel_table = sorted(({"call": k[0], "N": k[5], "M": k[6], "D": k[7], "concat": bool(k[9]), "ms_avg": v["ms_avg"],
                            "calls_per_step": v["calls"] / k_steps, "share_of_step": v["ms_total"] / k_steps / (ms / args.steps),
                            "GBps_algorithmic": algorithmic_bytes(k) / (v["ms_avg"] * 1e-3) / 1e9} for k, v in kernels.items()),
                          key=lambda r: -r["share_of_step"])

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        sample = args.cpu_sample_batch or auto_cpu_sample(args.workload, batch)
        v, dt, cores = cpu_oracle_steps(args.workload, sample, 3, 1)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"3 train steps of {sample} samples (of the {batch}-sample batch) through oracle/pit_oracle.py after 1 warm-up; {dt:.2f} s/step"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.precision == "highest" else "f32 (posatt) + tf32 (MLP Linears, as reference pit.py:2)",
        "data": "synthetic", "config": bench_config(args, batch),
        "e2e": {"value": total / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_ms / args.steps},
        "forward_only": None if fwd_ms is None else {"value": batch * world * args.steps / (fwd_ms * 1e-3), "unit": UNIT,
                                                     "ms_per_step": fwd_ms / args.steps, "what": "forward + loss under no_grad, device-resident inputs"},
        "gpu_launches": launches, "eager_ms_per_step": eager_ms, "roofline": roofline, "cpu_baseline": cpu, "clocks": clock_info, "kernels": kernel_table[:8],
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    elif args.gpus > 1:
        sys.exit("launch with torch.distributed.run for --gpus > 1")
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
