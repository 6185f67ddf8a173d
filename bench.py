#!/usr/bin/env python
"""PiT train-step benchmark (BASELINE.json metric: PiT train samples/s, fwd+bwd, at 1/2/4/8 B200; posatt TFLOP/s vs peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload darcy421] [--impl ours|reference]

One "step" is one training step of the workload's PiT model on one synthetic batch per GPU: zero_grad, forward,
RelLp loss, backward, gradient all-reduce(SUM) over NCCL when N > 1, Adam update.  Rank 0 prints ONE JSON line; see
DESIGN.md section "Measurement" for every field.

  value        samples/s, whole job, inputs resident in HBM, CUDA events, max over ranks.  K steps are timed per repeat;
               the repeat count R is raised until the timed region lasts >= 0.5 s (`timed_steps` = K*R)
  e2e          same step through the public module API with HOST (pinned) inputs: H2D copy of the batch and a D2H
               read of the loss inside the timed region, every step
  roofline     the dominant position-attention call, timed live with CUDA events on its stream: HBM fraction on the bytes
               the kernel has to move, tensor fraction on the TF32 MMA flops it issues against a TF32 GEMM timed in the
               same run, and -- for a fused stage -- the figure of the unfused stage it replaces under `unfused_equivalent`
  workloads    the same value / e2e / dominant kernel / dense-stage TFLOP/s for all five BASELINE configurations
  cpu_baseline the CPU oracle (restatement of the reference's dense algorithm) on this box's host cores (N = 1 only)
  gpu_eager_reference  context: the unmodified reference modules (baseline/_ref, TF32 'high' as shipped, eager) on the same GPU

--impl reference times the CPU oracle as the reference arm (the reference is a CPU/eager PyTorch path; /root/reference
does not exist on the GPU box).  That arm never imports the CUDA library.
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

OPTIMIZER_USED = {}
METRIC = "pit_train_samples_per_s"
UNIT = "samples/s"
STEP_DESC = "zero_grad+forward+RelLp loss+backward+grad allreduce(SUM)+Adam"
MIN_TIMED_S = 0.5


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
    ap.add_argument("--no-sweep", action="store_true", help="skip the other BASELINE workloads (the `workloads` dict then holds the primary one only)")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip the eager run of baseline/_ref on the GPU")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a captured CUDA graph")
    ap.add_argument("--optimizer", default="auto", choices=["auto", "torch", "fused"],
                    help="torch: torch.optim.Adam (fused kernel) after one NCCL all-reduce of the flat gradient; fused: the cross-rank SUM "
                         "over NVLink peer memory and Adam in ONE launch (pit_allreduce_adam); auto: fused when N > 1")
    ap.add_argument("--precision", default="high", choices=["high", "highest"],
                    help="torch matmul precision for the MLP Linears (reference pit.py:2 sets 'high')")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampler (100 ms period) read by a thread; `wait_first()` blocks until the first row has arrived so that a
    timed region never starts before the sampler is live."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.rows, self.thread = index, None, [], None
        self.first = threading.Event()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True, bufsize=1)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 9:
                self.rows.append((time.perf_counter(), f))
                self.first.set()

    def wait_first(self, timeout=8.0):
        return self.proc is not None and self.first.wait(timeout)

    def window(self, t0, t1):
        """Summary of the rows sampled in [t0, t1] (perf_counter times); falls back to all rows if the window caught none."""
        rows = [f for t, f in self.rows if t0 <= t <= t1] or [f for _, f in self.rows]
        sm, mx, pw, reasons = [], [], [], set()
        for f in rows:
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples" if self.proc else "nvidia-smi unavailable"], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw) if pw else None,
                "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()


# ----------------------------------------------------------------------------------------------
# algorithmic work of one position-attention call
# ----------------------------------------------------------------------------------------------
def unfused_bytes(key) -> int:
    """Q = 4*[B*M*D + B*N*H*D (+2*B*N*D concat) + sd*(N+M)*(B or 1)] for a forward stage; the
    backward re-reads U, reads dO (and O) and writes dU.  For the fused decoder tail this is the figure of the UNFUSED
    attention stage it replaces (H*D floats per point written and read back)."""
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


def tail_work(key, plan, out_dim):
    """Bytes the fused decoder tail has to move and TF32 MMA flops it issues (csrc/decoder_tail_plan.cuh), from the plan.

    bytes  forward: Y + plan (records, candidate lists, distances) + out + saved row sums; backward: the same reads + d_out
           + d_y (accumulated with REDs: counted twice)
    flops  one mma.sync.m16n8k8 = 2048 flop.  Forward, per tile and k-step of 8 candidates: H heads x (B*C/16) m-tiles x 4
           n-tiles x 3 split terms; the backward repeats that for the pre-activation and issues twice as many again for
           dY and dZ (k = the 32 tile rows: 4 k-steps x 2 products x 3 split terms per head, m-tile and group of 8)."""
    tag, _variant, _batched, B, H, N, M, C, _sd, _ = key
    ksteps = int(((plan.tile_cnt + 7) // 8).sum().item())
    n_tiles, n_cand = plan.n_tiles, plan.n_cand
    plan_bytes = n_tiles * 32 * 16 + n_cand * (128 + 2) + n_tiles * 8
    y_bytes, out_bytes, rs_bytes = B * M * H * C * 4, B * N * out_dim * 4, 2 * H * n_tiles * 32 * 4
    fwd_mma = ksteps * H * (B * C // 16) * 4 * 3
    if tag == "tail_fwd":
        return y_bytes + plan_bytes + out_bytes + rs_bytes, fwd_mma * 2048, ksteps
    return y_bytes + plan_bytes + out_bytes + rs_bytes + 2 * y_bytes, 3 * fwd_mma * 2048, ksteps


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measure_tf32_peak(dev):
    """cuBLAS TF32 GEMM rate on this GPU, in this run: 8192^3, best of 5 after a warm-up (TFLOP/s)."""
    prev = torch.get_float32_matmul_precision()
    torch.set_float32_matmul_precision("high")
    n = 8192
    a = torch.randn(n, n, device=dev)
    b = torch.randn(n, n, device=dev)
    torch.matmul(a, b)
    best = float("inf")
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        torch.matmul(a, b)
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    torch.set_float32_matmul_precision(prev)
    del a, b
    return 2 * n ** 3 / (best * 1e-3) / 1e12


# ----------------------------------------------------------------------------------------------
# CPU oracle arm (reference's dense algorithm on host cores) -- no import of the CUDA library anywhere below
# ----------------------------------------------------------------------------------------------
def cpu_oracle_steps(workload_name: str, sample_batch: int, steps: int, warmup: int):
    """Train steps of the CPU oracle on a bounded sample; returns (samples/s, seconds per step, cores)."""
    from oracle import pit_oracle
    from position_induced_transformer_b200 import workload_specs       # pure data: does not load libpit_posatt.so
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    spec = workload_specs.SPECS[workload_name]()
    params = {k: v.requires_grad_(True) for k, v in pit_oracle.init_params(spec).items()}
    opt = torch.optim.Adam(list(params.values()), lr=1e-3)
    ins, target = spec.make_batch(torch.Generator().manual_seed(1234), sample_batch)

    def step():
        opt.zero_grad()
        loss = pit_oracle.step_loss(params, spec, ins, target)
        loss.backward()
        opt.step()
        return float(loss.detach())

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return sample_batch / dt, dt, cores


def auto_cpu_sample(workload_name: str, batch: int) -> int:
    # the dense CPU path takes ~0.7 s per sample at Darcy-421 / NACA on 8 cores: keep the sample small there
    return {"darcy421": 2, "naca": 2, "elasticity": 2, "cylinder": 4, "vorticity": 1}.get(workload_name, batch)


def default_batch(name: str) -> int:
    return {"elasticity": 10, "naca": 20, "vorticity": 20, "cylinder": 200}.get(name, 8)


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
        "config": {"workload": args.workload, "per_gpu_batch": batch, "sample_batch": sample, "step": STEP_DESC, "launch": "cpu_eager",
                   "parallelism": "one CPU process (rank 0) whatever --gpus says: the reference has no multi-GPU path",
                   "mlp_matmul_precision": "fp32 (CPU)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{steps} train steps of {sample} samples each (of the {batch}-sample batch) through oracle/pit_oracle.py, "
                                   f"{warmup} warm-up; torch {torch.__version__} CPU, {cores} threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
        "note": "reference arm = CPU restatement of the reference's dense PyTorch path (the Python reference cannot travel to the GPU box); "
                "GPU-vs-CPU context, not a same-device comparison -- see `gpu_eager_reference` in the other arm's line for that",
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# the reference's own modules on the GPU, eager, as shipped (context; needs baseline/_ref)
# ----------------------------------------------------------------------------------------------
def gpu_eager_reference(workload_name, batch, dev, steps=5, warmup=2):
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.exists(os.path.join(ref_dir, "pit.py")):
        return {"unavailable": "baseline/_ref/pit.py not present (copied from /root/reference by __graft_entry__.build() in the build container)"}
    from position_induced_transformer_b200 import workload_specs
    spec = workload_specs.SPECS[workload_name]()
    if spec.family == "batched" or spec.rollout != 1 or spec.extra:
        return {"unavailable": f"no eager-reference wrapper for workload family {spec.family!r}"}
    prev = torch.get_float32_matmul_precision()
    rng = torch.get_rng_state()
    try:
        sp = importlib.util.spec_from_file_location("ref_pit_unmodified", os.path.join(ref_dir, "pit.py"))
        ref = importlib.util.module_from_spec(sp)
        sp.loader.exec_module(ref)          # sets matmul precision 'high' (TF32) and reseeds, as shipped (pit.py:2-6)
        base = {"fixed": ref.pit_fixed, "periodic1d": ref.pit_periodic1d, "periodic2d": ref.pit_periodic2d}[spec.family]

        class Model(base):                  # the forward every shared-mesh script defines (train_darcy.py:46-59)
            def forward(self, mesh_in, func_in, mesh_out):
                size = mesh_out.shape[:-1]
                mesh_in = mesh_in.reshape(-1, self.space_dim)
                func_in = func_in.reshape(func_in.shape[0], -1, self.in_dim)
                mesh_out = mesh_out.reshape(-1, self.space_dim)
                func_in = torch.cat((torch.tile(mesh_in.unsqueeze(0), [func_in.shape[0], 1, 1]), func_in), -1)
                h = self.encoder(mesh_in, func_in, self.mesh_ltt)
                h = self.processor(h, self.mesh_ltt)
                return self.decoder(self.mesh_ltt, h, mesh_out).reshape(func_in.shape[0], *size, self.out_dim)

        sd, in_dim, out_dim, hid, heads, blocks, en_loc, de_loc = spec.ctor
        model = Model(sd, in_dim, out_dim, hid, heads, blocks, spec.mesh_ltt.to(dev), en_loc, de_loc).to(dev)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        mesh = spec.mesh.to(dev)
        ins, tgt = spec.make_batch(torch.Generator().manual_seed(99), batch)
        x, tgt = ins[0].to(dev), tgt.to(dev)
        out_dim_l, p = spec.loss

        def step():
            opt.zero_grad()
            out = model(mesh, x, mesh)
            t, q = tgt.reshape(batch, -1, out_dim_l), out.reshape(batch, -1, out_dim_l)
            loss = (torch.norm(t - q, p=p, dim=1) / torch.norm(t, p=p, dim=1)).mean(-1).sum()      # utils.py:86-98
            loss.backward()
            opt.step()

        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            step()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / steps
        return {"value": batch / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "warmup": warmup,
                "what": "unmodified reference pit.py (baseline/_ref) on this GPU: eager, TF32 matmuls as shipped (pit.py:2), no torch.compile, "
                        "torch.optim.Adam, inputs resident -- the step a user of the reference runs today",
                "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30}
    except Exception as exc:  # noqa: BLE001 -- context only: never fail the bench over it
        return {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}
    finally:
        torch.set_float32_matmul_precision(prev)
        torch.set_rng_state(rng)
        torch.cuda.empty_cache()


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def timed_loop(fn, steps, world, est_ms=None):
    """K steps per repeat, repeats until >= MIN_TIMED_S; returns (ms per step, timed steps, host window)."""
    if est_ms is None:
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn(min(steps, 10))
        e.record()
        barrier(world)
        est_ms = max(s.elapsed_time(e) / min(steps, 10), 1e-3)
    repeats = max(1, math.ceil(MIN_TIMED_S * 1e3 / (est_ms * steps)))
    if world > 1:      # every rank must run the same number of steps
        import torch.distributed as dist
        r = torch.tensor([repeats], device="cuda")
        dist.all_reduce(r, op=dist.ReduceOp.MAX)
        repeats = int(r.item())
    barrier(world)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    s.record()
    for _ in range(repeats):
        fn(steps)
    e.record()
    barrier(world)
    t1 = time.perf_counter()
    return s.elapsed_time(e) / (steps * repeats), steps * repeats, (t0, t1)


class PeerTimeout(RuntimeError):
    """The peer-memory optimizer step gave up waiting for a rank: every rank raises it together (the flag is all-reduced)."""


def measure_workload(name, args, rank, world, dev, detail, tf32_peak, hbm_peak):
    """measure_workload_once, falling back from the peer-memory optimizer step to NCCL all-reduce + torch Adam if a warm-up step
    ever timed out waiting for a peer (every rank takes the same decision, so the job stays in step)."""
    try:
        return measure_workload_once(name, args, rank, world, dev, detail, tf32_peak, hbm_peak, False)
    except PeerTimeout as ex:
        if rank == 0:
            print(f"[bench] {name}: {ex}; repeating with NCCL all-reduce + torch Adam", file=sys.stderr)
        return measure_workload_once(name, args, rank, world, dev, detail, tf32_peak, hbm_peak, True)


def measure_workload_once(name, args, rank, world, dev, detail, tf32_peak, hbm_peak, force_torch):
    """value / e2e (and, with `detail`, forward-only and the full kernel table) of one workload on this rank."""
    import torch.distributed as dist
    from position_induced_transformer_b200 import _cabi, posatt, workloads
    from position_induced_transformer_b200.data_parallel import FlatGradients
    from position_induced_transformer_b200.graphed import GraphedTrainStep

    batch = (args.batch if name == args.workload and args.batch else 0) or default_batch(name)
    posatt.mesh_cache.clear()
    w = workloads.WORKLOADS[name](batch).to(dev)
    model = w.model
    use_graph = not args.no_graph
    fused_opt = not force_torch and (args.optimizer == "fused" or (args.optimizer == "auto" and world > 1))
    OPTIMIZER_USED[name] = "torch"
    if fused_opt:
        from position_induced_transformer_b200.fused_optimizer import FusedAllReduceAdam
        try:
            opt = FusedAllReduceAdam(model.parameters(), lr=1e-3)   # cross-rank SUM over peer memory + Adam, one launch
            ok = torch.ones(1, device=dev)
        except Exception as ex:  # noqa: BLE001  (no peer access / symmetric memory on this box)
            if rank == 0:
                print(f"[bench] fused optimizer unavailable ({type(ex).__name__}: {ex}); using NCCL all-reduce + torch Adam", file=sys.stderr)
            ok = torch.zeros(1, device=dev)
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)               # all ranks take the same path
        fused_opt = bool(ok.item())
        OPTIMIZER_USED[name] = "fused" if fused_opt else "torch"
    if not fused_opt:
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=use_graph, fused=True)   # torch's single-kernel Adam
    gen = torch.Generator().manual_seed(1234 + rank)
    n_buf = 4
    host = [w.make_batch(gen, batch) for _ in range(n_buf)]
    host = [(tuple(x.pin_memory() for x in ins), tgt.pin_memory()) for ins, tgt in host]
    resident = [(tuple(x.to(dev) for x in ins), tgt.to(dev)) for ins, tgt in host]
    h2d_bytes = sum(x.numel() * x.element_size() for x in host[0][0]) + host[0][1].numel() * 4

    def forward_loss(ins, target):
        return workloads.step_loss(w, ins, target)

    if use_graph:
        # the whole step (zero grads, forward, loss, backward, all-reduce, Adam) is captured once and replayed
        step = GraphedTrainStep(list(model.parameters()), forward_loss, opt, resident[0][0], resident[0][1], world)
    else:
        flat = FlatGradients(model.parameters(), 1 if fused_opt else world)

        def step(ins, target):
            flat.release()
            loss = forward_loss(ins, target)
            loss.backward()
            if world > 1 and not fused_opt:
                flat.gather()
                flat.all_reduce()
            opt.step()
            return loss

    # ---- device-resident measurement ----
    counter = [0]

    def run_resident(n):
        for _ in range(n):
            step(*resident[counter[0] % n_buf])
            counter[0] += 1

    run_resident(args.warmup)
    barrier(world)
    if fused_opt:
        bad = torch.tensor([float(opt.peer_timeout())], device=dev)
        if world > 1:
            dist.all_reduce(bad, op=dist.ReduceOp.MAX)          # every rank takes the same decision
        if float(bad) > 0 or os.environ.get("PIT_BENCH_FAULT_PEER_TIMEOUT") == name:      # (the variable is a test hook)
            del step, opt
            torch.cuda.synchronize()
            raise PeerTimeout("FusedAllReduceAdam: a warm-up step gave up waiting for a peer rank (2 s)")
    ms, timed_steps, window = timed_loop(run_resident, args.steps, world)

    # ---- per-kernel timing: the same step launched eagerly with CUDA events around every C-ABI call ----
    eager = step._eager if use_graph else (lambda: step(*resident[0]))
    eager()
    barrier(world)
    launches0 = _cabi.launch_count()
    timer = posatt.KernelTimer()
    posatt.set_kernel_timer(timer)
    k_steps = 3
    k_start, k_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k_start.record()
    for _ in range(k_steps):
        eager()
    k_end.record()
    barrier(world)
    posatt.set_kernel_timer(None)
    eager_ms = k_start.elapsed_time(k_end) / k_steps
    launches_per_step = (_cabi.launch_count() - launches0) // k_steps      # the graph replays exactly these launches
    kernels = timer.summary()

    # ---- end to end: host inputs, H2D every step, loss read back every step ----
    # The H2D copy of batch i+1 runs on a copy stream while step i computes: it lands in one of two staging tensors, which
    # step i+1 copies into the graph's static inputs as its first action (the static inputs are read by forward AND
    # backward, so they cannot be overwritten mid-step).  A staging slot is free again as soon as that device-to-device copy
    # has run.  Every step still waits for its own inputs and reads its own loss back, all inside the timed region.
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
    barrier(world)
    e2e_ms, e2e_steps, _ = timed_loop(e2e_loop, args.steps, world, est_ms=ms)

    # ---- forward only (inference): the same batches through model + loss under no_grad, replayed as a graph ----
    fwd_ms = None
    if use_graph and detail:
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
                forward_loss(static_in, static_tgt)

            def fwd_steps(n):
                for _ in range(n):
                    ins, target = resident[counter[0] % n_buf]
                    counter[0] += 1
                    for dst, src in zip(static_in, ins):
                        dst.copy_(src, non_blocking=True)
                    static_tgt.copy_(target, non_blocking=True)
                    fgraph.replay()

            fwd_steps(5)
            barrier(world)
            fwd_ms, _, _ = timed_loop(fwd_steps, args.steps, world)

    if world > 1:
        tms = torch.tensor([ms, e2e_ms, fwd_ms or 0.0], device=dev, dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(tms[0]), float(tms[1])
        fwd_ms = float(tms[2]) if fwd_ms is not None else None

    # ---- kernel table, roofline of the dominant call, dense-stage TFLOP/s ----
    plan = next((e.tail_plan for e in posatt.mesh_cache.entries.values() if e.tail_plan), None)
    out_dim = model.out_dim

    def describe(k, v):
        t_s = v["ms_avg"] * 1e-3
        row = {"call": k[0], "mesh_batched": bool(k[2]), "B": k[3], "H": k[4], "N": k[5], "M": k[6], "D": k[7], "concat": bool(k[9]),
               "ms_avg": v["ms_avg"], "calls_per_step": v["calls"] / k_steps, "share_of_step": v["ms_total"] / k_steps / ms}
        if k[0].startswith("tail") and plan is not None:
            nbytes, flops, ksteps = tail_work(k, plan, out_dim)
            row.update({"bytes": nbytes, "hbm_frac": nbytes / t_s / 1e9 / hbm_peak, "tf32_flops_issued": flops,
                        "tensor_frac": flops / t_s / 1e12 / tf32_peak, "plan_ksteps": ksteps,
                        "unfused_equivalent": {"bytes": unfused_bytes(k), "GBps": unfused_bytes(k) / t_s / 1e9,
                                               "frac": unfused_bytes(k) / t_s / 1e9 / hbm_peak}})
        elif k[0].startswith("processor"):
            # the fused processor (csrc/processor_block.cuh): k[9] carries the number of blocks.  Per block the attention products are
            # 2 H N N B D flop forward (C = P X) and twice that backward (dP and the value gradient), issued as 3xTF32; the Linear
            # products 2 B N D ((1+H) D + D) flop forward and twice that backward, issued as the matmul precision says.
            # Bytes: block inputs and the saved C / Z1 / Z2 (written forward, read backward), dC through the L2 scratch, in / out.
            _, _, _, B, H, N, _, D, _, nb = k
            bwd = k[0] == "processor_bwd"
            att = 2.0 * H * N * N * B * D * nb * (2 if bwd else 1)
            lin = 2.0 * B * N * D * ((1 + H) * D + D) * nb * (2 if bwd else 1)
            lin_terms = 3 if args.precision == "highest" else 1
            nbytes = 4 * B * N * (2 * D + nb * (H * D + 3 * D) + (2 * nb * H * D if bwd else 0))
            row.update({"concat": True, "n_blocks": nb, "bytes": nbytes, "hbm_frac": nbytes / t_s / 1e9 / hbm_peak,
                        "dense_flops": att, "linear_flops": lin, "posatt_tflops": att / t_s / 1e12,
                        "posatt_tflops_issued": 3 * att / t_s / 1e12, "tf32_flops_issued": 3 * att + lin_terms * lin,
                        "tensor_frac": (3 * att + lin_terms * lin) / t_s / 1e12 / tf32_peak})
        else:
            nbytes = unfused_bytes(k)
            row.update({"bytes": nbytes, "hbm_frac": nbytes / t_s / 1e9 / hbm_peak})
            if k[0] in ("fwd", "bwd") and k[9]:      # global self stage (locality 1.0): a dense contraction on tcgen05, 3xTF32
                _, _, batched, B, H, N, M, D, _, _ = k
                dense = 2.0 * H * N * M * B * D * (1 if k[0] == "fwd" else 2)
                row.update({"dense_flops": dense, "posatt_tflops": dense / t_s / 1e12, "posatt_tflops_issued": 3 * dense / t_s / 1e12,
                            "tensor_frac": 3 * dense / t_s / 1e12 / tf32_peak})
        return row

    table = sorted((describe(k, v) for k, v in kernels.items()), key=lambda r: -r["share_of_step"])
    top = next((r for r in table if r["call"] != "rowstat"), None)
    dense = [r for r in table if "posatt_tflops" in r]
    total = batch * world
    result = {
        "value": total / (ms * 1e-3), "ms_per_step": ms, "timed_steps": timed_steps, "per_gpu_batch": batch, "global_batch": total,
        "e2e": {"value": total / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_ms, "timed_steps": e2e_steps},
        "gpu_launches_per_step": launches_per_step, "eager_ms_per_step": eager_ms,
        "dominant_kernel": top,
        "posatt_dense": None if not dense else {
            "tflops_algorithmic": max(r["posatt_tflops"] for r in dense), "tflops_issued_tf32": max(r["posatt_tflops_issued"] for r in dense),
            "frac_of_tf32_gemm": max(r["tensor_frac"] for r in dense), "stage": {k: dense[0][k] for k in ("B", "H", "N", "M", "D")}},
        "tail_plan": None if plan is None else {"tiles": plan.n_tiles, "candidate_entries": plan.n_cand},
    }
    if detail:
        result["forward_only"] = None if fwd_ms is None else {"value": total / (fwd_ms * 1e-3), "unit": UNIT, "ms_per_step": fwd_ms,
                                                              "what": "forward + loss under no_grad, device-resident inputs"}
        result["kernels"] = table[:8]
        result["window"] = window
    del step, opt, model, w, resident, stage
    posatt.mesh_cache.clear()
    torch.cuda.empty_cache()
    return result


def run_ours(args, rank, world, local_rank):
    from position_induced_transformer_b200 import workload_specs
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU path)"
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    torch.set_float32_matmul_precision(args.precision)
    hbm_peak, peak_src = load_peaks()
    tf32_peak = measure_tf32_peak(dev)

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
        clocks.wait_first()          # no timed region starts before the sampler has produced a row

    primary = measure_workload(args.workload, args, rank, world, dev, True, tf32_peak, hbm_peak)
    sweep = {args.workload: primary}
    if not args.no_sweep:
        for name in workload_specs.BASELINE_WORKLOADS:
            if name not in sweep:
                sweep[name] = measure_workload(name, args, rank, world, dev, False, tf32_peak, hbm_peak)
    clock_info = clocks.window(*primary["window"]) if rank == 0 else None
    clock_all = clocks.window(0.0, float("inf")) if rank == 0 else None
    clocks.stop()
    if rank != 0:
        return

    batch = primary["per_gpu_batch"]
    top = primary["dominant_kernel"]
    roofline = None
    if top is not None:
        hbm_frac, tensor_frac = top["hbm_frac"], top.get("tensor_frac", 0.0)
        t_s = top["ms_avg"] * 1e-3
        if tensor_frac >= hbm_frac:
            bound, ach, peak, unit = "tensor", top.get("tf32_flops_issued", 3 * top.get("dense_flops", 0.0)) / t_s / 1e12, tf32_peak, "TFLOP/s"
        else:
            bound, ach, peak, unit = "hbm", top["bytes"] / t_s / 1e9, hbm_peak, "GB/s"
        roofline = {"bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                    "traffic": None, "hbm_frac": hbm_frac, "tensor_frac": tensor_frac,
                    "kernel": {k: top[k] for k in ("call", "mesh_batched", "B", "H", "N", "M", "D")},
                    "algorithmic_bytes_per_launch": top["bytes"], "tf32_flops_issued_per_launch": top.get("tf32_flops_issued"),
                    "avg_launch_ms": top["ms_avg"], "share_of_step": top["share_of_step"],
                    "peak_source": {"hbm": peak_src, "tensor": "cuBLAS TF32 GEMM 8192^3 timed in this run (3xTF32 kernels are compared with the TF32 rate they issue at)",
                                    "tf32_tflops": tf32_peak, "hbm_gbs": hbm_peak},
                    "unfused_equivalent": top.get("unfused_equivalent"),
                    "note": "the fused decoder tail is bound by instruction issue (GELU epilogue on the fp32 pipes), not by HBM or the tensor pipe: both "
                            "fractions are small by design of the fusion; profiles/ holds the ncu evidence" if top["call"].startswith("tail") else None}
        traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(traffic_path):
            t = json.load(open(traffic_path))
            roofline["traffic"] = t.get(f"{top['call']}:B{top['B']}:H{top['H']}:N{top['N']}:M{top['M']}:D{top['D']}")

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        sample = args.cpu_sample_batch or auto_cpu_sample(args.workload, batch)
        v, dt, cores = cpu_oracle_steps(args.workload, sample, 3, 1)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"3 train steps of {sample} samples (of the {batch}-sample batch) through oracle/pit_oracle.py after 1 warm-up; {dt:.2f} s/step"}
    gpu_ref = None
    if not args.no_gpu_reference and world == 1:
        gpu_ref = gpu_eager_reference(args.workload, batch, dev)

    line = {
        "metric": METRIC, "value": primary["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": primary["ms_per_step"], "timed_steps": primary["timed_steps"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.precision == "highest" else "f32 (posatt) + tf32 (MLP Linears, as reference pit.py:2)",
        "data": "synthetic",
        "config": {"workload": args.workload, "per_gpu_batch": batch, "global_batch": batch * world, "step": STEP_DESC,
                   "launch": "eager" if args.no_graph else "cuda_graph_replay", "parallelism": f"dp{world}", "mlp_matmul_precision": args.precision,
                   "optimizer": ("gradient SUM over NVLink peer memory + Adam in one launch (pit_allreduce_adam)"
                                 if OPTIMIZER_USED.get(args.workload) == "fused"
                                 else "torch.optim.Adam(fused)" + (" after one NCCL all-reduce of the flat gradient" if world > 1 else "")),
                   "timing": f"{args.steps} steps per repeat, repeated until the timed region lasts >= {MIN_TIMED_S} s",
                   "l2": "per-step working set (activations of the decoder stage) exceeds the 126 MB L2 and input batches rotate over 4 buffers; no explicit flush"},
        "e2e": primary["e2e"], "forward_only": primary.get("forward_only"),
        "gpu_launches": primary["gpu_launches_per_step"] * primary["timed_steps"], "gpu_launches_per_step": primary["gpu_launches_per_step"],
        "eager_ms_per_step": primary["eager_ms_per_step"], "roofline": roofline, "posatt_dense": primary["posatt_dense"],
        "cpu_baseline": cpu, "gpu_eager_reference": gpu_ref, "clocks": clock_info, "clocks_whole_run": clock_all,
        "kernels": primary.get("kernels"),
        "workloads": {name: {k: r[k] for k in ("value", "ms_per_step", "timed_steps", "per_gpu_batch", "e2e", "gpu_launches_per_step",
                                                "dominant_kernel", "posatt_dense", "tail_plan")} for name, r in sweep.items()},
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
    run_ours(args, rank, world, local_rank)
    if world > 1:
        # every rank has finished (rank 0 has printed its line): leave without tearing down NCCL communicators, captured graphs
        # and peer-mapped buffers one by one -- their destructors wait for each other across ranks in an order Python's
        # shutdown does not guarantee (observed: a hang at exit after a run with symmetric memory)
        import torch.distributed as dist
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
