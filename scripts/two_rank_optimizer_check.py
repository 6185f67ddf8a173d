"""Two (or more) ranks: FusedAllReduceAdam (gradient SUM over peer memory + Adam in one launch) against NCCL all-reduce +
torch.optim.Adam on the same model and per-rank data, eager and under CUDA-graph replay; also times both.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/two_rank_optimizer_check.py
"""
import copy
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from position_induced_transformer_b200 import workloads  # noqa: E402
from position_induced_transformer_b200.data_parallel import FlatGradients  # noqa: E402
from position_induced_transformer_b200.fused_optimizer import FusedAllReduceAdam  # noqa: E402
from position_induced_transformer_b200.graphed import GraphedTrainStep  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    torch.set_float32_matmul_precision("highest")
    wa = workloads.make("darcy43", 4).to(dev)
    wb = copy.deepcopy(wa)
    gen = torch.Generator().manual_seed(100 + rank)            # every rank has its own samples
    batches = [wa.make_batch(gen, 4) for _ in range(6)]
    batches = [(tuple(x.to(dev) for x in ins), tgt.to(dev)) for ins, tgt in batches]
    opt_a = FusedAllReduceAdam(wa.model.parameters(), lr=1e-3)
    opt_b = torch.optim.Adam(wb.model.parameters(), lr=1e-3)
    flat_b = FlatGradients(wb.model.parameters(), world)
    worst = 0.0
    for ins, tgt in batches:
        opt_a.zero_grad()
        workloads.step_loss(wa, ins, tgt).backward()
        opt_a.step()
        flat_b.release()
        workloads.step_loss(wb, ins, tgt).backward()
        flat_b.gather()
        flat_b.all_reduce()
        opt_b.step()
        for pa, pb in zip(wa.model.parameters(), wb.model.parameters()):
            worst = max(worst, float((pa - pb).abs().max() / pb.abs().max().clamp_min(1e-6)))
    # all ranks hold the same parameters
    mine = opt_a.flat_param.clone()
    ref = mine.clone()
    dist.broadcast(ref, 0)
    same = bool(torch.equal(mine, ref))
    # the optimizer step alone (gradients as the last backward left them), CUDA events over 200 calls
    def timed(fn, n=200):
        for _ in range(10):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / n * 1e3

    opt_t = torch.optim.Adam(wb.model.parameters(), lr=0.0, fused=True)

    def nccl_path():
        flat_b.gather()
        flat_b.all_reduce()
        opt_t.step()

    opt_a.set_lr(0.0)
    t_fused, t_nccl = timed(opt_a.step), timed(nccl_path)
    opt_a.set_lr(1e-3)
    if rank == 0:
        print(f"optimizer step alone (eager launches, {opt_a.total} parameters): fused {t_fused:.1f} us, pack + NCCL all-reduce + torch Adam {t_nccl:.1f} us", flush=True)
    # graph replay + timing of the step with either optimizer
    times = {}
    for tag, w, opt in (("fused", wa, opt_a), ("nccl+adam", wb, torch.optim.Adam(wb.model.parameters(), lr=1e-3, capturable=True, fused=True))):
        step = GraphedTrainStep(list(w.model.parameters()), lambda i, t, w=w: workloads.step_loss(w, i, t), opt, *batches[0], world)
        for _ in range(5):
            step(*batches[1])
        dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for k in range(200):
            step.replay()
        e.record()
        torch.cuda.synchronize()
        times[tag] = s.elapsed_time(e) / 200 * 1e3
    ok = worst <= 2e-4 and same and not opt_a.peer_timeout()
    flags = torch.tensor([float(ok)], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"world {world}: max relative parameter difference vs NCCL + torch Adam over 6 steps {worst:.2e}; ranks bit-identical: {same}; "
              f"graphed darcy43 step: fused {times['fused']:.1f} us, nccl+adam {times['nccl+adam']:.1f} us")
        print("OK" if float(flags) == 1.0 else "FAILED", flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0 if float(flags) == 1.0 else 1)      # no teardown of communicators / peer mappings across ranks (see bench.py)


if __name__ == "__main__":
    main()
