"""Probe: can two ranks on one box map each other's device memory (torch symmetric memory), and how long does NCCL's
all-reduce of a 316 KB bucket take?   torchrun --nproc-per-node 2 scripts/probe/symm_probe.py"""
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    n = 79_000
    x = torch.full((n,), float(rank + 1), device=dev)
    for _ in range(5):
        dist.all_reduce(x)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(50):
        dist.all_reduce(x)
    e.record()
    torch.cuda.synchronize()
    if rank == 0:
        print(f"nccl all_reduce of {4 * n} bytes over {world} ranks: {s.elapsed_time(e) / 50 * 1e3:.1f} us", flush=True)
    try:
        import torch.distributed._symmetric_memory as symm
        t = symm.empty(n, dtype=torch.float32, device=dev)
        hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
        t.fill_(float(rank + 1))
        torch.cuda.synchronize()
        dist.barrier()
        peer = hdl.get_buffer((rank + 1) % world, (n,), torch.float32)
        val = float(peer[0])
        print(f"rank {rank}: symmetric memory ok, peer value {val}, buffer_ptrs {[hex(p) for p in hdl.buffer_ptrs]}, "
              f"signal_pad_ptrs {len(hdl.signal_pad_ptrs)}", flush=True)
    except Exception as ex:  # noqa: BLE001
        print(f"rank {rank}: symmetric memory failed: {type(ex).__name__}: {ex}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
