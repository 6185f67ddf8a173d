"""world_size-2 gloo test of the data-parallel plumbing (no GPU): the flat SUM all-reduce over two ranks that
each hold half of the batch reproduces the single-process gradient of the whole batch."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from position_induced_transformer_b200.data_parallel import FlatGradients, shard_range
from position_induced_transformer_b200.utils import RelLpNorm


def _model():
    torch.manual_seed(3)
    return torch.nn.Sequential(torch.nn.Linear(5, 16), torch.nn.GELU(), torch.nn.Linear(16, 2))


def _data():
    g = torch.Generator().manual_seed(4)
    return torch.randn(6, 20, 5, generator=g), torch.randn(6, 20, 2, generator=g)


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = _model()
    flat = FlatGradients(model.parameters())
    x, y = _data()
    mine = list(shard_range(x.shape[0], rank, world))
    flat.zero()
    RelLpNorm(2, 2)(y[mine], model(x[mine])).backward()
    flat.all_reduce()
    torch.save(flat.buffer.clone(), os.path.join(out_dir, f"g{rank}.pt"))
    # the step the benchmark runs: grads released before backward, packed with one copy, then reduced
    flat.release()
    RelLpNorm(2, 2)(y[mine], model(x[mine])).backward()
    flat.gather()
    flat.all_reduce()
    torch.save(flat.buffer.clone(), os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_sum_equals_full_batch_gradient(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    model = _model()
    flat = FlatGradients(model.parameters(), world_size=1)
    x, y = _data()
    flat.zero()
    RelLpNorm(2, 2)(y, model(x)).backward()
    for r in range(2):
        for tag in ("g", "r"):          # accumulate-into-views path and release/gather path
            got = torch.load(os.path.join(str(tmp_path), f"{tag}{r}.pt"))
            assert torch.allclose(got, flat.buffer, rtol=1e-5, atol=1e-7)


def test_shard_range_covers_everything():
    for n in (1, 7, 8, 9, 64):
        for world in (1, 2, 3, 8):
            seen = [i for r in range(world) for i in shard_range(n, r, world)]
            assert seen == list(range(n))


def test_grads_stay_views_after_steps():
    model = _model()
    flat = FlatGradients(model.parameters(), world_size=1)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    x, y = _data()
    for _ in range(2):
        flat.zero()
        RelLpNorm(2, 2)(y, model(x)).backward()
        flat.all_reduce()
        opt.step()
    p = next(model.parameters())
    assert p.grad.untyped_storage().data_ptr() == flat.buffer.untyped_storage().data_ptr()
    assert float(flat.buffer.abs().sum()) > 0


def test_release_gather_equals_accumulate():
    """release() + backward + gather() packs the same gradients that zero() + backward accumulates into the views."""
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.GELU(), torch.nn.Linear(7, 2))
    x = torch.randn(11, 5)
    flat = FlatGradients(model.parameters(), world_size=1)
    flat.zero()
    model(x).square().sum().backward()
    want = flat.buffer.clone()
    flat.release()
    assert all(p.grad is None for p in model.parameters())
    model(x).square().sum().backward()
    flat.gather()
    flat.all_reduce()
    assert torch.equal(flat.buffer, want)
    for p in model.parameters():
        assert p.grad.untyped_storage().data_ptr() == flat.buffer.untyped_storage().data_ptr()
