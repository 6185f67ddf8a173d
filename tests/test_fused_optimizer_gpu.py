"""FusedAllReduceAdam: one launch = cross-rank gradient SUM over peer memory + Adam.  Single-rank parity with torch.optim.Adam
here; the two-rank run (peer memory over NVLink, compared with NCCL all-reduce + torch Adam) needs two GPUs and is skipped on
a one-GPU box (scripts/two_rank_optimizer_check.py is the same check as a torchrun script)."""
import copy
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _model(dev):
    torch.manual_seed(3)
    return torch.nn.Sequential(torch.nn.Linear(7, 33), torch.nn.GELU(), torch.nn.Linear(33, 5)).to(dev)


def test_single_rank_matches_torch_adam(cuda_device):
    from position_induced_transformer_b200.fused_optimizer import FusedAllReduceAdam
    a, b = _model(cuda_device), None
    b = copy.deepcopy(a)
    opt_a = FusedAllReduceAdam(a.parameters(), lr=3e-3)
    opt_b = torch.optim.Adam(b.parameters(), lr=3e-3)
    g = torch.Generator().manual_seed(0)
    for it in range(25):
        x = torch.randn(16, 7, generator=g).to(cuda_device)
        for m, o in ((a, opt_a), (b, opt_b)):
            o.zero_grad()
            m(x).square().mean().backward()
            o.step()
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert float((pa - pb).abs().max()) <= 2e-6 * max(1.0, float(pb.abs().max()))
    assert int(opt_a.step_count) == 25 and not opt_a.peer_timeout()
    assert all(p.data_ptr() >= opt_a.flat_param.data_ptr() for p in a.parameters())       # parameters are views of the flat buffer


def test_single_rank_large_model_grid_is_capped(cuda_device):
    """1.3 M parameters (the elasticity model's size): more work groups than CTAs that fit the chip at once -- the grid is capped by
    the occupancy query and every thread takes several groups."""
    from position_induced_transformer_b200.fused_optimizer import FusedAllReduceAdam
    torch.manual_seed(1)
    a = torch.nn.Sequential(torch.nn.Linear(1024, 1024), torch.nn.Linear(1024, 256)).to(cuda_device)
    b = copy.deepcopy(a)
    opt_a, opt_b = FusedAllReduceAdam(a.parameters(), lr=1e-3), torch.optim.Adam(b.parameters(), lr=1e-3)
    g = torch.Generator(device=cuda_device).manual_seed(5)
    for _ in range(3):
        # the same gradient tensors for both (two backward passes through differently aligned parameter storage may differ in the
        # last bit, which Adam amplifies without bound for gradients near eps)
        for pa, pb in zip(a.parameters(), b.parameters()):
            grad = torch.randn(pa.shape, generator=g, device=cuda_device) * 1e-2
            pa.grad, pb.grad = grad, grad.clone()
        opt_a.step()
        opt_b.step()
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert float((pa.detach() - pb.detach()).abs().max()) <= 2e-6 * max(1.0, float(pb.detach().abs().max()))


def test_single_rank_step_is_graph_capturable(cuda_device):
    from position_induced_transformer_b200.fused_optimizer import FusedAllReduceAdam
    from position_induced_transformer_b200.graphed import GraphedTrainStep
    a = _model(cuda_device)
    b = copy.deepcopy(a)
    x, y = torch.randn(8, 7, device=cuda_device), torch.randn(8, 5, device=cuda_device)
    loss_fn = lambda m: (lambda ins, tgt: (m(ins[0]) - tgt).square().mean())
    step_a = GraphedTrainStep(list(a.parameters()), loss_fn(a), FusedAllReduceAdam(a.parameters(), lr=1e-2), (x,), y, warmup=2)
    opt_b = torch.optim.Adam(b.parameters(), lr=1e-2)
    for _ in range(2):                     # the two eager warm-up steps already stepped a (capturing records, it does not run)
        opt_b.zero_grad()
        loss_fn(b)((x,), y).backward()
        opt_b.step()
    for _ in range(5):
        la = float(step_a((x,), y))
        opt_b.zero_grad()
        lb = loss_fn(b)((x,), y)
        lb.backward()
        opt_b.step()
        assert abs(la - float(lb)) <= 1e-5 * max(1.0, abs(float(lb)))


def test_two_ranks_match_nccl_plus_torch_adam():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import socket
    with socket.socket() as sock:          # a free port: a fixed one may still be in TIME_WAIT from an earlier run
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    proc = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                           "--master-port", str(port), os.path.join(ROOT, "scripts", "two_rank_optimizer_check.py")],
                          capture_output=True, text=True, timeout=240, cwd=ROOT)
    assert proc.returncode == 0 and "OK" in proc.stdout, proc.stdout[-2000:] + proc.stderr[-2000:]
