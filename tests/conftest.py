import glob
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_names(prefix):
    return sorted(os.path.basename(p)[len(prefix):-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def t(a, device="cpu"):
    a = np.asarray(a)
    return torch.from_numpy(np.ascontiguousarray(a)).reshape(a.shape).to(device)


def variant_of(cls_name: str) -> str:
    if "periodic1d" in cls_name:
        return "periodic1d"
    if "periodic2d" in cls_name:
        return "periodic2d"
    return "euclid"


def rel_linf(a: torch.Tensor, b: torch.Tensor, floor: float = 1e-30) -> float:
    """max|a-b| / max(max|b|, floor)."""
    return float((a - b).abs().max() / b.abs().max().clamp_min(floor))


@pytest.fixture
def host_scale_map():
    """Evaluate tan/sin of lmda with the CPU's libm, like the CPU run of the reference that made the goldens."""
    import position_induced_transformer_b200.pit as pit_mod
    pit_mod.use_host_scale_map(True)
    yield
    pit_mod.use_host_scale_map(False)


@pytest.fixture(scope="session")
def cuda_device():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
