"""Synthetic dataset writers: the files have the names, keys and layouts the reference's loaders read."""
import importlib.util
import os
import re

import numpy as np
import pytest

from position_induced_transformer_b200 import synthetic_data as sd

REF = "/root/reference"


def _load_data_of(script):
    """The script's own load_data(), cut out of the unmodified source (the rest of the script trains on a GPU)."""
    if not os.path.exists(os.path.join(REF, script)):
        pytest.skip("reference sources only exist in the build container")
    src = open(os.path.join(REF, script)).read()
    body = re.search(r"^def load_data\(.*?(?=^class )", src, re.S | re.M).group(0)
    ns = {}
    exec("import numpy as np\nimport torch\nfrom scipy.io import loadmat\n" + body, ns)
    return ns["load_data"]


def test_burgers_and_sod_files_feed_the_reference_loaders(tmp_path):
    sd.write_burgers(str(tmp_path), samples=12)
    sd.write_sod(str(tmp_path), samples=12)
    xb = _load_data_of("train_burgers.py")(str(tmp_path / "supplementary_data" / "data_burgers.mat"), 1024, 128)
    assert [tuple(t.shape) for t in xb] == [(12, 1024, 1)] * 4
    xs = _load_data_of("train_sod.py")(str(tmp_path / "supplementary_data" / "data_sod.mat"), 1024, 128)
    assert [tuple(t.shape) for t in xs] == [(12, 2048, 3)] * 4
    assert all(bool(np.isfinite(t.numpy()).all()) for t in xs)       # positive density: the primitive-variable division is safe


def test_darcy_files_feed_the_reference_loader(tmp_path):
    sd.write_darcy(str(tmp_path), samples=3)
    out = _load_data_of("train_darcy.py")(str(tmp_path / "piececonst_r421_N1024_smooth1.mat"),
                                          str(tmp_path / "piececonst_r421_N1024_smooth2.mat"), 10, 1024, 100)
    assert [tuple(t.shape) for t in out] == [(3, 43, 43, 1)] * 4
    assert set(np.unique(out[0].numpy())) <= {3.0, 12.0}


def test_point_cloud_files_feed_the_reference_loaders(tmp_path):
    root = str(tmp_path) + os.sep
    sd.write_elasticity(root, samples=7)
    x, ext, y, *_ = _load_data_of("train_elasticity.py")(root, 1000, 200)
    assert tuple(x.shape) == (7, 972, 44) and tuple(ext.shape) == (7, 972, 2) and tuple(y.shape) == (7, 972, 1)
    assert float(x[..., 2:].min()) >= 0.0 and float(x[..., 2:].max()) <= 1.0
    sd.write_naca(root, samples=5)
    c, g, q, *_ = _load_data_of("train_naca.py")(root, 1000, 200)
    assert tuple(c.shape) == (5, 120, 2) and tuple(g.shape) == (5, 221, 51, 2) and tuple(q.shape) == (5, 221, 51, 4)
