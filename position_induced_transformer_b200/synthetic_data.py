"""Synthetic dataset files in the formats the reference's loaders read.

The reference's data files are Git-LFS stubs or not distributable; each writer below produces files of the names, keys,
array layouts and dtypes that the corresponding ``load_data`` expects (train_burgers.py:7-16, train_sod.py:7-21,
train_darcy.py:7-23, train_elasticity.py:7-16, train_naca.py:7-15), filled by the seeded generators of ``workload_specs``.
Only the number of samples is smaller than the real sets: the loaders slice ``[:ntrain]`` / ``[-ntest:]``, which simply
yields fewer samples.  No CUDA, no torch ops beyond the generators.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import workload_specs as WS


def _gen(seed: int) -> torch.Generator:
    return torch.Generator().manual_seed(seed)


def write_burgers(root: str, samples: int = 48, seed: int = 0) -> None:
    """supplementary_data/data_burgers.mat: x, y (samples, 1024) (train_burgers.py:9-12, 58)."""
    from scipy.io import savemat
    (x,), y = WS.burgers().make_batch(_gen(seed), samples)
    os.makedirs(os.path.join(root, "supplementary_data"), exist_ok=True)
    savemat(os.path.join(root, "supplementary_data", "data_burgers.mat"), {"x": x[..., 0].numpy(), "y": y[..., 0].numpy()})


def write_sod(root: str, samples: int = 48, seed: int = 0) -> None:
    """supplementary_data/data_sod.mat: x, y (samples, 2048, 3) conservative variables with positive density so that the
    loader's conversion to primitive variables (train_sod.py:12-16) is well defined."""
    from scipy.io import savemat
    (x,), y = WS.sod().make_batch(_gen(seed), samples)
    os.makedirs(os.path.join(root, "supplementary_data"), exist_ok=True)
    savemat(os.path.join(root, "supplementary_data", "data_sod.mat"), {"x": x.numpy(), "y": y.numpy()})


def write_darcy(root: str, samples: int = 16, seed: int = 0) -> None:
    """piececonst_r421_N1024_smooth{1,2}.mat: coeff, sol (samples, 421, 421) (train_darcy.py:11-21, 62-63)."""
    from scipy.io import savemat
    spec = WS.darcy(421)
    for idx, name in enumerate(("piececonst_r421_N1024_smooth1.mat", "piececonst_r421_N1024_smooth2.mat")):
        (coeff,), sol = spec.make_batch(_gen(seed + idx), samples)
        savemat(os.path.join(root, name), {"coeff": (coeff[..., 0] * 4.5 + 7.5).numpy(), "sol": sol[..., 0].numpy()})


def write_elasticity(root: str, samples: int = 40, seed: int = 0) -> None:
    """Random_UnitCell_{rr,XY,sigma}_10.npy: (42, n), (972, 2, n), (972, n) (train_elasticity.py:9-14)."""
    (xy, feats, _), sigma = WS.elasticity().make_batch(_gen(seed), samples)
    codes = (feats[:, 0, 2:] + 1) / 5                     # the loader maps rr -> 5 rr - 1
    np.save(os.path.join(root, "Random_UnitCell_rr_10.npy"), codes.numpy().T.astype("float32"))
    np.save(os.path.join(root, "Random_UnitCell_XY_10.npy"), xy.permute(1, 2, 0).contiguous().numpy())
    np.save(os.path.join(root, "Random_UnitCell_sigma_10.npy"), sigma[..., 0].numpy().T.copy())


def write_naca(root: str, samples: int = 60, seed: int = 0) -> None:
    """shape_coords.npy (n, 120, 2), NACA_Cylinder_{X,Y}.npy (n, 221, 51), NACA_Cylinder_Q.npy (n, 5, 221, 51)
    (train_naca.py:8-13)."""
    (foil, _, grid), q = WS.naca().make_batch(_gen(seed), samples)
    np.save(os.path.join(root, "shape_coords.npy"), foil.numpy())
    np.save(os.path.join(root, "NACA_Cylinder_X.npy"), grid[..., 0].numpy())
    np.save(os.path.join(root, "NACA_Cylinder_Y.npy"), grid[..., 1].numpy())
    q5 = torch.cat((q, q[..., :1]), -1).permute(0, 3, 1, 2).contiguous()      # the loader keeps the first four of five fields
    np.save(os.path.join(root, "NACA_Cylinder_Q.npy"), q5.numpy())


WRITERS = {"train_burgers.py": write_burgers, "train_sod.py": write_sod, "train_darcy.py": write_darcy,
           "train_elasticity.py": write_elasticity, "train_naca.py": write_naca}
