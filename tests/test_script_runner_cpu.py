"""script_runner mechanics without a GPU: a stand-in script with the shape of the reference's train_*.py (star imports from `pit` /
`utils`, matplotlib import, loadmat of the loader's file, an Adam loop that would run for hundreds of epochs) is executed
unchanged, fed the synthetic data file and stopped after the step budget."""
import os
import sys
import textwrap

import torch

from position_induced_transformer_b200.script_runner import run_script

SCRIPT = textwrap.dedent('''
    from pit import *
    from timeit import default_timer
    import matplotlib.pyplot as plt
    from scipy.io import savemat, loadmat
    from utils import *

    data = loadmat('./supplementary_data/data_burgers.mat')
    x = torch.from_numpy(data["x"].astype('float32'))[..., None]
    y = torch.from_numpy(data["y"].astype('float32'))[..., None]
    model = kaiming_mlp(1, 8, 1)                      # a CPU module of the drop-in `pit` (no CUDA op involved)
    optimizer = torch.optim.Adam(model.parameters(), lr=1e-2)
    myloss = RelLpNorm(out_dim=1, p=2)
    plt.figure()
    for ep in range(500):
        for i in range(0, x.shape[0], 8):
            optimizer.zero_grad()
            loss = myloss(y[i:i + 8], model(x[i:i + 8]))
            loss.backward()
            optimizer.step()
    raise SystemExit("the step budget should have stopped the script long before")
''')


def test_runner_feeds_stops_and_restores(tmp_path):
    path = tmp_path / "train_burgers.py"
    path.write_text(SCRIPT)
    cwd, mods, step, backward = os.getcwd(), set(sys.modules), torch.optim.Adam.step, torch.Tensor.backward
    res = run_script(str(path), steps=7, compile_mode="off")
    assert res["steps"] == 7 and len(res["losses"]) == 7
    assert res["namespace"]["x"].shape == (48, 1024, 1)                      # the synthetic file in the loader's layout
    assert type(res["namespace"]["model"]).__module__ == "position_induced_transformer_b200.pit"
    # the world is put back: working directory, patched methods, stub modules, sys.path shadowing
    assert os.getcwd() == cwd and torch.optim.Adam.step is step and torch.Tensor.backward is backward
    assert "matplotlib" not in (set(sys.modules) - mods) and "pit" not in (set(sys.modules) - mods)
