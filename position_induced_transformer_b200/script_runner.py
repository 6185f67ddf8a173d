"""Run one of the reference's UNMODIFIED training scripts against this package for a bounded number of optimizer steps.

    python -m position_induced_transformer_b200.script_runner /path/to/train_burgers.py [--steps 12] [--compile asis|off]

The scripts take no arguments, read data files from the working directory, import matplotlib and train for hundreds of
epochs (train_burgers.py:51-58).  The runner leaves the script text alone and arranges the world around it instead:
  * a scratch working directory holding synthetic data files in the loaders' formats (synthetic_data.py);
  * `pit` / `utils` resolve to this repository's drop-in modules (repo root first on sys.path) -- `from pit import *`;
  * a stub `matplotlib.pyplot` (not installed here; the scripts only import it);
  * `torch.optim.Adam.step` counts steps and stops the script after `steps` of them; every scalar `.backward()` is
    recorded, so the caller sees the loss curve of the truncated run.
`torch.compile(model)` (train_burgers.py:73) is left as the script wrote it with --compile asis: the fused ops are opaque to
the tracing compiler (graph breaks around them, see posatt.py), everything else is compiled by Inductor; --compile off
replaces torch.compile by the identity.
"""
from __future__ import annotations

import argparse
import contextlib
import os
import sys
import tempfile
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Done(Exception):
    pass


def _stub_matplotlib():
    if "matplotlib" in sys.modules:
        return []
    mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt

    def _anything(name):
        if name.startswith("__"):
            raise AttributeError(name)
        return lambda *a, **k: None

    plt.__getattr__ = _anything
    mpl.__file__ = plt.__file__ = os.path.join(ROOT, "position_induced_transformer_b200", "script_runner.py")
    sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    return ["matplotlib", "matplotlib.pyplot"]


def run_script(path: str, steps: int = 12, compile_mode: str = "asis", workdir: str | None = None, data_seed: int = 0) -> dict:
    """Execute the script at `path` unchanged until `steps` optimizer steps have run.  Returns
    {"steps", "losses", "model_class_module", "launches", "namespace"}."""
    from . import _cabi, synthetic_data
    name = os.path.basename(path)
    if name not in synthetic_data.WRITERS:
        raise ValueError(f"no synthetic dataset writer for {name}")
    source = open(path).read()
    losses, state = [], {"steps": 0}
    orig_step, orig_backward, orig_compile = torch.optim.Adam.step, torch.Tensor.backward, torch.compile

    def counting_step(self, *a, **k):
        out = orig_step(self, *a, **k)
        state["steps"] += 1
        if state["steps"] >= steps:
            raise _Done()
        return out

    def recording_backward(self, *a, **k):
        if self.dim() == 0:
            losses.append(float(self.detach()))
        return orig_backward(self, *a, **k)

    stubbed = _stub_matplotlib()
    saved_path, saved_cwd = list(sys.path), os.getcwd()
    saved_mods = {m: sys.modules.pop(m) for m in ("pit", "utils") if m in sys.modules}
    before = _cabi.launch_count()
    ns = {"__name__": "__main__", "__file__": path}
    with contextlib.ExitStack() as stack:
        tmp = workdir or stack.enter_context(tempfile.TemporaryDirectory(prefix="pit_script_"))
        synthetic_data.WRITERS[name](tmp, seed=data_seed)
        try:
            sys.path.insert(0, ROOT)
            os.chdir(tmp)
            torch.optim.Adam.step, torch.Tensor.backward = counting_step, recording_backward
            if compile_mode == "off":
                torch.compile = lambda model=None, *a, **k: model
            try:
                exec(compile(source, path, "exec"), ns)
            except _Done:
                pass
        finally:
            torch.optim.Adam.step, torch.Tensor.backward, torch.compile = orig_step, orig_backward, orig_compile
            os.chdir(saved_cwd)
            sys.path[:] = saved_path
            for m in ("pit", "utils"):
                sys.modules.pop(m, None)
            sys.modules.update(saved_mods)
            for m in stubbed:
                sys.modules.pop(m, None)
    model = ns.get("model")
    inner = getattr(model, "_orig_mod", model)
    return {"steps": state["steps"], "losses": losses, "launches": _cabi.launch_count() - before,
            "model_class_module": type(inner).__mro__[1].__module__ if inner is not None else None, "namespace": ns}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("script")
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--compile", default="asis", choices=["asis", "off"])
    args = ap.parse_args()
    res = run_script(args.script, args.steps, args.compile)
    print(f"{os.path.basename(args.script)}: {res['steps']} optimizer steps, {res['launches']} launches of libpit_posatt.so, "
          f"loss {res['losses'][0]:.4f} -> {res['losses'][-1]:.4f}; model built from {res['model_class_module']}")


if __name__ == "__main__":
    main()
