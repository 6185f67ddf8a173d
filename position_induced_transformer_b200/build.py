"""Build recipe for libpit_posatt.so (in-tree, sm_100a only).

    python -m position_induced_transformer_b200.build [--force] [--verbose]

The library is plain CUDA C++ behind a C ABI (include/pit_posatt.h): no torch headers.  Every kernel family has
its own translation unit (csrc/tu_*.cu) so that the template instantiations compile in parallel; objects go to
csrc/_obj/ and are linked into one shared library.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(PKG, "libpit_posatt.so")
HEADER = os.path.join(os.path.dirname(PKG), "include", "pit_posatt.h")
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libpit_posatt.so cannot be built")
    return exe


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))] + [HEADER]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    built = os.path.getmtime(target)
    return any(os.path.getmtime(d) > built for d in deps)


def _compile(src: str, verbose: bool) -> str:
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    cmd = [_nvcc(), *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-c", src, "-o", obj]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile whatever is missing or older than its sources and link; returns the library path."""
    deps = _deps()
    if not force and not _stale(LIB, deps):
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    headers = [d for d in deps if not d.endswith(".cu")]
    todo = [s for s in _sources()
            if force or _stale(os.path.join(OBJ, os.path.basename(s)[:-3] + ".o"), [s, *headers])]
    workers = max(1, min(len(todo), os.cpu_count() or 1))
    if todo:
        with concurrent.futures.ThreadPoolExecutor(max_workers=workers) as pool:
            list(pool.map(lambda s: _compile(s, verbose), todo))
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + ".o") for s in _sources()]
    cmd = [_nvcc(), "-shared", "-o", LIB, *objs]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
