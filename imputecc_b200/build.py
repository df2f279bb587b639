"""Build libcrwr_b200.so (sm_100a) in-tree with nvcc.  `python -m imputecc_b200.build [--force]`."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcrwr_b200.so")
SOURCES = ["api.cu"]
DEPS = ["api.cu", "common.cuh", "propagate.cuh", "propagate_group.cuh", "structure.cuh", "structure_warp.cuh", "tail.cuh", "select.cuh", "transition.cuh", "finish.cuh", "peer.cuh", "binweights.cuh",
        os.path.join("..", "..", "include", "crwr_b200.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    out = os.environ.get("CRWR_LIB_OUT", LIB)
    extra = os.environ.get("CRWR_NVCC_FLAGS", "").split()
    cmd = [_nvcc(), "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-Xptxas", "-v", "-Xcompiler", "-fPIC", "-shared", "-o", out] + extra + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    log = proc.stdout + proc.stderr
    with open(os.path.join(HERE, "build.log"), "w") as fh:
        fh.write(" ".join(cmd) + "\n" + log)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log[-4000:])
    if verbose:
        print(log)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
