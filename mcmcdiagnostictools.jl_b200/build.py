"""Builds libmcmcdiag_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Each translation unit is compiled to an object file (in parallel, only when stale) and the objects are
linked into the shared library, so that touching one kernel family rebuilds in seconds."""
from __future__ import annotations

import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libmcmcdiag_b200.so")
COMMON = ["mcd_common.cuh", "mcd_slab.cuh", "mcd_fast.cuh", "mcd_rk2_api.cuh", "mcd_big_api.cuh"]
# translation unit -> headers it depends on (besides COMMON and the public header)
UNITS = {
    "mcd_api.cu": ["mcd_fastgen.cuh", "mcd_large.cuh", "mcd_crank.cuh"],
    "mcd_rk2.cu": ["mcd_rk2.cuh", "mcd_tma.cuh"],
    "mcd_big.cu": ["mcd_big.cuh", "mcd_big_api.cuh", "mcd_tma.cuh"],
}
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _mtime(p: str) -> float:
    return os.path.getmtime(p) if os.path.exists(p) else 0.0


def _deps(unit: str) -> list[str]:
    d = [os.path.join(CSRC, unit)] + [os.path.join(CSRC, h) for h in COMMON + UNITS[unit]]
    d.append(os.path.join(HERE, "..", "include", "mcmcdiag_b200.h"))
    d.append(os.path.abspath(__file__))
    return d


def _compile(unit: str, nvcc: str, verbose: bool) -> str:
    obj = os.path.join(OBJ, unit.replace(".cu", ".o"))
    cmd = [nvcc, *NVCC_FLAGS, "-c", "-o", obj, os.path.join(CSRC, unit)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    stale = []
    for unit in UNITS:
        obj = os.path.join(OBJ, unit.replace(".cu", ".o"))
        if force or _mtime(obj) < max(_mtime(d) for d in _deps(unit)):
            stale.append(unit)
    if stale:
        with concurrent.futures.ThreadPoolExecutor(max_workers=len(stale)) as ex:
            list(ex.map(lambda u: _compile(u, nvcc, verbose), stale))
    objs = [os.path.join(OBJ, u.replace(".cu", ".o")) for u in UNITS]
    if stale or _mtime(LIB) < max(_mtime(o) for o in objs):
        subprocess.run([nvcc, "-shared", "-o", LIB, *objs], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
