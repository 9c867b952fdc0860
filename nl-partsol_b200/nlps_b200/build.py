"""Build libnlps_b200.so (hand-written sm_100a CUDA + host setup) in-tree with nvcc.

The translation units are compiled in parallel into csrc/_obj/*.o (git-ignored) and linked; a unit is recompiled only
when it or one of the headers is newer than its object."""
from __future__ import annotations

import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(CSRC, "_obj")
SO = os.path.join(PKG, "libnlps_b200.so")
UNITS = ("nlps_engine.cu", "nlps_cellwarp.cu", "host_setup.cpp")
HEADERS = [os.path.join(CSRC, f) for f in ("nlps_device.cuh", "nlps_types.cuh", "nlps_cellwarp.h", "nlps_implicit.inl")] + \
          [os.path.join(PKG, "..", "include", "nlps_b200.h")]
EXTRA = os.environ.get("NLPS_NVCC_FLAGS", "").split()


def _obj(unit: str) -> str:
    return os.path.join(OBJ, os.path.splitext(unit)[0] + ".o")


def _unit_stale(unit: str) -> bool:
    o = _obj(unit)
    if not os.path.exists(o):
        return True
    t = os.path.getmtime(o)
    return any(os.path.getmtime(f) > t for f in [os.path.join(CSRC, unit)] + HEADERS)


def stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(f) > t for f in [os.path.join(CSRC, u) for u in UNITS] + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    base = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
            "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fPIC,-fopenmp,-O2"] + EXTRA
    if verbose:
        base.insert(1, "-Xptxas=-v")

    def compile_unit(unit: str) -> None:
        subprocess.check_call(base + ["-c", os.path.join(CSRC, unit), "-o", _obj(unit)])

    todo = [u for u in UNITS if force or _unit_stale(u)]
    with ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        list(ex.map(compile_unit, todo))
    subprocess.check_call([nvcc, "-shared", "-o", SO] + [_obj(u) for u in UNITS] + ["-lgomp", "-ldl"])
    return SO
