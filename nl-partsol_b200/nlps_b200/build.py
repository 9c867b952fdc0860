"""Build libnlps_b200.so (hand-written sm_100a CUDA + host setup) in-tree with nvcc.

The translation units are compiled in parallel into csrc/_obj/*.o (git-ignored) and linked; a unit is recompiled only
when it or one of the headers is newer than its object."""
from __future__ import annotations

import hashlib
import json
import os
import re
import subprocess
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(CSRC, "_obj")
SO = os.path.join(PKG, "libnlps_b200.so")
STAMP = os.path.join(PKG, "libnlps_b200.sass.json")   # {kernel: sha256 of its SASS}: which machine code is in the .so
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
        if not os.path.exists(STAMP) and all(os.path.exists(_obj(u)) for u in UNITS):
            sass_stamp()
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
    sass_stamp()
    return SO


def kernel_sass_hashes(obj: str) -> dict:
    """sha256 (16 hex digits) of the SASS text of every kernel of an object file, keyed by its demangled name without the
    parameter list.  The anonymous-namespace tag (a hash of the source path) is normalised away, so the same code built in
    another directory hashes the same.  Evidence tooling: profiles/ncu_traffic.json quotes an ncu capture only for
    kernels whose machine code is exactly the captured one (bench.py:ncu_traffic)."""
    cuobjdump = os.path.join(os.path.dirname(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")), "cuobjdump")
    txt = subprocess.run([cuobjdump, "-sass", obj], capture_output=True, text=True, check=True).stdout
    funcs, name, buf = {}, None, []
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                funcs[name] = buf
            name, buf = m.group(1), []
        elif name is not None:
            buf.append(re.sub(r"_GLOBAL__N__[0-9a-f]{8}", "_GLOBAL__N__X", line.rstrip()))
    if name:
        funcs[name] = buf
    names = list(funcs)
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
    out = {}
    for mangled, d in zip(names, dem):
        d = re.sub(r"\(anonymous namespace\)::", "", d)
        d = re.sub(r"^void ", "", re.sub(r"\(.*$", "", d)).strip()
        out[d] = hashlib.sha256("\n".join(funcs[mangled]).encode()).hexdigest()[:16]
    return out


def sass_stamp() -> dict:
    """Write the stamp next to the .so (a build artefact like it: git-ignored, travels with it)."""
    try:
        stamp = {}
        for u in UNITS:
            if u.endswith(".cu"):
                stamp.update(kernel_sass_hashes(_obj(u)))
        with open(STAMP, "w") as f:
            json.dump(stamp, f, indent=1, sort_keys=True)
        return stamp
    except Exception as ex:  # noqa: BLE001 -- evidence tooling must never break the build
        print(f"nlps_b200.build: no SASS stamp ({ex})")
        return {}
