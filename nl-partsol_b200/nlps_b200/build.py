"""Build libnlps_b200.so (hand-written sm_100a CUDA + host setup) in-tree with nvcc."""
from __future__ import annotations

import os
import subprocess

PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(PKG, "libnlps_b200.so")
SRCS = [os.path.join(PKG, "csrc", f) for f in ("nlps_engine.cu", "host_setup.cpp")]
DEPS = SRCS + [os.path.join(PKG, "csrc", "nlps_device.cuh"), os.path.join(PKG, "csrc", "nlps_implicit.inl"),
               os.path.join(PKG, "..", "include", "nlps_b200.h")]


def stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(f) > t for f in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fPIC,-fopenmp,-O2", "-shared", "-o", SO] + SRCS + ["-lgomp", "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    return SO
