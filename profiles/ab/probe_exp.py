"""Timing probe (profiles/ab/libprobe.so, built from a patched copy of nlps_cellwarp.cu -- see profiles/r02/session17.sh):
what would the 3D kinematics / G2P kernels cost without the exponential (bit 0) and without the shared-memory loads of
the node coordinates (bit 1)?  Three correct steps, then ONE step with the probe on (its results are wrong on purpose)
with per-kernel CUDA events."""
import ctypes
import json
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path[:0] = [os.path.join(ROOT, "nl-partsol_b200")]
from nlps_b200 import engine, synthetic  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 64
out = {}
for probe in (0, 1, 3):
    P = synthetic.cube_3d(cells=cells, nsteps=8)
    eng = engine.Engine(P)
    assert eng.initialize_lme() == 0
    assert eng.run(0, 3) == 0, eng.error()
    eng.profile(True)
    eng.kernel_times(reset=True)
    L = engine.lib()
    assert L.nlps_probe_set(probe) == 0
    eng.run(3, 1)                      # may latch an error with the probe on: only the kernel times are read
    kt = eng.kernel_times()
    L.nlps_probe_set(0)
    out[probe] = {k: round(ms / max(n, 1), 4) for k, (ms, n) in kt.items() if n and ms / n > 0.1}
    try:
        eng.close()
    except Exception:
        pass
print(json.dumps(out))
