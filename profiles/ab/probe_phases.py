"""Phase-elimination probe of cw_kin (profiles/ab/libprobe.so from profiles/ab/r02_probe_phases.patch): one phase of the
kernel switched off at a time -- 4 force scatter loop, 8 gather loop, 16 stress update, 32 flush of the cell sums,
64 mask union + node staging, 128 the whole particle part -- on the 64^3 cube after three correct steps."""
import json
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path[:0] = [os.path.join(ROOT, "nl-partsol_b200")]
from nlps_b200 import engine, synthetic  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 64
out = {}
for probe in [int(v) for v in sys.argv[2:]]:   # one process per probe value: a probed kernel may fault
    P = synthetic.cube_3d(cells=cells, nsteps=8)
    eng = engine.Engine(P)
    assert eng.initialize_lme() == 0
    assert eng.run(0, 3) == 0, eng.error()
    eng.profile(True)
    eng.kernel_times(reset=True)
    L = engine.lib()
    assert L.nlps_probe_set(probe) == 0
    eng.run(3, 1)
    kt = eng.kernel_times()
    L.nlps_probe_set(0)
    ms, n = kt["kin_stress_p2g_force"]
    out[probe] = round(ms / max(n, 1), 4)
    try:
        eng.close()
    except Exception:
        pass
print(json.dumps(out))
