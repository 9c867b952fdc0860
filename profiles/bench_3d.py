"""3D explicit throughput (BASELINE configs[2] shape: Neo-Hookean cube, H8 grid, GPxElement 8, LME gamma 6).
Not the driver's bench line (bench.py measures configs[1]).  python profiles/bench_3d.py [cells] [steps] [gamma]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("nl-partsol_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
from nlps_b200 import engine, synthetic  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
gamma = float(sys.argv[3]) if len(sys.argv) > 3 else 6.0
t0 = time.perf_counter()
P = synthetic.cube_3d(cells=cells, nsteps=steps + 8, gamma_lme=gamma)
t1 = time.perf_counter()
eng = engine.Engine(P, device=0)
assert eng.initialize_lme() == 0, eng.error()
t2 = time.perf_counter()
assert eng.run(0, 3) == 0, eng.error()
rc, ms = eng.timed_run(3, steps)
assert rc == 0, eng.error()
n_avg = float(eng.download()["NumberNodes"].mean()) if P.np_ <= 4_000_000 else None
eng.profile(True)
eng.kernel_times(reset=True)
assert eng.run(3 + steps, 2) == 0
kt = {k: round(v[0] / v[1], 4) for k, v in eng.kernel_times().items() if v[1]}
alg = 1096 + 4 * (n_avg or 41.6)   # 3D Neo-Hookean bytes per particle per step
out = {"workload": f"3D NH cube {cells}^3 particle cells x 8, gamma {gamma}", "particles": P.np_, "nodes": P.nn,
       "ms_per_step": ms / steps, "particle_updates_per_s": P.np_ * steps / (ms * 1e-3),
       "neighbours_per_particle": n_avg, "step_alg_gbs": alg * P.np_ * steps / (ms * 1e-3) / 1e9,
       "step_frac_of_6552": alg * P.np_ * steps / (ms * 1e-3) / 1e9 / 6552.6, "kernel_ms": kt,
       "setup_s": {"problem": round(t1 - t0, 2), "engine+lme": round(t2 - t1, 2)}}
print(json.dumps(out))
eng.close()
