"""Timing of the implicit Newmark-beta path (BASELINE configs[4] shape: 3D cantilever, Neo-Hookean,
LME gamma 6, dt = 10 x the explicit limit).  Not the driver's bench line (bench.py measures configs[1]);
prints one JSON line with particle-updates/s (steps = converged time steps), Newton / PCG counts and the
split assembly / PCG / residual.   python profiles/bench_implicit.py [cells_per_unit] [steps]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("nl-partsol_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
from nlps_b200 import engine, synthetic  # noqa: E402

c = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
t0 = time.perf_counter()
P = synthetic.beam_3d(cells_per_unit=c, nsteps=steps + 1)
t1 = time.perf_counter()
eng = engine.Engine(P, device=0)
assert eng.initialize_lme() == 0
t2 = time.perf_counter()
assert eng.newmark_setup(tol=1e-10, max_iter=10, pcg_rtol=float(os.environ.get("PCG_RTOL", "1e-6"))) == 0
t3 = time.perf_counter()
assert eng.newmark_step(0) == 0, eng.error()       # warm-up step (first launches, pattern)
s0 = eng.newmark_stats()
import torch  # noqa: E402  (only for the synchronize below)
torch.cuda.synchronize()
ta = time.perf_counter()
newton = []
for k in range(1, steps + 1):
    assert eng.newmark_step(k) == 0, eng.error()
    newton.append(eng.newmark_stats()["newton_iters"])
torch.cuda.synchronize()
tb = time.perf_counter()
s = eng.newmark_stats()
nnz_bytes = s["nnz_blocks"] * 76 + 5 * 3 * 8 * s["n_rows"]   # 72 B of values + one column index per 3x3 block, 5 vectors
pcg = s["pcg_iters_total"] - s0["pcg_iters_total"]
out = {"workload": f"3D cantilever 8x1x1, NH, gamma 6, implicit Newmark-beta, cfl {P.solver['cfl']}",
       "particles": P.np_, "nodes": P.nn, "steps": steps, "s_per_step": (tb - ta) / steps,
       "particle_updates_per_s": P.np_ * steps / (tb - ta), "newton_iters": newton, "pcg_iters": pcg,
       "rows": s["n_rows"], "nnz_blocks": s["nnz_blocks"],
       "ms_assemble_per_newton": (s["ms_assemble"] - s0["ms_assemble"]) / max(1, s["assemblies_total"] - s0["assemblies_total"]),
       "ms_per_pcg_iter": (s["ms_pcg"] - s0["ms_pcg"]) / max(1, pcg),
       "pcg_iteration_gbs": nnz_bytes / 1e9 / (1e-3 * (s["ms_pcg"] - s0["ms_pcg"]) / max(1, pcg)),
       "ms_residual_per_eval": (s["ms_residual"] - s0["ms_residual"]) / max(1, s["residual_evals_total"] - s0["residual_evals_total"]),
       "setup_s": {"problem": t1 - t0, "engine+lme": t2 - t1, "coupling_adjacency": t3 - t2}}
print(json.dumps(out))
eng.close()
