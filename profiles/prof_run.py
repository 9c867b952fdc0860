"""Short run of the bench workload for ncu (BASELINE configs[1] at full size, few steps)."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path[:0] = [os.path.join(ROOT, "nl-partsol_b200")]
from nlps_b200 import engine, synthetic
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
P = synthetic.column_collapse_2d(scale=scale, nsteps=steps + 1)
eng = engine.Engine(P)
assert eng.initialize_lme() == 0
assert eng.run(0, steps) == 0, eng.error()
print("ok", P.np_, "particles", eng.launch_count(), "launches")
eng.close()
