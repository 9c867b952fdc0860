"""Short run of the 3D cube (BASELINE configs[2] shape at 64^3) for ncu."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path[:0] = [os.path.join(ROOT, "nl-partsol_b200")]
from nlps_b200 import engine, synthetic
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
P = synthetic.cube_3d(cells=cells, nsteps=steps + 1)
eng = engine.Engine(P)
assert eng.initialize_lme() == 0
assert eng.run(0, steps) == 0, eng.error()
print("ok", P.np_, "particles", eng.launch_count(), "launches")
eng.close()
