"""the binary twin of the reference's particle VTK writer (nl-partsol_b200/host/b200_vtk_binary.h) against
the reference's own particle_results_vtk__InOutFun__ (InOutFun/Outputs/WriteVtk.c:95-268) on the same state, both run
inside oracle/_ref on a deck stepped by the reference's stage functions.  CPU only; needs oracle/_ref (this container)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import refharness
import vtkio

HERE = os.path.dirname(os.path.abspath(__file__))

WORKER = r"""
import ctypes, os, sys
sys.path[:0] = [{here!r}, os.path.join({here!r}, "golden"), os.path.join({here!r}, "..", "nl-partsol_b200")]
import deckgen, make_golden, refharness
case, out = sys.argv[1], sys.argv[2]
h = refharness.RefHarness(deckgen.write_deck(make_golden.spec_for(case), out), threads=1)
for k in range(12):
    assert h.step(k) == 0
L = h.lib
L.refh_set_outputs(1)
os.makedirs(os.path.join(out, "a"), exist_ok=True)
os.makedirs(os.path.join(out, "b"), exist_ok=True)
assert L.refh_write_vtk(11, 4, 0, os.path.join(out, "a").encode(), b"particles") == 0
assert L.refh_write_vtk(11, 4, 1, os.path.join(out, "b").encode(), b"particles") == 0
"""


@pytest.mark.parametrize("case", ["nh", "dp"])
def test_binary_writer_matches_the_reference_writer(case, tmp_path):
    if not refharness.available():
        pytest.skip("oracle/_ref not built")
    # the reference keeps its state in process globals: fresh interpreter per deck
    r = subprocess.run([sys.executable, "-c", WORKER.format(here=HERE), case, str(tmp_path)], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    a = vtkio.read_ascii(os.path.join(str(tmp_path), "a", "particles_11.vtk"))
    b = vtkio.read_binary(os.path.join(str(tmp_path), "b", "particles_11.vtk"))
    assert set(a) == set(b), (sorted(a), sorted(b))
    for k in a:
        assert a[k].shape == b[k].shape, k
        assert np.array_equal(a[k], b[k]), k            # "%.20g" round-trips a double exactly
    for k in ("POINTS", "X_GC", "MASS", "DENSITY", "ELEM_i", "MatIdx", "VELOCITY", "ACCELERATION", "DISPLACEMENT", "STRESS",
              "P", "DEFORMATION-GRADIENT", "Energy-Potential", "Energy-Kinetic", "EPS"):
        assert k in b, k
    assert np.abs(b["STRESS"]).max() > 0 and np.abs(b["STRESS"][:, 8]).max() > 0     # plane strain: sigma_33 from slot 4
