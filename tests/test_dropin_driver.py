"""The reference's OWN driver binary with the B200 scheme shim linked in place of U-Verlet.c
(nl-partsol_b200/host/Makefile -> oracle/_ref/nl-partsol-b200): a deck goes in, the reference's parser
and VTK writer run unchanged, the steps run on the GPU.  The lossless (%.20g) VTK particle positions
are compared with the golden trace frozen from the reference's CPU code."""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
BIN = os.path.join(ROOT, "oracle", "_ref", "nl-partsol-b200")
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def _points(vtk):
    lines = open(vtk).read().splitlines()
    i = next(k for k, l in enumerate(lines) if l.startswith("POINTS"))
    n = int(lines[i].split()[1])
    return np.array([[float(v) for v in lines[i + 1 + k].split()[:2]] for k in range(n)])


@pytest.mark.gpu
def test_reference_driver_with_b200_scheme(tmp_path):
    if not os.path.exists(BIN):
        pytest.skip("drop-in binary not built (needs /root/reference at build time)")
    import deckgen
    import make_golden
    from util import load_trace
    spec = make_golden.spec_for("nh")
    spec.out_every = 1
    deckgen.write_deck(spec, str(tmp_path))
    r = subprocess.run([BIN, "--FORMULATION-U", "-f", "deck.nlp"], cwd=str(tmp_path), capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "abnormally" not in r.stdout + r.stderr
    tr = load_trace("nh")
    for cp in (1, 2, 5, 20, 60):
        files = glob.glob(os.path.join(str(tmp_path), "Results", f"*_{cp - 1}.vtk"))
        assert files, f"no VTK for step {cp - 1}"
        x = _points(files[0])
        assert np.abs(x - tr[f"s{cp}_x_GC"]).max() <= 1e-10 * np.abs(tr[f"s{cp}_x_GC"]).max()
