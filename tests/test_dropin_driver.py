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


def _particle_vtk(d, k):
    """particle results files of step k (the nodal file Nodes_k.vtk lies beside them)"""
    return [f for f in glob.glob(os.path.join(str(d), "Results", f"*_{k}.vtk"))
            if not os.path.basename(f).startswith("Nodes_")]


def _points(vtk):
    lines = open(vtk).read().splitlines()
    i = next(k for k, l in enumerate(lines) if l.startswith("POINTS"))
    n = int(lines[i].split()[1])
    return np.array([[float(v) for v in lines[i + 1 + k].split()[:2]] for k in range(n)])


@pytest.mark.gpu
@pytest.mark.parametrize("case", ("nh", "vm", "almenh"))
def test_reference_driver_with_b200_scheme(tmp_path, case):
    """nh: the Neo-Hookean block; vm: a Von-Mises deck (mixed isotropic / kinematic hardening: the back stress lives in
    the reference's own Phi.Back_stress buffer and is handed to the engine by the shim); almenh: GramsShapeFun (Type=aLME)
    -- the reference's driver initialises metric and ellipsoid (initialize__aLME__), the shim hands Particle.Beta (n x 4)
    and Particle.Cut_off_Ellipsoid to the engine."""
    if not os.path.exists(BIN):
        pytest.skip("drop-in binary not built (needs /root/reference at build time)")
    import deckgen
    import make_golden
    from util import load_trace
    spec = make_golden.spec_for(case)
    spec.out_every = 1
    deckgen.write_deck(spec, str(tmp_path))
    r = subprocess.run([BIN, "--FORMULATION-U", "-f", "deck.nlp"], cwd=str(tmp_path), capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "abnormally" not in r.stdout + r.stderr
    tr = load_trace(case)
    for cp in [int(c) for c in tr["checkpoints"] if c <= 60]:
        files = _particle_vtk(tmp_path, cp - 1)
        assert files, f"no VTK for step {cp - 1}"
        x = _points(files[0])
        assert np.abs(x - tr[f"s{cp}_x_GC"]).max() <= 1e-10 * np.abs(tr[f"s{cp}_x_GC"]).max()
        # the nodal file of the same step (nodal_results_vtk__InOutFun__: ActiveNodes mask + reactions, %.20g)
        nodal = os.path.join(str(tmp_path), "Results", f"Nodes_{cp - 1}.vtk")
        assert os.path.exists(nodal), "no nodal VTK file"
        mask, R = _nodal(nodal)
        assert np.array_equal(mask, tr[f"s{cp}_active"].astype(np.int64))
        gR = tr[f"s{cp}_gR"]
        assert np.abs(R[:, :2] - gR).max() <= 1e-10 * max(np.abs(gR).max(), 1e-300)


def _nodal(vtk):
    lines = open(vtk).read().splitlines()
    i = next(k for k, l in enumerate(lines) if l.startswith("POINT_DATA"))
    n = int(lines[i].split()[1])
    j = next(k for k, l in enumerate(lines) if l.startswith("LOOKUP_TABLE"))
    mask = np.array([int(lines[j + 1 + k]) for k in range(n)])
    r = next(k for k, l in enumerate(lines) if l.startswith("VECTORS REACTIONS"))
    R = np.array([[float(v) for v in lines[r + 1 + k].split()[:3]] for k in range(n)])
    return mask, R


@pytest.mark.gpu
def test_reference_driver_binary_vtk_output(tmp_path):
    """NLPS_B200_VTK_BINARY=1: the shim writes the results steps with the binary twin of the reference's writer
    (host/b200_vtk_binary.h), overlapped with the following steps; positions and stresses of the plastic
    deck against the golden trace."""
    if not os.path.exists(BIN):
        pytest.skip("drop-in binary not built (needs /root/reference at build time)")
    import deckgen
    import make_golden
    import vtkio
    from util import load_trace
    spec = make_golden.spec_for("dp")
    spec.out_every = 1
    deckgen.write_deck(spec, str(tmp_path))
    r = subprocess.run([BIN, "--FORMULATION-U", "-f", "deck.nlp"], cwd=str(tmp_path), capture_output=True, text=True,
                       timeout=600, env=dict(os.environ, NLPS_B200_VTK_BINARY="1"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    tr = load_trace("dp")
    for cp in (1, 5, 60, 120):
        files = _particle_vtk(tmp_path, cp - 1)
        assert files, f"no VTK for step {cp - 1}"
        assert b"BINARY" in open(files[0], "rb").read(200)
        v = vtkio.read_binary(files[0])
        x, ref = v["POINTS"][:, :2], tr[f"s{cp}_x_GC"]
        assert np.abs(x - ref).max() <= 1e-10 * np.abs(ref).max()
        if "STRESS" in v:
            s_ref = tr[f"s{cp}_Stress"]
            got = v["STRESS"][:, [0, 1, 3, 4, 8]]
            assert np.abs(got - s_ref).max() <= 1e-9 * max(np.abs(s_ref).max(), 1.0)


@pytest.mark.gpu
def test_reference_driver_with_b200_implicit_scheme(tmp_path):
    """`NLPS-Solver (Type=Newmark-beta-Finite-Strains)`: the reference driver (compiled with -DUSE_PETSC against
    stand-in headers, no PETSc library) dispatches to U_Newmark_Beta, which here is the B200 shim.  The VTK
    positions are compared with the CPU restatement of the scheme (oracle, dense LU) on the same problem."""
    if not os.path.exists(BIN):
        pytest.skip("drop-in binary not built (needs /root/reference at build time)")
    import deckgen
    import make_golden
    import oracle
    from util import load_problem
    nsteps, cfl, tol = 6, 4.0, 1e-12
    spec = make_golden.spec_for("nh")
    spec.scheme = "Newmark-beta-Finite-Strains"
    spec.nsteps, spec.cfl, spec.out_every = nsteps, cfl, 1
    spec.solver_extra = {"Beta-Newmark-beta": 0.25, "Gamma-Newmark-beta": 0.5, "TOL-Newmark-beta": tol, "Max-Iter": 25,
                         "Epsilon": 0.0}
    deckgen.write_deck(spec, str(tmp_path))
    r = subprocess.run([BIN, "--FORMULATION-U", "-f", "deck.nlp"], cwd=str(tmp_path), capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "abnormally" not in r.stdout + r.stderr
    P = load_problem("nh")                      # the same deck as the reference's own setup produced it
    P.solver["cfl"], P.solver["nsteps"] = cfl, nsteps
    for b in P.bounds:
        b["dir"], b["val"] = b["dir"][:, :nsteps], b["val"][:, :nsteps]
    P.gravity = P.gravity[:, :nsteps]
    o = oracle.Oracle(P)
    assert o.init_lme() == 0
    o.newmark_setup(tol=tol, max_iter=25)
    for k in range(nsteps):
        assert o.newmark_step(k) == 0, o.error()
        files = _particle_vtk(tmp_path, k)
        assert files, f"no VTK for step {k}"
        x = _points(files[0])
        assert np.abs(x - o.field("x_GC")).max() <= 1e-8 * np.abs(o.field("x_GC")).max(), k
    assert np.abs(o.field("dis")).max() > 1e-6


@pytest.mark.gpu
def test_reference_driver_with_b200_static_scheme(tmp_path):
    """`NLPS-Solver (Type=Static)`: driver-nl-partsol.c:373-375 dispatches to U_Static, here the B200 shim (the implicit
    engine without inertia).  VTK positions against the CPU restatement."""
    if not os.path.exists(BIN):
        pytest.skip("drop-in binary not built (needs /root/reference at build time)")
    import deckgen
    import make_golden
    import oracle
    from util import load_problem
    nsteps, tol = 3, 1e-11
    spec = make_golden.spec_for("nh")
    spec.scheme = "Static"
    spec.nsteps, spec.out_every = nsteps, 1
    spec.solver_extra = {"TOL-Newmark-beta": tol, "Max-Iter": 25, "Epsilon": 0.0}
    deckgen.write_deck(spec, str(tmp_path))
    r = subprocess.run([BIN, "--FORMULATION-U", "-f", "deck.nlp"], cwd=str(tmp_path), capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "abnormally" not in r.stdout + r.stderr
    P = load_problem("nh")
    P.solver["nsteps"] = nsteps
    for b in P.bounds:
        b["dir"], b["val"] = b["dir"][:, :nsteps], b["val"][:, :nsteps]
    P.gravity = P.gravity[:, :nsteps]
    o = oracle.Oracle(P)
    assert o.init_lme() == 0
    o.static_setup(tol=tol, max_iter=25)
    for k in range(nsteps):
        assert o.newmark_step(k) == 0, o.error()
        files = _particle_vtk(tmp_path, k)
        assert files, f"no VTK for step {k}"
        x = _points(files[0])
        assert np.abs(x - o.field("x_GC")).max() <= 1e-8 * np.abs(o.field("x_GC")).max(), k
    assert np.abs(o.field("dis")).max() > 1e-6


@pytest.mark.gpu
def test_reference_driver_on_two_gpus_from_the_c_host(tmp_path):
    """NLPS_B200_GPUS=2: the U_Verlet shim (C, no torch) cuts the cloud into two slabs, runs one slab engine per device
    from two host threads (NCCL id made in C) and merges the rows into the reference's buffers; the VTK files of the
    run agree with those of the one-GPU run of the same deck to 1e-10."""
    import torch
    if not os.path.exists(BIN):
        pytest.skip("drop-in binary not built (needs /root/reference at build time)")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import deckgen
    spec = deckgen.DeckSpec(nx=48, ny=14, pnx=40, pny=8, porigin=(4 * 0.0625, 0.0), nsteps=40, out_every=20)
    runs = {}
    for tag, env in (("one", {}), ("two", {"NLPS_B200_GPUS": "2"})):
        d = tmp_path / tag
        d.mkdir()
        deckgen.write_deck(spec, str(d))
        r = subprocess.run([BIN, "--FORMULATION-U", "-f", "deck.nlp"], cwd=str(d), capture_output=True, text=True,
                           timeout=600, env=dict(os.environ, **env))
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        assert "abnormally" not in r.stdout + r.stderr
        runs[tag] = {k: _points(_particle_vtk(d, k)[0]) for k in (0, 20)}
    for k in (0, 20):
        a, b = runs["one"][k], runs["two"][k]
        assert a.shape == b.shape and np.abs(a - b).max() <= 1e-10 * np.abs(a).max(), k
    assert np.abs(runs["one"][20] - runs["one"][0]).max() > 1e-9       # the block did move


@pytest.mark.gpu
def test_reference_driver_scalable_setup_at_100k_particles(tmp_path):
    """102,400 particles through the reference's OWN driver: its quadratic set-up (get_sourrounding_elements,
    Read_GramsBox.c:293-330, and the element scan of initialize__LME__, LME.c:63-108: 69 s at this size on 8 host threads)
    is replaced at link time by host/Setup-b200.c (1.4 s), the first lists / beta / lambda come from the engine.  The
    positions after 21 steps agree with the CPU oracle on the same cloud to 1e-10."""
    if not os.path.exists(BIN):
        pytest.skip("drop-in binary not built (needs /root/reference at build time)")
    import time

    import deckgen
    import oracle
    from nlps_b200 import synthetic
    n, nsteps = 160, 21
    spec = deckgen.DeckSpec(nx=n + 10, ny=n + 10, h=1.0 / n, pnx=n, pny=n, ph=1.0 / n, porigin=(5.0 / n, 0.0), nsteps=nsteps,
                            cfl=0.5, out_every=10)
    deckgen.write_deck(spec, str(tmp_path))
    t0 = time.perf_counter()
    r = subprocess.run([BIN, "--FORMULATION-U", "-f", "deck.nlp"], cwd=str(tmp_path), capture_output=True, text=True,
                       timeout=900)
    wall = time.perf_counter() - t0
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "abnormally" not in r.stdout + r.stderr
    assert wall < 60.0, f"the whole driver run took {wall:.1f} s: the quadratic set-up is back"
    P = synthetic.structured_problem(2, (n + 10, n + 10), 1.0 / n, (n, n), (5, 0), synthetic.NH_C1, nsteps, 0.5, spec.cel,
                                     (0.0, -9.81))
    o = oracle.Oracle(P, threads=os.cpu_count() or 1)
    assert o.init_lme() == 0
    for k in range(nsteps):
        assert o.step(k) == 0, o.error()
    x = _points(_particle_vtk(tmp_path, 20)[0])
    ref = o.field("x_GC")
    assert x.shape == ref.shape and np.abs(x - ref).max() <= 1e-10 * np.abs(ref).max()
