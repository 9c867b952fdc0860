"""Parity of the CUDA engine (through the C ABI) with the oracle port and with the golden fixtures
frozen from the reference's own compiled 2D code.  Needs a B200: `pytest -m gpu`."""
import numpy as np
import pytest

import oracle
from nlps_b200 import engine
from util import (NODAL, RTOL, TRACE_FIELDS, assert_close, field_scales, load_points, load_problem,
                  load_trace)

pytestmark = pytest.mark.gpu
CASES = ("nh", "dp", "mn")


def _close_nodal(eng, o, P, what):
    for w, nm in enumerate(NODAL):
        assert_close(eng.nodal(w), o.nodal(w), f"{what} nodal {nm}")


@pytest.mark.parametrize("case", CASES)
def test_initialize_lme_matches_reference(case):
    P = load_problem(case)
    P0 = P.copy()
    P0.fields["Beta"][:] = 0.0
    P0.fields["lambda"][:] = 0.0
    eng = engine.Engine(P0)
    assert eng.initialize_lme() == 0, eng.error()
    f = eng.download()
    assert np.array_equal(f["Beta"], P.fields["Beta"])          # bit-exact (one division)
    assert_close(f["lambda"], P.fields["lambda"], "lambda", scale=1e-3 / P.dx)
    o = oracle.Oracle(P0)
    assert o.init_lme() == 0
    counts, lists = eng.lists()
    assert np.array_equal(counts, o.ints("NumberNodes"))
    assert np.array_equal(lists, o.lists())
    assert np.array_equal(eng.active(), o.active())
    eng.close()


@pytest.mark.parametrize("case", CASES)
def test_steps_match_golden_reference(case):
    """Multi-step run against fixtures produced by the reference's compiled code."""
    P = load_problem(case)
    tr = load_trace(case)
    eng = engine.Engine(P, compute_c_ep=1)
    scales = field_scales(P)
    cps = [int(c) for c in tr["checkpoints"]]
    done = 0
    for cp in cps:
        assert eng.run(done, cp - done) == 0, (cp, eng.error())
        done = cp
        t = f"s{cp}_"
        f = eng.download()
        counts, lists = eng.lists()
        assert np.array_equal(f["I0"], tr[t + "I0"]), f"I0 at step {cp}"
        assert np.array_equal(counts, tr[t + "NumberNodes"]), f"NumberNodes at step {cp}"
        assert np.array_equal(lists[:, :tr[t + "lists"].shape[1]], tr[t + "lists"]), f"lists at step {cp}"
        assert np.array_equal(eng.active(), tr[t + "active"])
        for name in TRACE_FIELDS:
            assert_close(f[name], tr[t + name], f"{case} step {cp} {name}", scale=scales.get(name))
        for w, nm in enumerate(NODAL):
            assert_close(eng.nodal(w), tr[t + "g" + nm], f"{case} step {cp} nodal {nm}")
    eng.close()


@pytest.mark.parametrize("case", CASES)
def test_stagewise_against_oracle(case):
    """Every stage from IDENTICAL state: the oracle state is re-uploaded before each step."""
    P = load_problem(case)
    o = oracle.Oracle(P)
    eng = engine.Engine(P, compute_c_ep=1)
    scales = field_scales(P)
    for k in range(6):
        eng.upload({name: o.field(name) for name in TRACE_FIELDS + ("F_n1", "J_n1", "b_e_n1", "EPS_n1", "Kappa_n1")
                    if name not in ("J_n",)} | {"J_n": o.field("J_n")})
        # search
        assert o.stage("search", k) == 0 and eng.stage("search", k) == 0
        f = eng.download()
        counts, lists = eng.lists()
        assert np.array_equal(f["I0"], o.ints("I0"))
        assert np.array_equal(counts, o.ints("NumberNodes")) and np.array_equal(lists, o.lists())
        assert np.array_equal(eng.active(), o.active())
        assert np.array_equal(f["Beta"], o.field("Beta"))
        assert_close(f["lambda"], o.field("lambda"), f"step {k} lambda", scale=scales["lambda"])
        # P2G mass / displacement increment + grid update
        for st in ("p2g_mass_disp", "grid_disp"):
            assert o.stage(st, k) == 0 and eng.stage(st, k) == 0
        for w in (0, 1):
            assert_close(eng.nodal(w), o.nodal(w), f"step {k} nodal {NODAL[w]}")
        # kinematics + stress
        assert o.stage("kin_stress", k) == 0 and eng.stage("kin_stress", k) == 0, eng.error()
        f = eng.download()
        for name in ("DF", "F_n1", "J_n1", "rho", "Stress", "W", "b_e_n1", "EPS_n1", "Kappa_n1", "C_ep"):
            assert_close(f[name], o.field(name), f"step {k} {name}", scale=scales.get(name))
        # forces + nodal equilibrium
        for st in ("force", "grid_acc"):
            assert o.stage(st, k) == 0 and eng.stage(st, k) == 0
        for w in (2, 3, 4):
            assert_close(eng.nodal(w), o.nodal(w), f"step {k} nodal {NODAL[w]}")
        # G2P + corrector
        assert o.stage("g2p", k) == 0 and eng.stage("g2p", k) == 0
        f = eng.download()
        for name in TRACE_FIELDS:
            assert_close(f[name], o.field(name), f"step {k} {name} after g2p", scale=scales.get(name))
    eng.close()


@pytest.mark.parametrize("case", ("dp", "mn"))
def test_material_points_match_reference(case):
    """Constitutive update on the strain paths frozen from the reference (incl. its own test path)."""
    z = load_points(case)
    X, Y = z["inputs"], z["outputs"]
    r = engine.stress_points(2, str(z["mat_type"]), z["mat_params"], float(z["tol_radial"]),
                             int(z["maxiter_radial"]), X[:, 0:5], X[:, 5:10], X[:, 10], X[:, 11:16], X[:, 16],
                             X[:, 17])
    assert np.all(r["status"] == 0)
    got = np.concatenate([r["stress"], r["b_e_n1"], r["eps_n1"][:, None], r["kappa_n1"][:, None], r["W"][:, None],
                          r["C_ep"]], axis=1)
    for sl, nm in ((slice(0, 5), "stress"), (slice(5, 10), "b_e"), (slice(10, 11), "eps"), (slice(11, 12), "kappa"),
                   (slice(12, 13), "W"), (slice(13, 17), "C_ep")):
        s = np.maximum(np.abs(Y[:, sl]).max(axis=1, keepdims=True), 1e-9)
        fin = np.isfinite(Y[:, sl])
        assert np.array_equal(np.isfinite(got[:, sl]), fin)
        e = np.where(fin, np.abs(got[:, sl] - Y[:, sl]) / s, 0.0).max()
        assert e <= RTOL, (nm, e)


def test_error_latch_negative_jacobian():
    """Device-side failure is surfaced through the reference's EXIT_FAILURE convention."""
    P = load_problem("nh")
    P.fields["vel"][:, 1] = -1.0e4 * (P.fields["x_GC"][:, 1] + 0.1)   # violent compression
    eng = engine.Engine(P)
    rc = 0
    for k in range(20):
        rc = eng.step(k)
        if rc:
            break
    assert rc == 1 and eng.error()[0] in (2, 3, 4, 5)
    eng.close()


def test_u_verlet_host_call_matches_engine():
    P = load_problem("nh")
    P.solver["nsteps"] = 12
    for b in P.bounds:
        b["dir"], b["val"] = b["dir"][:, :12], b["val"][:, :12]
    P.gravity = P.gravity[:, :12]
    f = engine.u_verlet(P, results_every=5)
    eng = engine.Engine(P)
    assert eng.run(0, 12) == 0
    g = eng.download()
    for name in ("x_GC", "vel", "Stress", "F_n"):
        assert np.array_equal(f[name], g[name]), name
    eng.close()
