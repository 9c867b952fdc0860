"""Implicit Newmark-beta on the device against the CPU restatement, through the C ABI:
stage by stage (nodal v_n/a_n, initial guess, residual, tangent entries) and converged time steps.
The oracle solves its Newton systems with dense LU, the engine with Jacobi-PCG on a block-CSR tangent:
converged states are compared, with the tolerance the nonlinear solve supports."""
import numpy as np
import pytest

import oracle
from nlps_b200 import engine, synthetic
from util import assert_close, field_scales, load_problem

pytestmark = pytest.mark.gpu

def _golden(case):
    def make(n):
        P = load_problem(case)       # the 2D fixtures exported from the reference's own parser (tests/golden);
        assert P.solver["nsteps"] >= n   # their load curves are tabulated for the fixture's own step count: keep it
        return P
    return make


CASES = {
    "block2d": lambda n: synthetic.block_2d(cells=8, nsteps=n),
    "cube3d": lambda n: synthetic.cube_3d(cells=3, nsteps=n),
    "dp2d": _golden("dp"),
    "mn2d": _golden("mn"),
    "vm2d": _golden("vm"),
    "hencky2d": _golden("hencky"),
    "dp3d": lambda n: synthetic.cube_3d(cells=3, nsteps=n, material=synthetic.DP_C2),
}


def _pair(case, nsteps, cfl, tol=1e-12, explicit_trial=False):
    P = CASES[case](nsteps)
    P.solver["cfl"] = cfl if cfl > 0 else -cfl * P.solver["cfl"]     # negative: multiple of the case's own CFL
    o = oracle.Oracle(P)
    assert o.init_lme() == 0
    o.newmark_setup(tol=tol, max_iter=25, explicit_trial=explicit_trial)
    eng = engine.Engine(P, device=0)
    assert eng.initialize_lme() == 0
    assert eng.newmark_setup(tol=tol, max_iter=25, explicit_trial=explicit_trial, pcg_rtol=1e-13) == 0
    return P, o, eng


def dense_from_csr(eng, P, o, a1):
    rows, rp, cols, vals = eng.newmark_tangent()
    d = P.ndim
    nd = P.nn * d
    K = np.zeros((nd, nd))
    for t, A in enumerate(rows):
        for q in range(rp[t], rp[t + 1]):
            B = cols[q]
            K[A * d:(A + 1) * d, B * d:(B + 1) * d] += vals[q]
    K[np.arange(nd), np.arange(nd)] += a1 * o.nodal(0).ravel()
    dead = ((o.active()[:, None] == 0) | (o.fixed() != 0)).ravel()
    K[dead, :] = 0.0
    K[:, dead] = 0.0
    K[dead, dead] = 1.0
    return K, rows, rp


@pytest.mark.parametrize("case", ["block2d", "cube3d"])
def test_stages_match_the_oracle(case):
    P, o, eng = _pair(case, 4, 4.0, explicit_trial=True)
    # one converged step first, so that F_n != I and the state is not trivial
    assert o.newmark_step(0) == 0 and eng.newmark_step(0) == 0, eng.error()
    sc = field_scales(P)
    assert o.newmark_begin(1) == 0 and eng.newmark_begin(1) == 0
    assert np.array_equal(eng.active(), o.active())
    h, dt = P.dx, P.dt()
    for which, scale in (("Vn", h / dt), ("An", h / dt ** 2), ("dU", h)):
        assert_close(eng.newmark_get(which), o.newmark_get(which), f"{case} {which}", rtol=1e-9, scale=1e-5 * scale)
    dU = o.newmark_get("dU")
    st, Ro = o.newmark_residual(1, dU)
    rc, Rg = eng.newmark_residual(1, dU)
    assert st == 0 and rc == 0
    fscale = sc["gF"] / 1e-4                       # natural force scale E h^(d-1)
    assert np.abs(Rg - Ro).max() <= 1e-10 * max(np.abs(Ro).max(), 1e-4 * fscale)
    # tangent: every entry of the oracle's dense matrix
    st, Ko = o.newmark_tangent()
    beta = 0.25
    Kg, rows, rp = dense_from_csr(eng, P, o, 1.0 / (beta * dt * dt))
    assert st == 0
    assert np.abs(Kg - Ko).max() <= 1e-10 * np.abs(Ko).max()
    assert np.array_equal(rows, np.nonzero(o.active())[0])
    assert rp[-1] < len(rows) ** 2 or len(rows) < 40   # sparse
    eng.close()


@pytest.mark.parametrize("case,cfl", [("block2d", 5.0), ("cube3d", 5.0)])
def test_converged_steps_match_the_oracle(case, cfl):
    nsteps = 6
    P, o, eng = _pair(case, nsteps, cfl)
    for k in range(nsteps):
        assert o.newmark_step(k) == 0, o.error()
        assert eng.newmark_step(k) == 0, eng.error()
        st = eng.newmark_stats()
        assert st["newton_iters"] <= 8 and st["residual"] <= max(100 * 1e-12, 1e-12 * st["residual0"]) * 10
    f = eng.download()
    sc = field_scales(P)
    # both Newton loops stop at |R| <= 1e-12 |R0|: the states agree to the conditioning of the tangent times that
    for name in ("x_GC", "dis", "vel", "acc", "F_n", "Stress", "rho", "J_n", "lambda"):
        assert_close(f[name], o.field(name), f"{case} {name}", rtol=1e-8, scale=sc.get(name))
    assert np.array_equal(f["I0"], o.ints("I0"))
    counts, lists = eng.lists()
    assert np.array_equal(lists, o.lists())
    s = eng.newmark_stats()
    assert s["pcg_iters_total"] > 0 and s["assemblies_total"] >= nsteps
    eng.close()


@pytest.mark.parametrize("case,cfl,pre", [("dp2d", -4.0, 5), ("mn2d", -2.0, 2), ("dp3d", 2.0, 2), ("vm2d", -2.0, 3),
                                          ("hencky2d", -2.0, 2)])
def test_elastoplastic_stages_match_the_oracle(case, cfl, pre):
    """Drucker-Prager / Matsuoka-Nakai / Von-Mises tangent (compute_stiffness_elastoplastic__Constitutive__,
    Elastoplastic-Tangent-Matrix.c:42-160; Von-Mises with the moduli of its own __tangent_moduli, Von-Mises.c:730-757) and
    the Hencky block (Hencky.c:98-232): every block of the device CSR against the oracle's dense matrix, on a state
    reached by `pre` converged implicit steps (plastic for dp2d)."""
    P, o, eng = _pair(case, pre + 2, cfl, tol=1e-10)
    for k in range(pre):
        assert o.newmark_step(k) == 0, o.error()
        assert eng.newmark_step(k) == 0, eng.error()
    if case == "dp2d":
        assert (o.field("EPS_n") > 0).sum() > 50
    assert o.newmark_begin(pre) == 0 and eng.newmark_begin(pre) == 0
    dU = o.newmark_get("dU")
    st, Ro = o.newmark_residual(pre, dU)
    rc, Rg = eng.newmark_residual(pre, dU)
    assert st == 0 and rc == 0
    assert np.abs(Rg - Ro).max() <= 1e-7 * np.abs(Ro).max()     # the two histories differ by the Newton tolerance
    st, Ko = o.newmark_tangent()
    Kg, rows, rp = dense_from_csr(eng, P, o, 1.0 / (0.25 * P.dt() ** 2))
    assert st == 0
    assert np.abs(Kg - Ko).max() <= 1e-6 * np.abs(Ko).max()
    if case in ("dp2d", "dp3d"):       # non-associated flow: the operator is not symmetric, hence BiCGStab
        assert np.abs(Ko - Ko.T).max() > 1e-9 * np.abs(Ko).max()
    eng.close()


@pytest.mark.parametrize("case,cfl,nsteps", [("dp2d", -4.0, 8), ("mn2d", -2.0, 4), ("dp3d", 2.0, 3)])
def test_elastoplastic_converged_steps_match_the_oracle(case, cfl, nsteps):
    P, o, eng = _pair(case, nsteps, cfl, tol=1e-11)
    for k in range(nsteps):
        assert o.newmark_step(k) == 0, o.error()
        assert eng.newmark_step(k) == 0, eng.error()
    f = eng.download()
    sc = field_scales(P)
    # linear convergence (the reference's tangent is not exact): both loops stop at 1e-11 |R0|, the states agree to ~1e-7
    for name in ("x_GC", "dis", "vel", "F_n", "Stress", "J_n", "EPS_n", "b_e_n"):
        assert_close(f[name], o.field(name), f"{case} {name}", rtol=2e-6, scale=sc.get(name))
    assert np.array_equal(f["I0"], o.ints("I0"))
    s = eng.newmark_stats()
    assert s["pcg_iters_total"] > 0
    eng.close()


@pytest.mark.parametrize("case", ["block2d", "cube3d"])
def test_static_scheme_matches_the_oracle(case):
    """U_Static (Formulations/Displacements/U-Static.c:83-322): the implicit loop without inertia -- residual
    f_int - f_trac - M b, tangent K only, positions / history updated, velocities untouched.
    The Drucker-Prager deck is not parity-tested here: without the alpha_1 M term the nodes at the edge of the cloud carry
    almost no stiffness and the tangent is numerically singular (condition number 2.5e16 in the oracle's dense matrix):
    a dense LU and a converged Jacobi-BiCGStab differ by 1e-3 in the solution of the same system."""
    nsteps = 3 if case != "dp2d" else 1    # the plastic deck: one converged step (the reference's inexact elastoplastic
    P = CASES[case](nsteps)                # tangent stagnates on the next ones, in the oracle as on the device)
    if case == "dp2d":
        P.gravity = P.gravity * 0.01       # a load the column carries (a static limit load has no solution); still yields
    o = oracle.Oracle(P)
    assert o.init_lme() == 0
    o.static_setup(tol=1e-11, max_iter=30)
    eng = engine.Engine(P, device=0)
    assert eng.initialize_lme() == 0
    assert eng.newmark_setup(tol=1e-11, max_iter=30, pcg_rtol=1e-13, quasi_static=True) == 0
    v0 = np.array(P.fields["vel"], copy=True)
    for k in range(nsteps):
        assert o.newmark_step(k) == 0, o.error()
        assert eng.newmark_step(k) == 0, eng.error()
    f = eng.download()
    sc = field_scales(P)
    for name in ("x_GC", "dis", "F_n", "Stress", "rho", "J_n"):
        assert_close(f[name], o.field(name), f"static {case} {name}", rtol=2e-6 if case == "dp2d" else 1e-8, scale=sc.get(name))
    assert np.array_equal(f["vel"], v0) and np.abs(f["acc"]).max() == 0.0
    assert np.abs(f["dis"]).max() > 0.0
    assert np.array_equal(f["I0"], o.ints("I0"))
    eng.close()


def test_implicit_refuses_invalid_parameters():
    P = synthetic.block_2d(cells=4, nsteps=2)
    eng = engine.Engine(P, device=0)
    assert eng.initialize_lme() == 0
    assert eng.newmark_setup(beta=0.0) != 0          # a1 = 1/(beta dt^2)
    assert eng.newmark_step(0) != 0                  # no scheme was set up
    assert eng.newmark_setup() == 0 and eng.newmark_step(0) == 0
    eng.close()


@pytest.mark.parametrize("key,rtol", [("nh", 1e-8), ("nh_trial", 1e-8), ("dp", 2e-6), ("mn", 2e-6), ("static_nh", 1e-8),
                                      ("vm", 2e-6), ("hencky", 2e-6), ("nhload", 1e-8), ("static_nhload", 1e-8), ("mixed", 2e-6), ("almenh", 1e-8), ("almedp", 2e-6)])
def test_converged_steps_match_the_reference_compiled_scheme(key, rtol):
    """The device scheme against the reference's OWN U_Newmark_Beta / U_Static: tests/golden/newmark_*.npz hold the states
    the reference's compiled scheme code reached on 2D decks (run against oracle/minipetsc -- PETSc is absent --, see
    tests/golden/make_golden.py::gen_newmark; the CPU suite checks the oracle against the same files to 1e-12).
    Tolerances as in the oracle comparisons above: what a Newton loop stopped at the scheme's tolerance supports."""
    from util import newmark_golden, newmark_problem
    g = newmark_golden(key)
    k = int(max(g["checkpoints"]))
    P = newmark_problem(g, k)
    static = str(g["scheme"]) == "Static"
    eng = engine.Engine(P, device=0)
    assert eng.initialize_lme() == 0
    assert eng.newmark_setup(tol=float(g["tol"]), max_iter=int(g["max_iter"]), explicit_trial=bool(g["explicit_trial"]),
                             pcg_rtol=1e-13, quasi_static=static) == 0
    for s in range(k):
        assert eng.newmark_step(s) == 0, eng.error()
    f = eng.download()
    sc = field_scales(P)
    names = ("x_GC", "dis", "vel", "acc", "F_n", "Stress", "rho", "J_n", "lambda") if rtol <= 1e-8 else \
            ("x_GC", "dis", "vel", "F_n", "Stress", "J_n") + (("EPS_n", "b_e_n") if key != "hencky" else ())
    if key.startswith("alme"):      # GramsShapeFun (Type=aLME): the convected metric and cut-off ellipsoid
        names = names + ("Beta", "Cut_off_Ellipsoid")
    for name in names:
        assert_close(f[name], g[f"s{k}_{name}"], f"reference scheme {key} {name}", rtol=rtol, scale=sc.get(name))
    assert np.array_equal(f["I0"], g[f"s{k}_I0"])
    counts, _ = eng.lists()
    assert np.array_equal(counts, g[f"s{k}_NumberNodes"])
    if key in ("dp", "mixed", "almedp"):
        assert (f["EPS_n"] > 0).sum() > 20
    eng.close()
