"""CPU checks of the implicit Newmark-beta restatement in oracle/.  The scheme is pinned in 2D to the reference's own
compiled U-Newmark-beta.c / U-Static.c (run against oracle/minipetsc because PETSc is absent; fixtures
tests/golden/newmark_*.npz), its per-particle tangent blocks to the reference's own compiled functions
(tests/golden/tangent_blocks.npz).  Beyond that: the tangent is the derivative
of the residual, the trapezoidal scheme converges to the explicit oracle as dt -> 0, Newton converges quadratically."""
import os

import numpy as np
import pytest

import oracle
from nlps_b200 import synthetic


def _implicit(P, tol=1e-12):
    o = oracle.Oracle(P)
    assert o.init_lme() == 0
    o.newmark_setup(tol=tol, max_iter=25)
    return o


def test_tangent_is_the_derivative_of_the_residual():
    P = synthetic.block_2d(cells=6, nsteps=4)
    P.solver["cfl"] = 4.0
    o = _implicit(P)
    assert o.newmark_step(0) == 0           # a deformed state to linearise about
    assert o.newmark_begin(1) == 0
    dU = o.newmark_get("dU")
    rng = np.random.default_rng(7)
    free = (o.active()[:, None] > 0) & (o.fixed() == 0)
    dU = dU + 1e-4 * P.dx * rng.standard_normal(dU.shape) * free
    st, R0 = o.newmark_residual(1, dU)
    assert st == 0
    st, K = o.newmark_tangent()
    assert st == 0 and np.abs(K - K.T).max() <= 1e-9 * np.abs(K).max()
    v = rng.standard_normal(dU.shape) * free
    eps = 1e-6 * P.dx
    _, Rp = o.newmark_residual(1, dU + eps * v)
    _, Rm = o.newmark_residual(1, dU - eps * v)
    fd = (Rp - Rm).ravel() / (2 * eps)
    an = K @ v.ravel()
    m = free.ravel()
    assert np.abs(fd[m] - an[m]).max() <= 2e-6 * np.abs(an[m]).max()


def test_implicit_matches_explicit_for_small_dt():
    errs = []
    for cfl, n in ((0.1, 20), (0.05, 40)):
        P = synthetic.block_2d(cells=6, nsteps=n)
        P.solver["cfl"] = cfl
        oe = oracle.Oracle(P)
        assert oe.init_lme() == 0
        oi = _implicit(P)
        for k in range(n):
            assert oe.step(k) == 0 and oi.newmark_step(k) == 0
        errs.append(np.abs(oe.field("dis") - oi.field("dis")).max() / np.abs(oe.field("dis")).max())
    assert errs[0] < 2e-2 and errs[1] < 0.6 * errs[0]   # first order or better in dt between the two schemes


def test_newton_converges_at_ten_times_the_explicit_step():
    P = synthetic.cube_3d(cells=3, nsteps=3)
    P.solver["cfl"] = 5.0                    # cel carries a factor 1.3: ~6.5 x the wave CFL
    o = _implicit(P, tol=1e-11)
    for k in range(3):
        assert o.newmark_step(k) == 0, o.error()
        assert o.newmark_iters() <= 6
        R = o.newmark_get("R")
        assert np.linalg.norm(R) <= 1e-7 * P.materials[0][1][1] * P.dx ** 2


GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tangent_blocks.npz")


@pytest.mark.parametrize("case", ["dp", "mn"])
def test_elastoplastic_tangent_block_matches_reference(case):
    g = np.load(GOLD)
    gin, gout = g[case + "_in"], g[case + "_out"]
    assert len(gin) >= 100
    worst = 0.0
    for row, ref in zip(gin, gout):
        K = oracle.stiffness_ep(2, row[0:2], row[2:4], row[4:8], row[8:12], row[12:16])
        worst = max(worst, np.abs(K - ref).max() / max(np.abs(ref).max(), 1e-300))
    assert worst <= 1e-12, worst


def test_neo_hookean_tangent_block_matches_reference():
    g = np.load(GOLD)
    for row, ref in zip(g["nh_in"], g["nh_out"]):
        K = oracle.stiffness_nh(2, row[0:2], row[2:4], row[4:6], row[6:8], row[8:12], row[12], row[13], row[14])
        assert np.array_equal(K, ref) or np.abs(K - ref).max() <= 1e-15 * np.abs(ref).max()


def test_tangent_blocks_against_the_compiled_reference_live():
    """same check against oracle/_ref itself where it exists (this container), on fresh random states"""
    import refharness
    if not refharness.available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(3)
    for _ in range(50):
        A = rng.standard_normal((2, 2))
        be = np.eye(2) + 0.1 * (A + A.T)
        B = rng.standard_normal((2, 2))
        tau = 1e4 * (B + B.T)
        cep = 1e6 * rng.standard_normal(4)
        u, v = rng.standard_normal(2), rng.standard_normal(2)
        ref = refharness.stiffness_ep(u, v, list(be.ravel()) + [1.0], list(tau.ravel()) + [0.0], cep)
        K = oracle.stiffness_ep(2, u, v, be.ravel(), tau.ravel(), cep)
        assert np.abs(K - ref).max() <= 1e-12 * np.abs(ref).max()


@pytest.mark.parametrize("case,mult,nsteps", [("dp", 4.0, 6), ("mn", 2.0, 3)])
def test_elastoplastic_newton_converges(case, mult, nsteps):
    """Drucker-Prager / Matsuoka-Nakai with the reference's spectral tangent: it is not the exact derivative of the
    residual (Newton converges linearly, as in the reference), but every step must reach the tolerance"""
    from util import load_problem
    P = load_problem(case)
    P.solver["cfl"] *= mult
    o = _implicit(P, tol=1e-10)
    for k in range(nsteps):
        assert o.newmark_step(k) == 0, o.error()
        assert o.newmark_iters() < 25
    if case == "dp":
        assert (o.field("EPS_n") > 0).sum() > 50      # the tangent was evaluated on plastic states
    st, K = o.newmark_tangent()
    assert st == 0 and np.isfinite(K).all()


def test_static_scheme_balances_gravity():
    """U_Static (U-Static.c): Newton on f_int - f_trac - M b = 0.  After a converged step the internal forces carry the
    weight (the vertical reactions on the fixed nodes sum to m g), velocities and accelerations stay untouched."""
    P = synthetic.block_2d(cells=6, nsteps=3)
    o = oracle.Oracle(P)
    assert o.init_lme() == 0
    o.static_setup(tol=1e-11, max_iter=25)
    v0 = o.field("vel").copy()
    for k in range(2):
        assert o.newmark_step(k) == 0, o.error()
        assert o.newmark_iters() <= 8
    assert np.array_equal(o.field("vel"), v0) and np.abs(o.field("acc")).max() == 0.0
    # residual at the converged increment: zero on the free dofs
    assert o.newmark_begin(2) == 0
    dU = o.newmark_get("dU")
    st, R = o.newmark_residual(2, dU)
    free = (o.active()[:, None] > 0) & (o.fixed() == 0)
    weight = 9.81 * float(P.fields["mass"].sum())
    assert st == 0 and np.abs(R[free]).max() <= 1e-6 * weight    # one more step is already (nearly) in equilibrium
    assert np.abs(o.field("dis")[:, 1]).max() > 1e-4               # the block did settle


# ---- the scheme itself, pinned: the reference's OWN compiled U-Newmark-beta.c / U-Static.c (run against oracle/minipetsc,
# tests/golden/make_golden.py::gen_newmark) froze converged states of 2D decks; the restatement reproduces them
NEWMARK_KEYS = ["nh", "nh_trial", "dp", "mn", "static_nh", "vm", "vm_plastic", "hencky", "nhload", "static_nhload", "mixed", "almenh", "almedp"]
NEWMARK_FIELDS = ("x_GC", "dis", "vel", "acc", "F_n", "Stress", "rho", "J_n", "W", "b_e_n", "EPS_n", "Kappa_n", "lambda")


@pytest.mark.parametrize("key", NEWMARK_KEYS)
def test_scheme_matches_the_reference_compiled_scheme(key):
    """orc_newmark_* against the reference's own U_Newmark_Beta / U_Static (every stage function compiled from the
    reference; Newton + step halving and dense LU from oracle/mini_petsc.c, the same algorithm as the restatement's):
    all particle fields after 1 .. 8 implicit steps at 2 .. 8 times the explicit time step, plastic flow included (dp),
    and the same number of Newton iterations and residual evaluations."""
    from util import assert_close, field_scales, newmark_golden, newmark_problem
    g = newmark_golden(key)
    for k in g["checkpoints"]:
        k = int(k)
        P = newmark_problem(g, k)
        o = oracle.Oracle(P)
        assert o.init_lme() == 0
        if str(g["scheme"]) == "Static":
            o.static_setup(tol=float(g["tol"]), max_iter=int(g["max_iter"]))
        else:
            o.newmark_setup(tol=float(g["tol"]), max_iter=int(g["max_iter"]), explicit_trial=bool(g["explicit_trial"]))
        iters = 0
        for s in range(k):
            assert o.newmark_step(s) == 0, o.error()
            iters += o.newmark_iters()
        stats = g[f"s{k}_stats"]
        assert stats[0] == k
        if key != "vm_plastic":      # (see make_golden.py: the reference's in-place back stress stalls its own line search)
            assert stats[3] == 0     # every solve of the reference run converged
        assert iters == int(stats[1]), (iters, stats)
        sc = field_scales(P)
        for name in NEWMARK_FIELDS + (("Back_stress",) if key.startswith("vm") else ()) + \
                (("Beta", "Cut_off_Ellipsoid") if key.startswith("alme") else ()):
            assert_close(o.field(name), g[f"s{k}_{name}"], f"newmark {key} step {k} {name}", rtol=1e-12, scale=sc.get(name))
        assert np.array_equal(o.ints("I0"), g[f"s{k}_I0"])
        assert np.array_equal(o.ints("NumberNodes"), g[f"s{k}_NumberNodes"])
    last = int(max(g["checkpoints"]))
    assert np.abs(g[f"s{last}_dis"]).max() > 1e-5
    if key in ("dp", "vm_plastic", "mixed", "almedp"):
        assert (g[f"s{last}_EPS_n"] > 0).sum() > 20   # plastic flow reached
