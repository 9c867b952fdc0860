"""CPU checks of the implicit Newmark-beta restatement in oracle/ (parity unpinned: PETSc is absent, the
reference scheme cannot run here): the tangent is the derivative of the residual, the trapezoidal
scheme converges to the explicit oracle as dt -> 0, Newton converges quadratically."""
import numpy as np

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
