"""Shared helpers for the parity tests."""
from __future__ import annotations

import os

import numpy as np

from nlps_b200.problem import Problem

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# SURVEY 8(d) parity protocol / north_star: <= 1e-10 relative in fp64, with an absolute floor of
# 1e-14 x the field's scale (a field that is ~0 everywhere compares against that floor).
RTOL = 1e-10
FLOOR = 1e-14

TRACE_FIELDS = ("x_GC", "dis", "D_dis", "vel", "acc", "F_n", "DF", "Stress", "rho", "J_n", "W",
                "b_e_n", "EPS_n", "Kappa_n", "lambda", "Beta")
NODAL = ("M", "dU", "F", "A", "R")


def load_problem(case) -> Problem:
    return Problem.from_npz(np.load(os.path.join(GOLDEN, f"{case}_problem.npz")))


def load_trace(case):
    return np.load(os.path.join(GOLDEN, f"{case}_trace.npz"))


def load_points(case):
    return np.load(os.path.join(GOLDEN, f"{case}_points.npz"))


def rel_err(a, b, scale=None):
    a, b = np.asarray(a, float), np.asarray(b, float)
    fin = np.isfinite(b)
    a, b = a[fin], b[fin]
    s = max(float(np.abs(b).max()) if b.size else 0.0, scale or 0.0, 1e-300)
    return float(np.abs(a - b).max()) / s if a.size else 0.0


def assert_close(a, b, what, rtol=RTOL, scale=None):
    e = rel_err(a, b, scale)
    # the reference itself yields inf/nan in places (e.g. the apex tangent with psi = 0 divides by
    # alpha_Q = 0, Drucker-Prager.c:1225): the non-finite pattern must agree, finite entries compare.
    assert np.array_equal(np.isfinite(np.asarray(a, float)), np.isfinite(np.asarray(b, float))), \
        f"{what}: non-finite pattern differs"
    assert e <= rtol, f"{what}: rel err {e:.3e} > {rtol:.1e}"


def field_scales(prob):
    """Absolute floor of each field = 1e-14 x its NATURAL scale (SURVEY 8(d) parity protocol), passed
    to assert_close as scale = 1e-4 x natural because rtol = 1e-10.  Natural scales: stresses and
    the strain energy density are differences of O(E) terms (W = f(J) + G/2 (I1 - d) cancels to
    E*strain^2, so its round-off is ~E*eps no matter how small W is); kinematic fields scale with
    the cell size h and the time step."""
    E = max(m[1][1] for m in prob.materials)
    h = prob.dx
    dt = prob.dt()
    nat = dict(Stress=E, W=E, vel=h / dt, acc=h / dt ** 2, dis=h, D_dis=h, x_GC=1.0, C_ep=E,
               **{"lambda": 1.0 / h})
    d = prob.ndim
    rho = max(m[1][0] for m in prob.materials)
    nat.update(gM=rho * h ** d, gdU=h, gF=E * h ** (d - 1), gA=h / dt ** 2, gR=E * h ** (d - 1))
    return {k: 1e-4 * v for k, v in nat.items()}


# ---- BASELINE configs at their stated shape (fixtures tests/golden/{c1,c2twin}_config.npz, tests/golden/make_golden.py)
CONFIG_CASES = {
    "c1": dict(grid=(20, 20), h=0.0625, block=(16, 16), origin=(2, 0), mat="nh_c1", nsteps=200, cel=31.622776601683793,
               checkpoints=(100, 200), kick=0.0),
    "c2twin": dict(grid=(264, 110), h=0.2 / 44, block=(44, 88), origin=(0, 0), mat="dp_c2", nsteps=120,
                   cel=(1e7 / 2000.0) ** 0.5 * 1.3, checkpoints=(120,), kick=-0.12),
}
CONFIG_FIELDS_SMALL = ("x_GC", "vel", "Stress", "F_n", "J_n", "EPS_n", "Kappa_n", "lambda", "rho")


def config_problem(case):
    """The Problem of a CONFIG_CASES entry from the synthetic generators (what the GPU tests run)."""
    from nlps_b200 import synthetic
    c = CONFIG_CASES[case]
    mat = synthetic.NH_C1 if c["mat"] == "nh_c1" else synthetic.DP_C2
    P = synthetic.structured_problem(2, c["grid"], c["h"], c["block"], c["origin"], mat, c["nsteps"], 0.5, c["cel"],
                                     (0.0, -9.81))
    if c["kick"]:
        P.fields["vel"][:, 1] = c["kick"] * P.solver["cel"]
    return P




def load_config(case):
    """(Problem rebuilt by the generator with the fixture's bit-exact seeds laid over it, golden npz)."""
    g = np.load(os.path.join(GOLDEN, f"{case}_config.npz"))
    P = config_problem(case)
    for k in g.files:
        if k.startswith("init_"):
            P.fields[k[5:]] = g[k].copy()
    return P, g


def newmark_golden(key):
    return np.load(os.path.join(GOLDEN, f"newmark_{key}.npz"))


def newmark_problem(g, nsteps):
    """The deck of the golden run as a Problem: the 2D fixture exported from the reference's own parser with the run's
    CFL and step count (its load curves are constant, so the tables are cut to the run's length)."""
    P = load_problem(str(g["case"]))
    P.solver["cfl"], P.solver["nsteps"] = float(g["cfl"]), nsteps
    for b in list(P.bounds) + list(P.neumann):
        b["dir"], b["val"] = b["dir"][:, :nsteps], b["val"][:, :nsteps]
    P.gravity = P.gravity[:, :nsteps]
    assert abs(P.dt() - float(g["dt"])) <= 1e-15 * float(g["dt"])
    return P
