"""The 3D branches of the oracle's elastoplastic laws against the reference's OWN compiled 3D code.

The reference's 3D build does not compile as a whole, but Drucker-Prager.c, Matsuoka-Nakai.c and
Elastoplastic-Tangent-Matrix.c do (they need LAPACKE only): oracle/Makefile `ref3d` builds them with NumberDimensions == 3
behind oracle/ref_harness3d.c, tests/golden/make_golden.py `points3d` froze 600+ material-point updates per law on sheared
paths (half of them plastic) and 120+ tangent blocks.  Finding pinned here: the compiled 3D laws index the eigenvector
matrix by ROW in their plastic branches exactly as the 2D ones do -- the oracle reproduces them bit for
bit with quirk_transposed = 1 and differs by O(1) on rotated plastic states with the mathematically intended column form
(quirk 0, the default of the 3D engine)."""
import os

import numpy as np
import pytest

import oracle
from nlps_b200 import synthetic

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LAWS = {"dp": synthetic.DP_C2, "mn": synthetic.MN_C4}


def _oracle(case, quirk):
    P = synthetic.cube_3d(cells=2, nsteps=2, material=LAWS[case])
    g = np.load(os.path.join(GOLD, f"{case}_points3d.npz"))
    P.solver["tol_radial"], P.solver["maxiter_radial"] = float(g["tol_radial"]), int(g["maxiter_radial"])
    o = oracle.Oracle(P)
    o.set_flags(quirk, 1)
    return o, g


def _run(o, g):
    out = []
    for r in g["inputs"]:
        a = o.stress_point(0, r[0:9], r[9:18], r[18], r[19:28], r[28], r[29])
        assert a["status"] == 0
        out.append(np.concatenate([a["stress"], a["b_e_n1"], [a["eps_n1"], a["kappa_n1"], a["W"]], a["C_ep"]]))
    return np.array(out)


@pytest.mark.parametrize("case", ["dp", "mn"])
def test_3d_laws_match_the_compiled_reference(case):
    o, g = _oracle(case, quirk=1)
    got, ref = _run(o, g), g["outputs"]
    plastic = (ref[:, 18] != g["inputs"][:, 28]) | (ref[:, 19] != g["inputs"][:, 29])      # EPS or kappa moved
    assert plastic.sum() > 200 and (~plastic).sum() > 100
    s = np.abs(ref[:, :9]).max(axis=1, keepdims=True) + 1.0
    assert np.abs(got[:, :9] - ref[:, :9]).max() / 1.0 <= 1e-9 * s.max()           # Kirchhoff stress
    assert (np.abs(got[:, :9] - ref[:, :9]) / s).max() <= (1e-12 if case == "dp" else 1e-8)
    assert np.abs(got[:, 9:18] - ref[:, 9:18]).max() <= (1e-12 if case == "dp" else 1e-8)      # b_e
    assert np.abs(got[:, 18] - ref[:, 18]).max() <= (1e-14 if case == "dp" else 1e-9)          # EPS
    # C_ep: Matsuoka-Nakai's 3D plastic branch never writes it in the reference (the oracle and the engine
    # do, for the implicit tangent): compared on the rows where the reference writes it
    rows = np.ones(len(ref), bool) if case == "dp" else ~plastic
    fin = np.isfinite(ref[rows, 21:])
    assert np.array_equal(np.isfinite(got[rows, 21:]), fin)
    c = np.abs(ref[rows, 21:][fin]).max()
    assert np.abs(got[rows, 21:][fin] - ref[rows, 21:][fin]).max() <= 1e-10 * c


@pytest.mark.parametrize("case", ["dp", "mn"])
def test_3d_row_indexed_eigenvectors_are_what_the_reference_computes(case):
    """quirk 0 (column form, the intended mathematics): identical on elastic steps, different on rotated plastic ones"""
    o, g = _oracle(case, quirk=0)
    got, ref = _run(o, g), g["outputs"]
    plastic = (ref[:, 18] != g["inputs"][:, 28]) | (ref[:, 19] != g["inputs"][:, 29])
    s = np.abs(ref[:, :9]).max(axis=1) + 1.0
    d = np.abs(got[:, :9] - ref[:, :9]).max(axis=1) / s
    assert d[~plastic].max() <= 1e-9
    assert d[plastic].max() > 1e-2           # F10-i is live in the compiled 3D laws


@pytest.mark.parametrize("case", ["dp", "mn"])
def test_3d_elastoplastic_tangent_block_matches_the_compiled_reference(case):
    g = np.load(os.path.join(GOLD, f"{case}_points3d.npz"))
    worst = 0.0
    for row, ref in zip(g["tang_in"], g["tang_out"]):
        K = oracle.stiffness_ep(3, row[0:3], row[3:6], row[6:15], row[15:24], row[24:33])
        if not np.isfinite(ref).all():
            continue
        worst = max(worst, np.abs(K - ref).max() / max(np.abs(ref).max(), 1e-300))
    assert worst <= 1e-9, worst


def test_3d_lme_matches_the_compiled_reference():
    """K0 in 3D: the Newton solve for lambda (LME.c:272-353, from the cloud's own lambda and from zero), N = p__LME__ and
    grad N = dp__LME__ (:700-891) of the oracle against the reference's own LME.c compiled with NumberDimensions == 3
    (oracle/ref_harness3d_lme.c, tests/golden/lme_points3d.npz: 684 evaluations, 24-111 neighbours, gamma 3 and 6)."""
    g = np.load(os.path.join(GOLD, "lme_points3d.npz"))
    P = synthetic.cube_3d(cells=2, nsteps=2)
    P.solver["tol_wrapper"], P.solver["max_iter_lme"] = float(g["tol_wrapper"]), int(g["max_iter"])
    o = oracle.Oracle(P)
    worst = dict(lam=0.0, N=0.0, dN=0.0)
    for k in range(len(g["n"])):
        n = int(g["n"][k])
        st, lam, N, dN = o.lme_point(g["l"][k, :n], g["start"][k], float(g["beta"][k]))
        assert st == 0
        assert abs(N.sum() - 1.0) < 1e-13
        worst["lam"] = max(worst["lam"], np.abs(lam - g["lam"][k]).max() / max(np.abs(g["lam"][k]).max(), 1.0))
        worst["N"] = max(worst["N"], np.abs(N - g["N"][k, :n]).max())
        worst["dN"] = max(worst["dN"], np.abs(dN - g["dN"][k, :n]).max() / np.abs(g["dN"][k, :n]).max())
    assert worst["lam"] <= 1e-12 and worst["N"] <= 1e-13 and worst["dN"] <= 1e-11, worst


def test_3d_neo_hookean_matches_the_compiled_reference():
    """BASELINE configs[2] / [4]: Kirchhoff stress + strain energy (Neo-Hookean.c:38-85) and the tangent block (:89-141) of
    the oracle in 3D against the reference's compiled Neo-Hookean.c (tests/golden/nh_points3d.npz)."""
    g = np.load(os.path.join(GOLD, "nh_points3d.npz"))
    for row, ref in zip(g["t_in"], g["t_out"]):
        K = oracle.stiffness_nh(3, row[0:3], row[3:6], row[6:9], row[9:12], row[12:21], row[21], row[22], row[23])
        assert np.abs(K - ref).max() <= 1e-14 * max(np.abs(ref).max(), 1e-300)
    worst_s = worst_w = 0.0
    for row, ref in zip(g["s_in"], g["s_out"]):
        F, J, E, nu = row[0:9], float(row[9]), float(row[10]), float(row[11])
        P = synthetic.cube_3d(cells=2, nsteps=2, material=("Neo-Hookean-Wriggers", [1000.0, E, nu] + [0.0] * 13))
        o = oracle.Oracle(P)
        a = o.stress_point(0, np.eye(3).ravel(), F, J, np.eye(3).ravel(), 0.0, 0.0)
        assert a["status"] == 0
        worst_s = max(worst_s, np.abs(a["stress"] - ref[:9]).max() / max(np.abs(ref[:9]).max(), 1e-300))
        worst_w = max(worst_w, abs(a["W"] - ref[9]) / max(abs(ref[9]), 1e-300))
    assert worst_s <= 1e-13 and worst_w <= 1e-10, (worst_s, worst_w)


@pytest.mark.parametrize("gamma", [6.0, 3.0])
def test_3d_neighbour_lists_match_the_compiled_reference(gamma):
    """K0 in 3D, the search: the ordered neighbour lists of a moving jittered cloud (closest node after the move, 2-ring
    in chain order, ActiveNode flags, radius from the previous beta) against the reference's own tributary__LME__
    (LME.c:1019-1099) compiled in 3D -- bit-exact, order included (tests/golden/lists3d.npz)."""
    import sys
    sys.path.insert(0, GOLD)
    import make_golden
    P, o, x, beta_old, I0, active = make_golden.lists3d_state(gamma)
    g = np.load(os.path.join(GOLD, "lists3d.npz"))
    ref_l, ref_c = g[f"g{int(gamma)}_lists"], g[f"g{int(gamma)}_counts"]
    assert np.array_equal(I0, g[f"g{int(gamma)}_I0"])          # get_closest_node__MeshTools__ (Nodes-Tools.c:476-538)
    assert np.array_equal(o.ints("NumberNodes"), ref_c)
    got = o.lists()
    assert np.array_equal(got[:, :ref_l.shape[1]], ref_l)
    assert ref_c.min() >= 4 and (ref_c != ref_c[0]).any()


def test_3d_kinematics_match_the_compiled_reference():
    """K2a in 3D: DF = I + sum_A dU_A (x) grad N_A and F_n1 = DF F_n (compute-Strains.c:20-44, 76-105, compiled in 3D) for
    every particle of the moving jittered cloud -- the oracle's kinematics stage against tests/golden/kin3d.npz."""
    import sys
    sys.path.insert(0, GOLD)
    import make_golden
    P, o, k = make_golden.kin3d_state()
    assert o.stage("kin_stress", k) == 0
    g = np.load(os.path.join(GOLD, "kin3d.npz"))
    assert np.abs(g["DF"] - np.eye(3).ravel()).max() > 1e-4
    assert np.abs(o.field("DF") - g["DF"]).max() <= 1e-15
    assert np.abs(o.field("F_n1") - g["F_n1"]).max() <= 1e-15
