"""Generate the golden fixtures that pin the oracle (and through it the CUDA engine) to the
reference's OWN compiled 2D code.

Run here (the container that has /root/reference), after `make -C oracle ref ref3d ref-newmark`:

    python tests/golden/make_golden.py                 # everything
    python tests/golden/make_golden.py sim almedp      # one explicit case: <case>_problem.npz + <case>_trace.npz
    python tests/golden/make_golden.py newmark dp      # one implicit case: newmark_<key>.npz

Explicit cases (`sim`): nh, dp, mn, vm, hencky (one law each), nhload (Neumann traction + moving platen), mixed (two
materials in one cloud), almenh / almedp (GramsShapeFun Type=aLME).  Implicit cases (`newmark`, NEWMARK_CASES): the
reference's OWN U_Newmark_Beta / U_Static, compiled unmodified against oracle/minipetsc (oracle/_ref/
libnlps2d_newmark_ref.so), run over the same decks.

Each case is produced in a fresh subprocess (the reference keeps its state in process globals):
a deck is written with tests/deckgen.py, parsed by the reference's parser, initialised by the
reference's initialize__LME__, and stepped by oracle/ref_harness.c::refh_step (reference stage
functions driven by the restated NPC-FS loop, 1 OpenMP thread for determinism).  The point-wise
files drive one material point through Stress_integration__Constitutive__ along strain paths that
include the reference's own stand-alone test path (tests/Constitutive/
Drucker-Prager-Backward-Euler.c:377-472: E=1e4, nu=0.2, kappa0=40, phi=39, psi=6, H=0.1, m=1,
DF=diag(1,0.999,1)) and sheared paths that rotate the principal axes (so that the transposed
eigenvector indexing of the plastic branches, SURVEY F10-i, is exercised).
"""
from __future__ import annotations

import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", "..", "nl-partsol_b200"))

import deckgen  # noqa: E402

MATERIALS = {
    "nh": ("Neo-Hookean-Wriggers", dict(rho=1000.0, E=1.0e6, nu=0.3), None),
    "dp": ("Drucker-Prager", {"rho": 2000.0, "E": 1e6, "nu": 0.3, "m": 1.0, "Hardening-modulus": 1.0,
                              "Reference-plastic-strain": 1e-2, "kappa-0": 20.0, "Friction-angle": 30.0,
                              "Dilatancy-angle": 0.0}, (1e7 / 2000) ** 0.5),
    "mn": ("Matsuoka-Nakai", {"rho": 2000.0, "E": 1e7, "nu": 0.3, "alpha": 0.5, "a1": 20000.0,
                              "a2": 0.005, "a3": 35.0, "Friction-angle": 30.0, "Cohesion": 1e3,
                              "kappa-0": 8.0 / 3.0}, 1.2 * (1e7 / 2000) ** 0.5),
    # the next laws of the K2 slot.  Von-Mises with mixed isotropic (linear + Voce) / kinematic hardening
    # (Constitutive/Plasticity/Von-Mises.c), Hencky (Constitutive/Hyperelastic/Hencky.c)
    "vm": ("Von-Mises", {"rho": 2000.0, "E": 1e6, "nu": 0.3, "Yield-stress": 1500.0, "Hardening-Modulus": 2e4,
                         "theta": 0.6, "K-0": 100.0, "K-inf": 600.0, "delta": 40.0}, (1e7 / 2000) ** 0.5),
    "hencky": ("Hencky", dict(rho=1000.0, E=1.0e6, nu=0.3), None),
}
POINT_MATERIALS = {
    "dp": ("Drucker-Prager", {"rho": 2000.0, "E": 1e4, "nu": 0.2, "m": 1.0, "Hardening-modulus": 0.1,
                              "kappa-0": 40.0, "Friction-angle": 39.0, "Dilatancy-angle": 6.0}),
    "mn": ("Matsuoka-Nakai", {"rho": 2000.0, "E": 1e7, "nu": 0.3, "alpha": 0.5, "a1": 20000.0,
                              "a2": 0.005, "a3": 35.0, "Friction-angle": 30.0, "Cohesion": 1e3,
                              "kappa-0": 8.0 / 3.0}),
    # Lade-Duncan has no cohesion in the reference's reader (the key aborts): the law only makes sense from a compressed
    # state, so its paths start from an isotropically pre-compressed elastic left Cauchy-Green tensor
    "ld": ("Lade-Duncan", {"rho": 2000.0, "E": 1e7, "nu": 0.3, "alpha": 0.5, "a1": 20000.0, "a2": 0.005, "a3": 35.0,
                           "Friction-angle": 30.0, "Atmospheric-pressure": 100.0}),
}
CHECKPOINTS = {"nh": (1, 2, 5, 20, 60), "dp": (1, 2, 5, 20, 60, 120), "mn": (1, 2, 5, 20, 60),
               "vm": (1, 2, 5, 20, 60, 120), "hencky": (1, 5, 60), "nhload": (1, 5, 40)}
MATERIALS["nhload"] = MATERIALS["nh"]
# two materials in one cloud (MatIdx per particle): the lower half Drucker-Prager, the upper half Neo-Hookean
MATERIALS["mixed"] = MATERIALS["dp"]
CHECKPOINTS["mixed"] = (1, 5, 60)
# GramsShapeFun (Type=aLME) (Nodes/aLME.c): the anisotropic shape functions, elastic and in plastic flow
MATERIALS["almenh"], MATERIALS["almedp"] = MATERIALS["nh"], MATERIALS["dp"]
CHECKPOINTS["almenh"], CHECKPOINTS["almedp"] = (1, 5, 40), (1, 20, 120)
TRACE_FIELDS = ("x_GC", "dis", "D_dis", "vel", "acc", "F_n", "DF", "Stress", "rho", "J_n", "W", "b_e_n",
                "EPS_n", "Kappa_n", "lambda", "Beta", "C_ep")   # (aLME: Beta is the d x d metric; + Cut_off_Ellipsoid)


def spec_for(case):
    model, params, cel = MATERIALS[case]
    spec = deckgen.DeckSpec()
    spec.material = deckgen.Material(model, params)
    if cel:
        spec.cel = cel
    spec.nsteps = max(CHECKPOINTS[case])
    if case.startswith("alme"):
        spec.shape = "aLME"
    if case == "mixed":
        nh_model, nh_params, _ = MATERIALS["nh"]
        upper = [j * spec.pnx + i for j in range(spec.pny // 2, spec.pny) for i in range(spec.pnx)]
        spec.more_materials = [(deckgen.Material(nh_model, dict(nh_params, rho=2000.0)), upper)]
    if case == "nhload":
        # the loads the other decks do not have: a Neumann traction (K3: U-Verlet.c:805-902, __nodal_traction_forces of
        # U-Newmark-beta.c:1388-1500) on the right column of particle cells, and a platen -- Dirichlet set with non-zero
        # displacement increments (U-Newmark-beta.c:909-946) -- on the top boundary of the grid, which the block reaches
        spec.pny = 12
        spec.dirichlet = spec.dirichlet + [("Top", "top", {"V.x": None, "V.y": -2.0e-5}, "CONSTANT_CURVE")]
        right = [j * spec.pnx + (spec.pnx - 1) for j in range(spec.pny)]
        spec.neumann = [("RightFace", right, {"T.x": 4.0e3, "T.y": -1.0e3}, "CONSTANT_CURVE")]
    return spec


def gen_sim(case):
    import refexport
    import refharness
    tmp = tempfile.mkdtemp(prefix="nlps_golden_")
    h = refharness.RefHarness(deckgen.write_deck(spec_for(case), tmp), threads=1)
    P = refexport.problem_from_ref(h)
    np.savez_compressed(os.path.join(HERE, f"{case}_problem.npz"), **P.to_npz_dict())
    cap = int((P.r2p[1:] - P.r2p[:-1]).max())
    trace = {}
    for k in range(max(CHECKPOINTS[case])):
        assert h.step(k) == 0, (case, k)
        if k + 1 in CHECKPOINTS[case]:
            t = f"s{k + 1}_"
            for f in TRACE_FIELDS:
                trace[t + f] = h.field(f)
            if case == "vm":
                trace[t + "Back_stress"] = h.field("Back_stress")
            if case.startswith("alme"):
                trace[t + "Cut_off_Ellipsoid"] = h.field("Cut_off_Ellipsoid")
            for w, nm in enumerate(("M", "dU", "F", "A", "R")):
                trace[t + "g" + nm] = h.nodal(w)
            lp, li = h.table(4)
            trace[t + "lists"] = refexport.lists_dense(lp, li, cap)
            trace[t + "I0"] = h.ints("I0")
            trace[t + "NumberNodes"] = h.ints("NumberNodes")
            trace[t + "active"] = h.active()
            n2m = np.zeros(P.nn, np.int32)
            d2m = np.zeros(P.nn * P.ndim, np.int32)
            import ctypes
            ip = ctypes.POINTER(ctypes.c_int)
            na = h.lib.refh_masks(k, n2m.ctypes.data_as(ip), d2m.ctypes.data_as(ip))
            trace[t + "nodes2mask"] = n2m
            trace[t + "dofs2mask"] = d2m[:na * P.ndim]
    trace["checkpoints"] = np.array(CHECKPOINTS[case])
    np.savez_compressed(os.path.join(HERE, f"{case}_trace.npz"), **trace)
    print(case, "ok: np", P.np_, "nn", P.nn, "max EPS", float(h.field("EPS_n").max()))


def gen_points(case):
    import refexport
    import refharness
    model, params = POINT_MATERIALS[case]
    spec = deckgen.DeckSpec(nx=6, ny=6, pnx=2, pny=2, porigin=(0.125, 0.125))
    spec.material = deckgen.Material(model, params)
    tmp = tempfile.mkdtemp(prefix="nlps_golden_")
    h = refharness.RefHarness(deckgen.write_deck(spec, tmp), threads=1)
    P = refexport.problem_from_ref(h)
    rng = np.random.default_rng(20261018)
    rows_in, rows_out = [], []
    for path in range(24):
        be = np.array([1, 0, 0, 1, 1.0])
        if case == "ld":
            be = be * (1.0 - 0.002 * (1 + path % 5)) ** 2
        eps, kap = float(P.fields["EPS_n"][0]), float(P.fields["Kappa_n"][0])
        F = np.array([1, 0, 0, 1, 1.0])
        amp = 10.0 ** rng.uniform(-4, -2)
        drift = rng.standard_normal(4)
        for step in range(50):
            D = np.eye(2) + amp * (drift.reshape(2, 2) + 0.3 * rng.standard_normal((2, 2)))
            if path == 0:
                D = np.diag([1.0, 0.999])  # the reference's own stand-alone test path
            elif path % 4 == 0:
                D = np.diag([1.0, 1 - amp])
            if case == "ld" and path % 4 != 0:  # shear-dominated increments keep the state inside the compression octant
                D = np.eye(2) + 0.2 * (D - np.eye(2)) + np.diag([0.0, -0.5 * amp])
            DF = np.array([D[0, 0], D[0, 1], D[1, 0], D[1, 1], 1.0])
            Fm = D @ F[:4].reshape(2, 2)
            F = np.array([Fm[0, 0], Fm[0, 1], Fm[1, 0], Fm[1, 1], 1.0])
            J = float(np.linalg.det(Fm))
            a = h.stress_point(0, DF, F, J, be, eps, kap)
            if a["status"] != 0 or not np.all(np.isfinite(a["stress"])):
                break
            rows_in.append(np.concatenate([DF, F, [J], be, [eps, kap]]))
            rows_out.append(np.concatenate([a["stress"], a["b_e_n1"], [a["eps_n1"], a["kappa_n1"], a["W"]],
                                            a["C_ep"]]))
            be, eps, kap = a["b_e_n1"], a["eps_n1"], a["kappa_n1"]
    mt, mp = P.materials[0]
    np.savez_compressed(os.path.join(HERE, f"{case}_points.npz"), inputs=np.array(rows_in),
                        outputs=np.array(rows_out), mat_type=np.array(mt), mat_params=mp,
                        tol_radial=P.solver["tol_radial"], maxiter_radial=P.solver["maxiter_radial"])
    nplastic = int((np.array(rows_out)[:, 10] != np.array(rows_in)[:, 16]).sum())
    print(case, "points:", len(rows_in), "plastic:", nplastic)


def gen_tangent_blocks(_case="all"):
    """Golden vectors of the implicit scheme's tangent blocks (K5) from the reference's own compiled functions:
    compute_stiffness_elastoplastic__Constitutive__ (Elastoplastic-Tangent-Matrix.c:42-160) on the material-point
    states of {dp,mn}_points.npz (plastic and elastic, rotated principal axes) and
    compute_stiffness_density_Neo_Hookean (Neo-Hookean.c:89-141) on random deformed states."""
    import refharness
    rng = np.random.default_rng(20261019)
    out = {}
    for case in ("dp", "mn"):
        pts = np.load(os.path.join(HERE, f"{case}_points.npz"))
        o = pts["outputs"]
        sel = np.linspace(0, len(o) - 1, 160).astype(int)
        rows_in, rows_out = [], []
        for k in sel:
            stress, be, cep = o[k, 0:5], o[k, 5:10], o[k, 13:17]
            u, v = rng.standard_normal(2), rng.standard_normal(2)
            K = refharness.stiffness_ep(u, v, be, stress, cep)
            rows_in.append(np.concatenate([u, v, be[:4], stress[:4], cep]))
            rows_out.append(K)
        out[case + "_in"], out[case + "_out"] = np.array(rows_in), np.array(rows_out)
    rows_in, rows_out = [], []
    for k in range(120):
        F = np.eye(2) + 0.2 * rng.standard_normal((2, 2))
        J = float(np.linalg.det(np.eye(2) + 0.05 * rng.standard_normal((2, 2)) ) * np.linalg.det(F))
        E, nu = 10.0 ** rng.uniform(4, 8), rng.uniform(0.0, 0.45)
        u, v, un, vn = (rng.standard_normal(2) for _ in range(4))
        K = refharness.stiffness_nh(u, v, un, vn, F.ravel(), J, E, nu)
        rows_in.append(np.concatenate([u, v, un, vn, F.ravel(), [J, E, nu]]))
        rows_out.append(K)
    out["nh_in"], out["nh_out"] = np.array(rows_in), np.array(rows_out)
    np.savez_compressed(os.path.join(HERE, "tangent_blocks.npz"), **out)
    print("tangent blocks:", {k: v.shape for k, v in out.items()})


def gen_points3d(case):
    """3D material points through the reference's OWN compiled 3D laws (oracle/_ref/libnlps3d_laws_ref.so = Drucker-Prager.c /
    Matsuoka-Nakai.c / Elastoplastic-Tangent-Matrix.c built with NumberDimensions == 3 + oracle/ref_harness3d.c; `make -C oracle
    ref3d`): sheared strain paths that rotate the principal axes, so that the row-indexed eigenvectors of the plastic branches
     are exercised in 3D, plus tangent blocks on the resulting states."""
    import ctypes
    sys.path.insert(0, os.path.join(HERE, "..", "..", "nl-partsol_b200"))
    from nlps_b200 import synthetic
    L = ctypes.CDLL(os.path.join(HERE, "..", "..", "oracle", "_ref", "libnlps3d_laws_ref.so"))
    dp_ = ctypes.POINTER(ctypes.c_double)
    arr = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    law, mpar = {"dp": synthetic.DP_C2, "mn": synthetic.MN_C4}[case]
    mpar = arr(mpar)
    tol, mi = {"dp": (1e-14, 10), "mn": (1e-10, 20)}[case]
    rng = np.random.default_rng(20261020 + (case == "mn"))
    rows_in, rows_out, tang_in, tang_out = [], [], [], []
    for path in range(16):
        be, eps, kap, F = np.eye(3).ravel(), 0.0, float(mpar[4]), np.eye(3)
        amp = 10.0 ** rng.uniform(-3.6, -2.2)
        drift = rng.standard_normal((3, 3))
        for step in range(40):
            D = np.eye(3) + amp * (-np.abs(drift) * np.eye(3) * (1.0 if case == "mn" else 2.0) + 0.5 * drift +
                                   0.2 * rng.standard_normal((3, 3)))
            if path == 0:
                D = np.diag([1.0, 0.999, 1.0])      # the reference's stand-alone test path, in 3D
            F = D @ F
            st, be1, cep = np.zeros(9), np.zeros(9), np.zeros(9)
            e1, k1, W = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
            Dc, Fc, bec = arr(D.ravel()), arr(F.ravel()), arr(be)
            rc = L.refh3_stress_point(law.encode(), mpar.ctypes.data_as(dp_), ctypes.c_double(tol), mi, Dc.ctypes.data_as(dp_),
                                      Fc.ctypes.data_as(dp_), bec.ctypes.data_as(dp_), ctypes.c_double(eps),
                                      ctypes.c_double(kap), st.ctypes.data_as(dp_), be1.ctypes.data_as(dp_), ctypes.byref(e1),
                                      ctypes.byref(k1), ctypes.byref(W), cep.ctypes.data_as(dp_))
            if rc != 0 or not np.all(np.isfinite(st)) or not np.all(np.isfinite(cep)):
                break
            rows_in.append(np.concatenate([D.ravel(), F.ravel(), [float(np.linalg.det(F))], be, [eps, kap]]))
            rows_out.append(np.concatenate([st, be1, [e1.value, k1.value, W.value], cep]))
            if step % 5 == 4:
                u, v = rng.standard_normal(3), rng.standard_normal(3)
                K = np.zeros(9)
                b2, s2, c2 = be1.copy(), st.copy(), cep.copy()
                assert L.refh3_stiffness_ep(K.ctypes.data_as(dp_), arr(u).ctypes.data_as(dp_), arr(v).ctypes.data_as(dp_),
                                            b2.ctypes.data_as(dp_), s2.ctypes.data_as(dp_), c2.ctypes.data_as(dp_)) == 0
                tang_in.append(np.concatenate([u, v, be1, st, cep]))
                tang_out.append(K)
            be, eps, kap = be1, e1.value, k1.value
    rows_in, rows_out = np.array(rows_in), np.array(rows_out)
    np.savez_compressed(os.path.join(HERE, f"{case}_points3d.npz"), inputs=rows_in, outputs=rows_out, mat_type=np.array(law),
                        mat_params=mpar, tol_radial=tol, maxiter_radial=mi, tang_in=np.array(tang_in),
                        tang_out=np.array(tang_out))
    print(case, "3D points:", len(rows_in), "plastic:", int((rows_out[:, 18] != rows_in[:, 28]).sum()), "tangents:", len(tang_in))


def gen_lme3d(_case="all"):
    """3D LME evaluations through the reference's OWN compiled LME.c (oracle/_ref/libnlps3d_lme_ref.so = Nodes/LME.c +
    its Matlib helpers built with NumberDimensions == 3 behind oracle/ref_harness3d_lme.c): for particles of a jittered 3D
    cloud (neighbour sets of 20-110 nodes, gamma 3 and 6) the converged lambda of __lambda_Newton_Rapson from two start
    values (the cloud's own lambda and zero), N = p__LME__ and grad N = dp__LME__."""
    import ctypes
    sys.path.insert(0, os.path.join(HERE, "..", "..", "nl-partsol_b200"))
    sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
    import oracle
    from nlps_b200 import synthetic
    L = ctypes.CDLL(os.path.join(HERE, "..", "..", "oracle", "_ref", "libnlps3d_lme_ref.so"))
    dp_ = ctypes.POINTER(ctypes.c_double)
    rows = []
    for gamma in (6.0, 3.0):
        P = synthetic.structured_problem(3, (8, 8, 8), 0.125, (4, 4, 4), (2, 2, 0), synthetic.NH_C1, 4, 0.5, 40.0,
                                         (0.0, 0.0, -9.81), gamma_lme=gamma, jitter=0.2,
                                         rollers=("left", "right", "front", "back"))
        o = oracle.Oracle(P)
        assert o.init_lme() == 0
        for k in range(2):
            assert o.step(k) == 0
        x, lam, beta, lists, nn = o.field("x_GC"), o.field("lambda"), o.field("Beta"), o.lists(), o.ints("NumberNodes")
        for p in range(0, P.np_, 3):
            n = int(nn[p])
            l = np.ascontiguousarray(x[p][None, :] - P.coords[lists[p, :n]])
            for start in (lam[p].copy(), np.zeros(3)):
                lm = np.ascontiguousarray(start, dtype=np.float64).copy()
                N, dN = np.zeros(n), np.zeros((n, 3))
                st = L.refh3_lme_point(n, l.ctypes.data_as(dp_), lm.ctypes.data_as(dp_), ctypes.c_double(float(beta[p])),
                                       ctypes.c_double(P.solver["tol_wrapper"]), int(P.solver["max_iter_lme"]),
                                       N.ctypes.data_as(dp_), dN.ctypes.data_as(dp_))
                assert st == 0
                rows.append(dict(l=l, start=np.array(start, float), beta=float(beta[p]), lam=lm, N=N, dN=dN))
    cap = max(len(r["N"]) for r in rows)
    pad = lambda a, shape: np.pad(a, [(0, s - t) for s, t in zip(shape, a.shape)])
    np.savez_compressed(os.path.join(HERE, "lme_points3d.npz"),
                        n=np.array([len(r["N"]) for r in rows]), l=np.array([pad(r["l"], (cap, 3)) for r in rows]),
                        start=np.array([r["start"] for r in rows]), beta=np.array([r["beta"] for r in rows]),
                        lam=np.array([r["lam"] for r in rows]), N=np.array([pad(r["N"], (cap,)) for r in rows]),
                        dN=np.array([pad(r["dN"], (cap, 3)) for r in rows]), tol_wrapper=1e-10, max_iter=10)
    ns = [len(r["N"]) for r in rows]
    print("3D LME points:", len(rows), "neighbours", min(ns), "-", max(ns))


def gen_nh3d(_case="all"):
    """Neo-Hookean in 3D through the reference's compiled Neo-Hookean.c (oracle/_ref/libnlps3d_laws_ref.so): Kirchhoff stress
    and strain energy (Neo-Hookean.c:38-85) on random deformation gradients, tangent blocks (:89-141) on random states."""
    import ctypes
    L = ctypes.CDLL(os.path.join(HERE, "..", "..", "oracle", "_ref", "libnlps3d_laws_ref.so"))
    dp_ = ctypes.POINTER(ctypes.c_double)
    arr = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    rng = np.random.default_rng(20261021)
    s_in, s_out, t_in, t_out = [], [], [], []
    for k in range(200):
        F = np.eye(3) + 10.0 ** rng.uniform(-3, -0.7) * rng.standard_normal((3, 3))
        if np.linalg.det(F) <= 0.1:
            continue
        J, E, nu = float(np.linalg.det(F)), 10.0 ** rng.uniform(4, 8), rng.uniform(0.0, 0.45)
        T, W, Fc = np.zeros(9), ctypes.c_double(), arr(F.ravel())
        assert L.refh3_stress_nh(ctypes.c_double(E), ctypes.c_double(nu), Fc.ctypes.data_as(dp_), ctypes.c_double(J),
                                 T.ctypes.data_as(dp_), ctypes.byref(W)) == 0
        s_in.append(np.concatenate([F.ravel(), [J, E, nu]]))
        s_out.append(np.concatenate([T, [W.value]]))
        u, v, un, vn = (arr(rng.standard_normal(3)) for _ in range(4))
        K, Fn = np.zeros(9), arr(F.ravel())
        Jt = float(J * (1 + 0.05 * rng.standard_normal()))
        assert L.refh3_stiffness_nh(K.ctypes.data_as(dp_), u.ctypes.data_as(dp_), v.ctypes.data_as(dp_), un.ctypes.data_as(dp_),
                                    vn.ctypes.data_as(dp_), Fn.ctypes.data_as(dp_), ctypes.c_double(Jt), ctypes.c_double(E),
                                    ctypes.c_double(nu)) == 0
        t_in.append(np.concatenate([u, v, un, vn, F.ravel(), [Jt, E, nu]]))
        t_out.append(K)
    np.savez_compressed(os.path.join(HERE, "nh_points3d.npz"), s_in=np.array(s_in), s_out=np.array(s_out), t_in=np.array(t_in),
                        t_out=np.array(t_out))
    print("3D NH points:", len(s_in), "tangent blocks:", len(t_in))


LISTS3D = dict(grid=(8, 8, 8), h=0.125, block=(4, 4, 4), origin=(2, 2, 0), jitter=0.2, steps=3)


def lists3d_state(gamma):
    """the 3D oracle state whose neighbour lists are frozen in lists3d.npz (also replayed by the test)"""
    sys.path.insert(0, os.path.join(HERE, "..", "..", "nl-partsol_b200"))
    sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
    import oracle
    from nlps_b200 import synthetic
    c = LISTS3D
    P = synthetic.structured_problem(3, c["grid"], c["h"], c["block"], c["origin"], synthetic.NH_C1, 6, 0.5, 40.0,
                                     (0.0, 0.0, -9.81), gamma_lme=gamma, jitter=c["jitter"],
                                     rollers=("left", "right", "front", "back"))
    P.fields["vel"][:, 2] = -0.2 * P.solver["cel"]          # particles change cells, lists change
    o = oracle.Oracle(P)
    assert o.init_lme() == 0
    for k in range(c["steps"]):
        assert o.step(k) == 0
    x, beta_old, I0_old = o.field("x_GC").copy(), o.field("Beta").copy(), o.ints("I0").copy()
    assert o.search_closest() == 0
    I0, active = o.ints("I0").copy(), o.active().copy()
    assert o.search_lists() == 0
    o.I0_before_search = I0_old
    return P, o, x, beta_old, I0, active


def gen_lists3d(_case="all"):
    """3D neighbour lists through the reference's OWN tributary__LME__ (LME.c:1019-1099, compiled in 3D inside
    oracle/_ref/libnlps3d_lme_ref.so): for every particle of a moving jittered cloud, the 2-ring of its closest node in
    chain order + ActiveNode flags + the previous beta go in, the ordered list comes out."""
    import ctypes
    L = ctypes.CDLL(os.path.join(HERE, "..", "..", "oracle", "_ref", "libnlps3d_lme_ref.so"))
    dp_, ip_ = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
    out = {}
    for gamma in (6.0, 3.0):
        P, o, x, beta_old, I0, active = lists3d_state(gamma)
        I0_ref = np.zeros(P.np_, np.int32)      # get_closest_node__MeshTools__ over the 1-ring of the previous closest node
        for p in range(P.np_):
            cand = P.r1i[P.r1p[o.I0_before_search[p]]:P.r1p[o.I0_before_search[p] + 1]]
            cc, xp = np.ascontiguousarray(P.coords[cand]), np.ascontiguousarray(x[p])
            I0_ref[p] = cand[L.refh3_closest(len(cand), cc.ctypes.data_as(dp_), xp.ctypes.data_as(dp_))]
        out[f"g{int(gamma)}_I0"] = I0_ref
        assert (I0_ref != o.I0_before_search).sum() > 20          # the cloud did move
        cap = int((P.r2p[1:] - P.r2p[:-1]).max())
        lists = np.full((P.np_, cap), -1, np.int32)
        counts = np.zeros(P.np_, np.int32)
        for p in range(P.np_):
            cand = P.r2i[P.r2p[I0[p]]:P.r2p[I0[p] + 1]]
            cc = np.ascontiguousarray(P.coords[cand])
            act = np.ascontiguousarray(active[cand].astype(np.uint8))
            res = np.zeros(len(cand), np.int32)
            xp = np.ascontiguousarray(x[p])
            n = L.refh3_tributary(len(cand), cc.ctypes.data_as(dp_), act.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)),
                                  xp.ctypes.data_as(dp_), ctypes.c_double(float(beta_old[p])),
                                  ctypes.c_double(P.solver["tol_zero"]), res.ctypes.data_as(ip_))
            counts[p] = n
            lists[p, :n] = cand[res[:n]]
        out[f"g{int(gamma)}_lists"], out[f"g{int(gamma)}_counts"] = lists, counts
        print("gamma", gamma, "particles", P.np_, "neighbours", counts.min(), "-", counts.max())
    np.savez_compressed(os.path.join(HERE, "lists3d.npz"), **out)


def kin3d_state():
    """3D oracle state stopped between the nodal increments and the kinematics of a step (replayed by the test)"""
    P, o, x, beta_old, I0, active = lists3d_state(6.0)
    k = LISTS3D["steps"]
    assert o.stage("p2g_mass_disp", k) == 0 and o.stage("grid_disp", k) == 0
    return P, o, k


def gen_kin3d(_case="all"):
    """3D kinematics through the reference's compiled Particles/compute-Strains.c (:20-44, :76-105 inside
    oracle/_ref/libnlps3d_laws_ref.so): DF = I + sum_A dU_A (x) grad N_A and F_n1 = DF F_n for every particle of the moving
    jittered cloud, from the oracle's nodal increments, neighbour lists and gradients."""
    import ctypes
    L = ctypes.CDLL(os.path.join(HERE, "..", "..", "oracle", "_ref", "libnlps3d_laws_ref.so"))
    dp_ = ctypes.POINTER(ctypes.c_double)
    P, o, k = kin3d_state()
    dU, lists, nn, Fn = o.nodal(1), o.lists(), o.ints("NumberNodes"), o.field("F_n")
    DF, F1 = np.zeros((P.np_, 9)), np.zeros((P.np_, 9))
    for p in range(P.np_):
        n = int(nn[p])
        N, dN = o.shape(p)
        du, gr, fn = (np.ascontiguousarray(a, dtype=np.float64) for a in (dU[lists[p, :n]], dN, Fn[p]))
        L.refh3_kinematics(n, du.ctypes.data_as(dp_), gr.ctypes.data_as(dp_), fn.ctypes.data_as(dp_),
                           DF[p].ctypes.data_as(dp_), F1[p].ctypes.data_as(dp_))
    np.savez_compressed(os.path.join(HERE, "kin3d.npz"), DF=DF, F_n1=F1)
    print("3D kinematics:", P.np_, "particles, max |DF - I|", float(np.abs(DF - np.eye(3).ravel()).max()))


# ---- BASELINE configs at their stated shape: C1 = 1024 particles x 200 steps, C2 twin = 1/8 linear scale
# (15,488 particles) x 120 steps with plastic flow.  The problems themselves are NOT stored: the generators of
# nl-partsol_b200/nlps_b200/synthetic.py rebuild them on the GPU box, and this script asserts that what they build is
# bit-identical to what the reference's parser, mesh reader and particle seeding produced from the deck.
from util import CONFIG_CASES, CONFIG_FIELDS_SMALL, config_problem  # noqa: E402


def gen_config(case):
    import hashlib
    import refexport
    import refharness
    c = CONFIG_CASES[case]
    spec = deckgen.DeckSpec(nx=c["grid"][0], ny=c["grid"][1], h=c["h"], pnx=c["block"][0], pny=c["block"][1], ph=c["h"],
                            porigin=(c["origin"][0] * c["h"], c["origin"][1] * c["h"]), nsteps=c["nsteps"], cfl=0.5, cel=c["cel"])
    if c["mat"] == "dp_c2":
        spec.material = deckgen.Material("Drucker-Prager", {
            "rho": 2000.0, "E": 1e7, "nu": 0.3, "m": 1.0, "Hardening-modulus": 1.0,
            "Reference-plastic-strain": 1e-2, "kappa-0": 1e4, "Friction-angle": 30.0, "Dilatancy-angle": 0.0})
    tmp = tempfile.mkdtemp(prefix="nlps_golden_")
    h = refharness.RefHarness(deckgen.write_deck(spec, tmp), threads=1)
    if c["kick"]:
        v = h.field("vel")
        v[:, 1] = c["kick"] * c["cel"]
        h.set_field("vel", v)
    Pref = refexport.problem_from_ref(h)
    Psyn = config_problem(case)
    for nm in ("coords", "r1p", "r1i", "r2p", "r2i", "h_avg", "I0", "MatIdx"):
        assert np.array_equal(getattr(Pref, nm), getattr(Psyn, nm)), (case, nm)
    skip = ("tol_radial", "maxiter_radial") if c["mat"] == "nh_c1" else ()   # globals no material of the deck sets (F10-iv)
    assert Pref.dx == Psyn.dx and all(Pref.solver[k] == Psyn.solver[k] for k in Pref.solver if k not in skip), (Pref.solver, Psyn.solver)
    # particle fields: the reference interpolates the seeds with the element shape functions and integrates the element
    # volume, the generator uses closed forms: equal to 1 ulp.  The few fields that differ in the last bit (and the
    # NaN the reference leaves in Kappa_n of a Neo-Hookean deck) are stored with the fixture and laid over the generated
    # problem by the test (lambda and Beta are the product of initialize__LME__, which the test runs itself)
    init = {}
    for k, v in Pref.fields.items():
        if k in ("lambda", "Beta") or np.array_equal(v, Psyn.fields[k], equal_nan=True):
            continue
        assert v.shape == Psyn.fields[k].shape and (np.isnan(v).any() or np.allclose(v, Psyn.fields[k], rtol=1e-13, atol=1e-300)
                                                     or k == "b_e_n1"), (case, "field", k)
        init[k] = v
    for a, b in zip(Pref.bounds, Psyn.bounds):
        # (the reference holds the node set in chain order = file order reversed; the order of a Dirichlet set has no effect)
        assert np.array_equal(np.sort(a["nodes"]), np.sort(b["nodes"])) and all(np.array_equal(a[k], b[k]) for k in ("dir", "val")), (case, "bounds")
    assert np.array_equal(Pref.gravity, Psyn.gravity)
    cap = int((Pref.r2p[1:] - Pref.r2p[:-1]).max())
    small = Pref.np_ > 4096
    fields = CONFIG_FIELDS_SMALL if small else TRACE_FIELDS
    out = {"checkpoints": np.array(c["checkpoints"]), "np": Pref.np_, "nn": Pref.nn}
    for k, v in init.items():
        out["init_" + k] = v
    for k in range(c["nsteps"]):
        assert h.step(k) == 0, (case, k)
        if k + 1 in c["checkpoints"]:
            t = f"s{k + 1}_"
            for f in fields:
                out[t + f] = h.field(f)
            out[t + "I0"] = h.ints("I0")
            out[t + "NumberNodes"] = h.ints("NumberNodes")
            lp, li = h.table(4)
            lists = refexport.lists_dense(lp, li, cap)
            if small:  # the ordered lists as a digest (15 k x 25 ints would be the largest array of the file)
                out[t + "lists_sha256"] = np.array(hashlib.sha256(np.ascontiguousarray(lists).tobytes()).hexdigest())
            else:
                out[t + "lists"] = lists
            out[t + "active"] = h.active()
    np.savez_compressed(os.path.join(HERE, f"{case}_config.npz"), **out)
    print(case, "ok: np", Pref.np_, "nn", Pref.nn, "max EPS", float(h.field("EPS_n").max()),
          "plastic particles", int((h.field("EPS_n") > 0).sum()))


# ---- the reference's OWN implicit schemes (U-Newmark-beta.c, U-Static.c) run against oracle/minipetsc
NEWMARK_SO = os.path.join(HERE, "..", "..", "oracle", "_ref", "libnlps2d_newmark_ref.so")
NEWMARK_CASES = {
    # key: (deck material, scheme, CFL, TOL-Newmark-beta, Max-Iter, Explicit-trial, checkpoints (steps run))
    "nh": ("nh", "Newmark-beta-Finite-Strains", 4.0, 1e-12, 25, 0, (1, 3, 6)),
    "nh_trial": ("nh", "Newmark-beta-Finite-Strains", 4.0, 1e-12, 25, 1, (6,)),
    "dp": ("dp", "Newmark-beta-Finite-Strains", 2.0, 1e-11, 25, 0, (1, 4, 8)),
    "mn": ("mn", "Newmark-beta-Finite-Strains", 1.0, 1e-11, 25, 0, (1, 4)),
    "static_nh": ("nh", "Static", 0.5, 1e-11, 30, 0, (1, 3)),
    # Von-Mises / Hencky tangents (Constitutive.c:284-297,316-329).  "vm" stays elastic and converges; "vm_plastic" yields:
    # the reference updates the back stress IN PLACE at every residual evaluation (Constitutive.c:110-143 restarts b_e and
    # EPS from step n, Phi.Back_stress has no n / n1 pair), so its residual is not a function of dU any more, the line
    # search stalls on some steps (stats[3] > 0) and the result depends on the sequence of evaluations -- reproduced by
    # the oracle (same sequence), not asked of the device
    "vm": ("vm", "Newmark-beta-Finite-Strains", 1.0, 1e-11, 50, 0, (1, 5)),
    "vm_plastic": ("vm", "Newmark-beta-Finite-Strains", 2.0, 1e-12, 25, 0, (8,)),
    "hencky": ("hencky", "Newmark-beta-Finite-Strains", 1.0, 1e-10, 100, 0, (1, 3)),
    # Neumann traction + moving platen (spec_for("nhload")), dynamic and quasi-static
    "nhload": ("nhload", "Newmark-beta-Finite-Strains", 4.0, 1e-12, 25, 0, (1, 6)),
    "static_nhload": ("nhload", "Static", 0.5, 1e-11, 30, 0, (1, 3)),
    "mixed": ("mixed", "Newmark-beta-Finite-Strains", 2.0, 1e-11, 25, 0, (1, 6)),
    "almenh": ("almenh", "Newmark-beta-Finite-Strains", 4.0, 1e-12, 25, 0, (1, 6)),
    "almedp": ("almedp", "Newmark-beta-Finite-Strains", 2.0, 1e-11, 25, 0, (1, 8)),
}
NEWMARK_FIELDS = ("x_GC", "dis", "vel", "acc", "F_n", "Stress", "rho", "J_n", "W", "b_e_n", "EPS_n", "Kappa_n", "lambda",
                  "Beta", "Back_stress")


def newmark_spec(key, nsteps):
    case, scheme, cfl, tol, max_iter, trial, _ = NEWMARK_CASES[key]
    spec = spec_for(case)
    spec.scheme = scheme
    spec.nsteps, spec.cfl, spec.out_every = nsteps, cfl, 100000
    spec.solver_extra = {"Beta-Newmark-beta": 0.25, "Gamma-Newmark-beta": 0.5, "TOL-Newmark-beta": tol,
                         "Max-Iter": max_iter, "Epsilon": 0.0}
    if trial:
        spec.solver_extra["Explicit-trial"] = 1
    return spec


def gen_newmark_run(key_steps):
    """One run of the reference's compiled U_Newmark_Beta / U_Static over `steps` time steps (fresh process: the reference
    keeps its simulation in process globals); prints nothing, writes a temporary npz that gen_newmark collects."""
    import ctypes
    import refexport
    import refharness
    key, steps, out = key_steps.split(":")
    steps = int(steps)
    tmp = tempfile.mkdtemp(prefix="nlps_golden_")
    h = refharness.RefHarness(deckgen.write_deck(newmark_spec(key, steps), tmp), so=NEWMARK_SO, threads=1)
    P = refexport.problem_from_ref(h)
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        rc = h.lib.refh_static_run() if NEWMARK_CASES[key][1] == "Static" else h.lib.refh_newmark_run()
    finally:
        os.chdir(cwd)
    assert rc == 0, (key, steps, rc)
    st = np.zeros(5)
    h.lib.refh_newmark_stats(st.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    res = {f: h.field(f) for f in NEWMARK_FIELDS}
    if key.startswith("alme"):
        res["Cut_off_Ellipsoid"] = h.field("Cut_off_Ellipsoid")
    res["I0"] = h.ints("I0")
    res["NumberNodes"] = h.ints("NumberNodes")
    res["stats"] = st
    res["x0"] = P.fields["x_GC"]
    res["dt"] = np.array(P.dt())
    np.savez(out, **res)


def gen_newmark(key):
    """Golden states of the implicit schemes from the reference's own compiled scheme code: every stage function is the
    reference's; the Newton loop / linear solve are oracle/mini_petsc.c (Newton + step halving, dense LU).  `stats` =
    SNES solves, Newton iterations, residual evaluations, solves stopped by Max-Iter, last |F|."""
    case, scheme, cfl, tol, max_iter, trial, cps = NEWMARK_CASES[key]
    out = dict(case=np.array(case), scheme=np.array(scheme), cfl=np.array(cfl), tol=np.array(tol),
               max_iter=np.array(max_iter), explicit_trial=np.array(trial), checkpoints=np.array(cps))
    ref_problem = np.load(os.path.join(HERE, f"{case}_problem.npz"))
    for k in cps:
        tmpf = os.path.join(tempfile.mkdtemp(prefix="nlps_golden_"), "run.npz")
        subprocess.run([sys.executable, __file__, "newmark_run", f"{key}:{k}:{tmpf}"], check=True,
                       stdout=subprocess.DEVNULL)
        r = np.load(tmpf)
        # the fixture {case}_problem.npz (written by gen_sim from the same deck generator) is the same initial state
        assert np.array_equal(r["x0"], ref_problem["f_x_GC"])
        for f in r.files:
            if f not in ("x0", "dt"):
                out[f"s{k}_{f}"] = r[f]
        out["dt"] = r["dt"]
    np.savez_compressed(os.path.join(HERE, f"newmark_{key}.npz"), **out)
    last = max(cps)
    print("newmark", key, "ok: stats", out[f"s{last}_stats"], "max |dis|", float(np.abs(out[f"s{last}_dis"]).max()),
          "max EPS", float(out[f"s{last}_EPS_n"].max()))


if __name__ == "__main__":
    if len(sys.argv) == 3:
        {"sim": gen_sim, "points": gen_points, "tangent": gen_tangent_blocks, "points3d": gen_points3d, "lme3d": gen_lme3d, "nh3d": gen_nh3d, "lists3d": gen_lists3d, "kin3d": gen_kin3d, "config": gen_config, "newmark": gen_newmark, "newmark_run": gen_newmark_run}[sys.argv[1]](sys.argv[2])
    else:
        for c in ("nh", "dp", "mn", "vm", "hencky", "nhload", "mixed", "almenh", "almedp"):
            subprocess.run([sys.executable, __file__, "sim", c], check=True, stdout=subprocess.DEVNULL
                           if os.environ.get("QUIET") else None)
        for c in ("dp", "mn", "ld"):
            subprocess.run([sys.executable, __file__, "points", c], check=True)
        subprocess.run([sys.executable, __file__, "tangent", "all"], check=True)
        for c in ("dp", "mn"):
            subprocess.run([sys.executable, __file__, "points3d", c], check=True)
        subprocess.run([sys.executable, __file__, "lme3d", "all"], check=True)
        subprocess.run([sys.executable, __file__, "nh3d", "all"], check=True)
        subprocess.run([sys.executable, __file__, "lists3d", "all"], check=True)
        subprocess.run([sys.executable, __file__, "kin3d", "all"], check=True)
        for c in ("c1", "c2twin"):
            subprocess.run([sys.executable, __file__, "config", c], check=True)
        for c in NEWMARK_CASES:     # needs `make -C oracle ref-newmark`
            subprocess.run([sys.executable, __file__, "newmark", c], check=True)
