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


@pytest.mark.parametrize("case", CASES + ("almenh",))
def test_initialize_lme_matches_reference(case):
    """almenh: initialize__aLME__ (Nodes/aLME.c:32-166) -- the isotropic metric and cut-off ellipsoid, lists, lambda."""
    P = load_problem(case)
    P0 = P.copy()
    P0.fields["Beta"][:] = 0.0
    P0.fields["lambda"][:] = 0.0
    if case.startswith("alme"):
        P0.fields["Cut_off_Ellipsoid"][:] = 0.0
    eng = engine.Engine(P0)
    assert eng.initialize_lme() == 0, eng.error()
    f = eng.download()
    assert np.array_equal(f["Beta"], P.fields["Beta"])          # bit-exact (one division)
    if case.startswith("alme"):
        assert np.array_equal(f["Cut_off_Ellipsoid"], P.fields["Cut_off_Ellipsoid"])
    assert_close(f["lambda"], P.fields["lambda"], "lambda", scale=1e-3 / P.dx)
    o = oracle.Oracle(P0)
    assert o.init_lme() == 0
    counts, lists = eng.lists()
    assert np.array_equal(counts, o.ints("NumberNodes"))
    assert np.array_equal(lists, o.lists())
    assert np.array_equal(eng.active(), o.active())
    eng.close()


@pytest.mark.parametrize("case", CASES + ("vm", "hencky", "nhload", "mixed", "almenh", "almedp"))
def test_steps_match_golden_reference(case):
    """Multi-step run against fixtures produced by the reference's compiled code.  nhload: Neumann traction on a column of
    particles + a platen (Dirichlet set with non-zero displacement increments); mixed: two materials in one cloud
    (Drucker-Prager below, Neo-Hookean above: the kernels that read the law per particle); almenh / almedp:
    GramsShapeFun (Type=aLME), Nodes/aLME.c -- metric tensor and cut-off ellipsoid convected with DF^-1 at every search,
    elastic and with 7 % equivalent plastic strain (TRACE_FIELDS holds Beta: here the 2 x 2 metric)."""
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
        # (Von-Mises: also the tangent moduli of its return mapping, Von-Mises.c:730-757, which the implicit tangent reads)
        for name in TRACE_FIELDS + (("Back_stress", "C_ep") if case == "vm" else ()) + \
                (("Cut_off_Ellipsoid",) if case.startswith("alme") else ()):
            assert_close(f[name], tr[t + name], f"{case} step {cp} {name}", scale=scales.get(name))
        for w, nm in enumerate(NODAL):
            assert_close(eng.nodal(w), tr[t + "g" + nm], f"{case} step {cp} nodal {nm}", scale=scales["g" + nm])
    if case == "vm":
        assert (f["EPS_n"] > 0).sum() > 50 and np.abs(f["Back_stress"]).max() > 0
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
            assert_close(eng.nodal(w), o.nodal(w), f"step {k} nodal {NODAL[w]}", scale=scales["g" + NODAL[w]])
        # kinematics + stress
        assert o.stage("kin_stress", k) == 0 and eng.stage("kin_stress", k) == 0, eng.error()
        f = eng.download()
        for name in ("DF", "F_n1", "J_n1", "rho", "Stress", "W", "b_e_n1", "EPS_n1", "Kappa_n1", "C_ep"):
            assert_close(f[name], o.field(name), f"step {k} {name}", scale=scales.get(name))
        # forces + nodal equilibrium
        for st in ("force", "grid_acc"):
            assert o.stage(st, k) == 0 and eng.stage(st, k) == 0
        for w in (2, 3, 4):
            assert_close(eng.nodal(w), o.nodal(w), f"step {k} nodal {NODAL[w]}", scale=scales["g" + NODAL[w]])
        # G2P + corrector
        assert o.stage("g2p", k) == 0 and eng.stage("g2p", k) == 0
        f = eng.download()
        for name in TRACE_FIELDS:
            assert_close(f[name], o.field(name), f"step {k} {name} after g2p", scale=scales.get(name))
    eng.close()


GROUPS = ((slice(0, 5), "stress"), (slice(5, 10), "b_e"), (slice(10, 11), "eps"), (slice(11, 12), "kappa"),
          (slice(12, 13), "W"), (slice(13, 17), "C_ep"))


def _pack(r):
    return np.concatenate([r["stress"], r["b_e_n1"], [r["eps_n1"], r["kappa_n1"], r["W"]], r["C_ep"]])


def _group_scale(y, sl, nm, E):
    """max |reference value| of the group, floored at 1e-4 x the natural scale of the quantity (so the
    absolute floor is 1e-14 x natural scale): E for stresses / energy / moduli, 1 for strains."""
    nat = {"stress": E, "W": E, "C_ep": E, "b_e": 1.0, "eps": 1.0, "kappa": 1.0}[nm]
    return max(np.abs(y[sl]).max(), 1e-4 * nat)


def _reference_sensitivity(case, z):
    """Per output group: how much the (oracle restatement of the) reference's own result moves when the
    inputs are perturbed by 1e-15 relative.  Drucker-Prager iterates to 1e-14 and is insensitive; the
    Matsuoka-Nakai Newton stops at a RELATIVE RESIDUAL of 1e-10 (TOL_Radial_Returning,
    InOutFun/Material/Plasticity/Matsuoka-Nakai.c:82) on a 5x5 system whose reciprocal condition
    number the reference itself reports below 1e-12 on some points (Matsuoka-Nakai.c:670-676), so its
    internal variables are only defined to solver tolerance x conditioning: no arithmetic reproduces
    them to 1e-10 and the tolerance must follow the measured sensitivity."""
    P = load_problem("mn" if case == "ld" else case)
    P.materials = [(str(z["mat_type"]), z["mat_params"])]
    P.solver["tol_radial"] = float(z["tol_radial"])
    P.solver["maxiter_radial"] = int(z["maxiter_radial"])
    E = float(z["mat_params"][1])
    o = oracle.Oracle(P)
    rng = np.random.default_rng(7)
    X = z["inputs"]
    sens = np.zeros((len(X), len(GROUPS)))
    run = lambda x: _pack(o.stress_point(0, x[0:5], x[5:10], x[10], x[11:16], x[16], x[17]))
    for i, x in enumerate(X):
        base = run(x)
        for _ in range(4):
            r = run(x * (1 + 1e-15 * rng.standard_normal(x.shape)))
            for g, (sl, nm) in enumerate(GROUPS):
                with np.errstate(all="ignore"):
                    dev = np.abs(r[sl] - base[sl]).max() / _group_scale(base, sl, nm, E)
                sens[i, g] = max(sens[i, g], dev if np.isfinite(dev) else np.inf)
    return sens


@pytest.mark.parametrize("case", ("dp", "mn", "ld"))
def test_material_points_match_reference(case):
    """Constitutive update on the strain paths frozen from the reference (incl. its own test path).  ld = Lade-Duncan,
    the Matsuoka-Nakai return mapping with another yield surface, pinned on points only (its paths start pre-compressed:
    the reference's cloud runs of this law diverge from the unstressed state)."""
    z = load_points(case)
    X, Y = z["inputs"], z["outputs"]
    E = float(z["mat_params"][1])
    r = engine.stress_points(2, str(z["mat_type"]), z["mat_params"], float(z["tol_radial"]),
                             int(z["maxiter_radial"]), X[:, 0:5], X[:, 5:10], X[:, 10], X[:, 11:16], X[:, 16],
                             X[:, 17])
    assert np.all(r["status"] == 0)
    got = np.concatenate([r["stress"], r["b_e_n1"], r["eps_n1"][:, None], r["kappa_n1"][:, None], r["W"][:, None],
                          r["C_ep"]], axis=1)
    sens = _reference_sensitivity(case, z)
    # Drucker-Prager iterates its scalar Newton to 1e-14: every point must agree to 1e-10.
    # Matsuoka-Nakai stops its 5x5 Newton at a relative residual of 1e-10, so its internal variables
    # (Lambda = EPS, kappa, and C_ep which is a function of the last iterate) are only DEFINED to
    # solver tolerance x conditioning; stress, b_e and W are compared at 1e-10 + the measured
    # sensitivity of the reference's own answer, the internal variables at 1e-6 + sensitivity.
    loose = ("eps", "kappa", "C_ep") if case in ("mn", "ld") else ()
    if case == "dp":
        assert (sens < 1e-9).all(), float(sens.max())
    smax = sens.max(axis=1)      # a point whose Newton flips is unstable in every output
    for g, (sl, nm) in enumerate(GROUPS):
        s = np.array([_group_scale(y, sl, nm, E) for y in Y])
        fin = np.isfinite(Y[:, sl])
        stable = np.isfinite(sens[:, g]) & (sens[:, g] < 1e-9)
        assert stable.mean() > 0.9
        assert np.array_equal(np.isfinite(got[:, sl])[stable], fin[stable]), nm
        with np.errstate(all="ignore"):
            e = np.where(fin & np.isfinite(got[:, sl]), np.abs(got[:, sl] - Y[:, sl]) / s[:, None], 0.0).max(axis=1)
        tol = (1e-6 if nm in loose else RTOL) + 1000.0 * np.where(np.isfinite(smax), smax, 1e300)
        bad = e > tol
        # the MN tangent is the inverse of a matrix that goes singular towards the apex: tolerate a
        # handful of outliers there (3 of ~1100 points), nowhere else
        assert bad.sum() <= (3 if nm in loose else 0), (nm, int(bad.sum()), float(e[bad].max()))


@pytest.mark.parametrize("case", ["dp", "mn"])
def test_3d_material_points_against_the_compiled_reference(case):
    """The device's 3D Drucker-Prager / Matsuoka-Nakai updates on the strain paths frozen from the reference's OWN compiled
    3D laws (tests/golden/*_points3d.npz, oracle/ref_harness3d.c).  With quirk_transposed_eigvec = 1 (row-indexed
    eigenvectors in the plastic branches, SURVEY F10-i, which the compiled 3D laws do have) the kernel reproduces the
    reference; the engine's 3D default (0, the intended column form) agrees on the elastic steps only."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"{case}_points3d.npz"))
    X, Y = z["inputs"], z["outputs"]
    plastic = (Y[:, 18] != X[:, 28]) | (Y[:, 19] != X[:, 29])
    s = np.abs(Y[:, :9]).max(axis=1) + 1.0

    def run(quirk):
        r = engine.stress_points(3, str(z["mat_type"]), z["mat_params"], float(z["tol_radial"]), int(z["maxiter_radial"]),
                                 X[:, 0:9], X[:, 9:18], X[:, 18], X[:, 19:28], X[:, 28], X[:, 29], quirk=quirk)
        assert np.all(r["status"] == 0)
        return r
    r1 = run(1)
    d = np.abs(r1["stress"] - Y[:, :9]).max(axis=1) / s
    tol = 1e-10 if case == "dp" else 1e-7      # MN: 5x5 Newton stopped at 1e-10 on badly conditioned systems
    # a rotated plastic state whose trial eigenvalues nearly coincide has no stable eigenvector signs: allow a few
    assert (d > tol).sum() <= (0 if case == "dp" else 6), (int((d > tol).sum()), float(d.max()))
    de = np.abs(r1["eps_n1"] - Y[:, 18])
    if case == "dp":
        assert de.max() <= 1e-12
    else:   # MN's internal variables are defined to solver tolerance x conditioning only (see the 2D test above)
        assert (de > 1e-6).mean() <= 0.05, (int((de > 1e-6).sum()), float(de.max()))     # measured: 16 of 608, max 5e-4
    r0 = run(0)
    d0 = np.abs(r0["stress"] - Y[:, :9]).max(axis=1) / s
    assert d0[~plastic].max() <= 1e-9 and d0[plastic].max() > 1e-2


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


def test_async_download_snapshot_is_stream_ordered():
    """nlps_b200_run_async / _download_begin / _download_end / _sync: the snapshot taken after step k
    is what arrives on the host, although further steps were enqueued before the copy was issued."""
    P = load_problem("dp")
    eng = engine.Engine(P)
    assert eng.run(0, 5) == 0
    ref5 = eng.download()
    eng.close()
    eng = engine.Engine(P)
    assert eng.run_async(0, 5) == 0 and eng.sync() == 0
    eng.download_begin()
    assert eng.run_async(5, 20) == 0          # the GPU steps on while the snapshot waits for its copy
    got = eng.download_end()
    assert eng.sync() == 0
    for n in ("x_GC", "vel", "Stress", "F_n", "EPS_n", "I0", "NumberNodes"):
        assert np.array_equal(got[n], ref5[n]), n
    after = eng.download()
    assert not np.array_equal(after["x_GC"], ref5["x_GC"])
    ref = engine.Engine(P)
    assert ref.run(0, 25) == 0
    f25 = ref.download()
    assert np.array_equal(after["x_GC"], f25["x_GC"]) and np.array_equal(after["Stress"], f25["Stress"])
    ref.close()
    eng.close()


@pytest.mark.parametrize("sync_io", ("", "1"))
def test_u_verlet_results_callback_sees_its_own_step(sync_io, monkeypatch):
    """Results steps (TimeStep % ResultsTimeStep == 0, U-Verlet.c:1097): the download of step k overlaps the following
    steps (snapshot on the device, copies on their own stream); when cb(k) runs the host buffers must hold step k."""
    if sync_io:
        monkeypatch.setenv("NLPS_SYNC_IO", "1")
    P = load_problem("dp")
    P.solver["nsteps"] = 13
    for b in P.bounds:
        b["dir"], b["val"] = b["dir"][:, :13], b["val"][:, :13]
    P.gravity = P.gravity[:, :13]
    seen = {}

    def cb(k, host):
        seen[k] = {n: host[n].copy() for n in ("x_GC", "vel", "Stress", "EPS_n", "I0")}

    f = engine.u_verlet(P, results_every=4, callback=cb)
    assert sorted(seen) == [0, 4, 8, 12]
    eng = engine.Engine(P)
    done = 0
    for k in sorted(seen):
        assert eng.run(done, k + 1 - done) == 0
        done = k + 1
        g = eng.download()
        for n, v in seen[k].items():
            assert np.array_equal(v, g[n]), (k, n)
    for n in ("x_GC", "vel", "Stress", "EPS_n"):
        assert np.array_equal(f[n], g[n]), n
    eng.close()


@pytest.mark.parametrize("name", ("column2d_dp", "block2d_nh", "cube3d_nh", "cube3d_dp", "cube3d_mn", "cube3d_vm",
                                  "cube3d_hencky", "column2d_vm"))
def test_synthetic_clouds_against_oracle(name):
    """Synthetic inputs of the bench shapes (2D and 3D) through engine and oracle.  3D has no
    compilable reference: this is parity with the restatement, physics checks included."""
    from nlps_b200 import synthetic
    make = dict(column2d_dp=lambda: synthetic.column_collapse_2d(scale=0.03, nsteps=12),
                block2d_nh=lambda: synthetic.block_2d(cells=12, nsteps=12),
                cube3d_nh=lambda: synthetic.cube_3d(cells=5, nsteps=8),
                cube3d_dp=lambda: synthetic.cube_3d(cells=5, nsteps=8, material=synthetic.DP_C2),
                cube3d_mn=lambda: synthetic.cube_3d(cells=5, nsteps=8, material=synthetic.MN_C4),
                cube3d_vm=lambda: synthetic.cube_3d(cells=5, nsteps=12, material=synthetic.VM_SOFT),
                cube3d_hencky=lambda: synthetic.cube_3d(cells=5, nsteps=8, material=synthetic.HENCKY_C1),
                column2d_vm=lambda: synthetic.column_collapse_2d(scale=0.03, nsteps=12, material=synthetic.VM_SOFT))[name]
    P = make()
    if name.endswith("_vm"):  # a push that takes the cloud past the yield surface within the run
        P.fields["vel"][:, -1] = -0.1 * P.solver["cel"]
    n = P.nsteps
    eng = engine.Engine(P, compute_c_ep=1)
    o = oracle.Oracle(P)
    assert eng.initialize_lme() == 0 and o.init_lme() == 0
    counts, lists = eng.lists()
    assert np.array_equal(counts, o.ints("NumberNodes")) and np.array_equal(lists, o.lists())
    assert eng.run(0, n) == 0, eng.error()
    for k in range(n):
        assert o.step(k) == 0, o.error()
    f = eng.download()
    counts, lists = eng.lists()
    assert np.array_equal(f["I0"], o.ints("I0"))
    assert np.array_equal(counts, o.ints("NumberNodes")) and np.array_equal(lists, o.lists())
    assert np.array_equal(eng.active(), o.active())
    sc = field_scales(P)
    for nm in TRACE_FIELDS + (("Back_stress",) if name.endswith("_vm") else ()):
        assert_close(f[nm], o.field(nm), f"{name} {nm}", scale=sc.get(nm))
    for w, nm in enumerate(NODAL):
        assert_close(eng.nodal(w), o.nodal(w), f"{name} nodal {nm}", scale=sc["g" + nm])
    if name.endswith("_vm"):
        assert (f["EPS_n"] > 0).sum() > 20 and np.abs(f["Back_stress"]).max() > 0, "the Von-Mises cloud must yield"
    m0 = P.fields["mass"].sum()
    assert abs(eng.nodal(0)[:, 0].sum() - m0) <= 1e-12 * m0       # partition of unity on the device
    eng.close()


@pytest.mark.parametrize("case", ("c1", "c2twin"))
def test_config_shapes_against_the_reference(case):
    """BASELINE configs[0] at its stated shape (1024 particles x 200 steps) and the 1/8-scale twin of configs[1]
    (15,488 particles x 120 steps, half of them in plastic flow) against fixtures produced by the reference's OWN
    compiled code (tests/golden/make_golden.py config): closest nodes, neighbour counts, ordered lists and ActiveNode
    bit-exact, fields <= 1e-10."""
    import hashlib

    from util import CONFIG_FIELDS_SMALL, load_config
    P, g = load_config(case)
    eng = engine.Engine(P)
    assert eng.initialize_lme() == 0, eng.error()
    sc = field_scales(P)
    done = 0
    for cp in g["checkpoints"]:
        assert eng.run(done, int(cp) - done) == 0, eng.error()
        done = int(cp)
        t = f"s{done}_"
        f = eng.download()
        counts, lists = eng.lists()
        assert np.array_equal(f["I0"], g[t + "I0"]) and np.array_equal(counts, g[t + "NumberNodes"])
        assert np.array_equal(eng.active(), g[t + "active"])
        if t + "lists" in g.files:
            assert np.array_equal(lists, g[t + "lists"])
        else:
            assert hashlib.sha256(np.ascontiguousarray(lists).tobytes()).hexdigest() == str(g[t + "lists_sha256"])
        for nm in (TRACE_FIELDS if t + "W" in g.files else CONFIG_FIELDS_SMALL):
            if nm == "Kappa_n" and np.isnan(g[t + nm]).any():
                continue    # the reference leaves NaN in Kappa_n of a Neo-Hookean deck (kappa_0 is never parsed)
            assert_close(f[nm], g[t + nm], f"{case} step {done} {nm}", scale=sc.get(nm))
    eng.close()


@pytest.mark.parametrize("law", ("dp", "mn"))
def test_3d_plastic_cloud_with_the_reference_row_form(law):
    """A 3D plastic CLOUD with quirk_transposed_eigvec = 1, i.e. the reference's compiled behaviour (its plastic branches
    index the eigenvector matrix by row): engine against the oracle, whose 3D laws are pinned bit for bit to
    the reference's compiled 3D Drucker-Prager.c / Matsuoka-Nakai.c with the same flag (tests/test_oracle_3d_laws.py).
    The cloud is sheared so that the trial states have three distinct eigenvalues (row and column form then differ by
    O(1), and the eigenvectors are well conditioned)."""
    from nlps_b200 import synthetic
    mat = synthetic.DP_C2 if law == "dp" else synthetic.MN_C4
    P = synthetic.cube_3d(cells=5, nsteps=10, material=mat, compress=False)
    x = P.fields["x_GC"]
    cel = P.solver["cel"]
    P.fields["vel"][:, 0] = 0.04 * cel * (x[:, 2] - x[:, 2].min())           # simple shear in x-z ...
    P.fields["vel"][:, 1] = -0.02 * cel * (x[:, 0] - x[:, 0].mean())         # ... plus a twist: no two equal stretches
    P.fields["vel"][:, 2] = -0.05 * cel
    out = {}
    for quirk in (1, 0):
        # Matsuoka-Nakai in the row form: the stress then depends on the orientation the eigen-solver happens to return
        # for the (nearly degenerate) trial states of the first steps, and the 5x5 Newton stops at 1e-10 on systems the
        # reference itself reports as ill-conditioned (deviations 3 and 6): the difference between two correct
        # implementations grows with every step of this strongly sheared cloud (5e-6 in x after 10 steps in the column form,
        # 5e-4 in the row form), so Matsuoka-Nakai is compared after 3 steps (column form) / ONE step (row form) at 1e-6
        nsteps = (1 if quirk == 1 else 3) if law == "mn" else P.nsteps
        eng = engine.Engine(P, quirk=quirk)
        o = oracle.Oracle(P)
        o.set_flags(quirk, 0)
        assert eng.initialize_lme() == 0 and o.init_lme() == 0
        assert eng.run(0, nsteps) == 0, eng.error()
        for k in range(nsteps):
            assert o.step(k) == 0, o.error()
        f = eng.download()
        sc = field_scales(P)
        tol = 1e-10 if law == "dp" else 1e-6      # Matsuoka-Nakai: Newton stopped at 1e-10 on ill-conditioned systems
        for nm in ("x_GC", "vel", "F_n", "Stress", "b_e_n", "EPS_n", "Kappa_n"):
            assert_close(f[nm], o.field(nm), f"3D {law} cloud, quirk {quirk}: {nm}", rtol=tol, scale=sc.get(nm))
        out[quirk] = f
        eng.close()
    if law == "mn":
        return
    assert (out[1]["EPS_n"] > (P.fields["EPS_n"] if law == "mn" else 0)).sum() > 20       # plastic particles exist
    # and the two forms do differ on this cloud: the test above is not vacuous
    assert np.abs(out[1]["Stress"] - out[0]["Stress"]).max() > 1e-6 * np.abs(out[0]["Stress"]).max()
