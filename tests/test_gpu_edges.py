"""Edge cases of the engine through the C ABI: mixed material laws in one cloud (per-particle dispatch), Neumann
tractions, particles sitting exactly on nodes / cell faces (ties of the closest-node search), jittered
(ragged) clouds, an initially EMPTY slab that fills by migration, a one-particle cloud, upload/download round
trips, invalid inputs."""
import threading

import numpy as np
import pytest

import oracle
from nlps_b200 import engine, synthetic
from slabcases import COMPARE, merge, moving_block
from util import TRACE_FIELDS, assert_close, field_scales

pytestmark = pytest.mark.gpu


def _run_both(P, nsteps):
    eng = engine.Engine(P, device=0)
    o = oracle.Oracle(P)
    assert eng.initialize_lme() == 0 and o.init_lme() == 0
    assert eng.run(0, nsteps) == 0, eng.error()
    for k in range(nsteps):
        assert o.step(k) == 0, o.error()
    f = eng.download()
    counts, lists = eng.lists()
    assert np.array_equal(f["I0"], o.ints("I0"))
    assert np.array_equal(lists, o.lists())
    assert np.array_equal(eng.active(), o.active())
    sc = field_scales(P)
    for name in TRACE_FIELDS:
        assert_close(f[name], o.field(name), name, scale=sc.get(name))
    eng.close()
    return f


def test_mixed_materials_dispatch_per_particle():
    P = synthetic.column_collapse_2d(scale=0.05, nsteps=30)
    P.materials = [synthetic.DP_C2, synthetic.NH_C1, ("Matsuoka-Nakai", np.asarray(synthetic.MN_C4[1], float))]
    P.materials = [(t, np.asarray(p, float)) for t, p in P.materials]
    x = P.fields["x_GC"]
    P.MatIdx = (np.floor(x[:, 1] / (x[:, 1].max() / 3 + 1e-12)).clip(0, 2)).astype(np.int32)
    P.solver["tol_radial"], P.solver["maxiter_radial"] = 1e-10, 20
    P.fields["vel"][:, 1] = -0.05 * P.solver["cel"]
    f = _run_both(P, 30)
    assert len(np.unique(P.MatIdx)) == 3


def test_neumann_traction_on_a_block():
    nsteps = 25
    P = synthetic.block_2d(cells=8, nsteps=nsteps)
    x = P.fields["x_GC"]
    top = np.nonzero(x[:, 1] > x[:, 1].max() - 0.3 * P.dx)[0].astype(np.int32)
    dr = np.zeros((2, nsteps), np.int32)
    dr[1, :] = 1
    val = np.zeros((2, nsteps))
    val[1, :] = -5e3 * np.linspace(0.2, 1.0, nsteps)
    P.neumann.append(dict(nodes=top, dir=dr, val=val))
    half = top[: len(top) // 2]                      # a second load overlapping the first one
    P.neumann.append(dict(nodes=half, dir=np.ones((2, nsteps), np.int32), val=np.full((2, nsteps), 1e3)))
    _run_both(P, nsteps)


def test_neumann_traction_3d_uses_area_0():
    """3D Neumann loads act on Phi.Area_0 (U-Verlet.c:847-849), not on Vol_0 / Thickness_Plain_Stress: a cube pushed
    down on its top particle layer, against the oracle; without Area_0 the engine refuses the deck."""
    nsteps = 12
    P = synthetic.cube_3d(cells=6, nsteps=nsteps, compress=False)
    x = P.fields["x_GC"]
    top = np.nonzero(x[:, 2] > x[:, 2].max() - 0.3 * P.dx)[0].astype(np.int32)
    dr = np.zeros((3, nsteps), np.int32)
    dr[2, :] = 1
    val = np.zeros((3, nsteps))
    val[2, :] = -2e4 * np.linspace(0.2, 1.0, nsteps)
    P.neumann.append(dict(nodes=top, dir=dr, val=val))
    with pytest.raises(RuntimeError, match="Area_0"):
        engine.Engine(P, device=0)
    P.fields["Area_0"] = np.full(P.np_, 0.25 * P.dx ** 2) * (1.0 + 0.1 * np.sin(np.arange(P.np_)))   # not Vol_0 / h
    f = _run_both(P, nsteps)
    assert np.abs(f["acc"][top, 2]).max() > 0.0


def test_two_engines_with_different_materials_in_one_process():
    """The material table belongs to the engine (a kernel parameter), not to the process: a second engine with another
    deck, and a call of the point-wise stress entry in between, leave the first engine's law untouched."""
    nsteps = 10
    Pa = synthetic.column_collapse_2d(scale=0.04, nsteps=nsteps)
    Pb = synthetic.block_2d(cells=8, nsteps=nsteps)                       # Neo-Hookean, other constants
    for Q in (Pa, Pb):
        Q.fields["vel"][:, 1] = -0.05 * Q.solver["cel"]
    ea = engine.Engine(Pa, device=0)
    assert ea.initialize_lme() == 0 and ea.run(0, nsteps // 2) == 0
    eb = engine.Engine(Pb, device=0)                                       # used to overwrite the process-wide table
    assert eb.initialize_lme() == 0 and eb.run(0, nsteps) == 0
    ident = np.tile(np.array([1.0, 0.0, 0.0, 1.0, 1.0]), (4, 1))            # 2D tensors: 2x2 block + slot 4
    engine.stress_points(2, "Neo-Hookean-Wriggers", synthetic.NH_C1[1], 1e-14, 10, ident, ident, np.ones(4), ident,
                         np.zeros(4), np.zeros(4))
    assert ea.run(nsteps // 2, nsteps - nsteps // 2) == 0, ea.error()
    fa = ea.download()
    oa = oracle.Oracle(Pa)
    assert oa.init_lme() == 0
    for k in range(nsteps):
        assert oa.step(k) == 0
    sc = field_scales(Pa)
    for name in TRACE_FIELDS:
        assert_close(fa[name], oa.field(name), "engine A after engine B: " + name, scale=sc.get(name))
    ea.close()
    eb.close()


def test_particles_on_nodes_and_faces_tie_breaking():
    """x exactly on grid lines: equal distances to several nodes; the first strict minimum in chain order wins."""
    P = synthetic.structured_problem(2, (10, 10), 0.125, (6, 6), (2, 2), synthetic.NH_C1, 12, 0.5,
                                     (1e6 / 1000.0) ** 0.5 * 1.3, (0.0, -9.81))   # block away from the hull of the nodes
    x = P.fields["x_GC"]
    h = P.dx
    x[:] = np.round(x / (0.5 * h)) * (0.5 * h)        # snap to half-cell lattice: nodes, edge and cell centres
    P.fields["dis"][:] = 1e-300                       # makes the search run (|dis| > 0) without moving anything
    d2 = ((x[:, None, :] - P.coords[None, :, :]) ** 2).sum(-1)
    P.I0 = d2.argmin(1).astype(np.int32)
    P.fields["vel"][:, 0] = 0.01 * P.solver["cel"]
    _run_both(P, 12)


def test_jittered_cloud_3d():
    P = synthetic.structured_problem(3, (9, 9, 9), 0.125, (4, 4, 4), (2, 2, 1), synthetic.NH_C1, 10, 0.5,
                                     (1e6 / 1000.0) ** 0.5 * 1.3, (0.0, 0.0, -9.81), gamma_lme=6.0, jitter=0.2,
                                     rollers=("left", "right", "front", "back"))
    _run_both(P, 10)


def test_one_particle():
    P = synthetic.block_2d(cells=4, nsteps=5)
    keep = np.array([P.np_ // 2])
    P.fields = {k: np.ascontiguousarray(v[keep]) for k, v in P.fields.items()}
    P.I0, P.MatIdx = P.I0[keep].copy(), P.MatIdx[keep].copy()
    _run_both(P, 5)


def test_upload_download_round_trip():
    P = synthetic.block_2d(cells=6, nsteps=4)
    eng = engine.Engine(P, device=0)
    assert eng.initialize_lme() == 0 and eng.run(0, 2) == 0
    f = eng.download()
    rng = np.random.default_rng(3)
    new = {k: f[k] + 1e-3 * rng.standard_normal(f[k].shape) for k in ("vel", "acc", "Stress", "rho")}
    eng.upload(new)
    g = eng.download()
    for k, v in new.items():
        assert np.array_equal(g[k], v)
    for k in ("x_GC", "F_n", "lambda"):
        assert np.array_equal(g[k], f[k])
    eng.close()


def test_empty_slab_fills_by_migration():
    """cuts chosen so that the upper slab starts with no particle at all; the block flies into it."""
    nsteps = 60
    P = moving_block(nsteps=nsteps)
    e1 = engine.Engine(P, device=0)
    assert e1.initialize_lme() == 0 and e1.run(0, nsteps) == 0
    f1 = e1.download()
    c1, l1 = e1.lists()
    e1.close()
    x = P.coords[P.I0, 0]
    cuts = np.array([x.max() + 0.5 * P.dx])           # everything below the cut
    comms = engine.ThreadComm.group(2)
    res, errs = [None, None], []

    def work(r):
        try:
            eng = engine.Engine(P, device=0, slab=dict(rank=r, world=2, axis=0, cuts=cuts, comm=comms[r], migrate_every=3,
                                                       capacity_factor=2.0))
            n0 = eng.local_count()
            assert eng.initialize_lme() == 0 and eng.run(0, nsteps) == 0, eng.error()
            f, ids = eng.download_local()
            counts, lists = eng.lists()
            res[r] = (f, ids, counts, lists, n0, eng.local_count())
            eng.close()
        except BaseException as ex:  # noqa: BLE001
            errs.append((r, ex))
            comms[r].sh.barrier.abort()

    th = [threading.Thread(target=work, args=(r,)) for r in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join(600)
    assert not errs, errs
    assert res[1][4] == 0 and res[1][5] > 100          # empty at the start, populated at the end
    m = merge([r[:4] for r in res], P.np_)
    assert np.array_equal(m["I0"], f1["I0"]) and np.array_equal(m["_lists"], l1)
    sc = field_scales(P)
    for k in COMPARE:
        assert_close(m[k], f1[k], "empty slab: " + k, scale=sc.get(k))


def test_invalid_inputs_are_refused():
    P = synthetic.block_2d(cells=4, nsteps=2)
    bad = synthetic.block_2d(cells=4, nsteps=2)
    bad.I0 = bad.I0.copy()
    bad.I0[0] = bad.nn + 5
    with pytest.raises(RuntimeError, match="I0 out of range"):
        engine.Engine(bad, device=0)
    bad = synthetic.block_2d(cells=4, nsteps=2)
    bad.MatIdx = bad.MatIdx.copy()
    bad.MatIdx[3] = 7
    with pytest.raises(RuntimeError, match="MatIdx out of range"):
        engine.Engine(bad, device=0)
    with pytest.raises(RuntimeError, match="slab"):
        engine.Engine(P, device=0, slab=dict(rank=2, world=2, axis=0, cuts=np.array([0.5]), comm=None))
