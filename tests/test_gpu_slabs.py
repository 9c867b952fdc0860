"""Multi-slab engine against the single-slab engine, through the C ABI.

Single GPU: the slabs run as threads of this process and exchange through the custom-transport
hook of the ABI (device-to-device copies behind a barrier) -- the same kernels, halo lists,
migration and buffers as with NCCL.  With >= 2 GPUs the NCCL transport is exercised through
torchrun.  Bar: integer outputs bit-exact, fields <= 1e-10 (the nodal sums of the shared nodes are
added in a different order than on one slab)."""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

from nlps_b200 import engine
from slabcases import COMPARE, merge, moving_block, sinking_column
from util import assert_close, field_scales

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_single(P, nsteps):
    eng = engine.Engine(P, device=0)
    assert eng.initialize_lme() == 0
    assert eng.run(0, nsteps) == 0, eng.error()
    f = eng.download()
    counts, lists = eng.lists()
    act = eng.active()
    eng.close()
    return f, counts, lists, act


def run_slabs_threads(P, nsteps, world, migrate_every, device=0, per_rank=None):
    """per_rank: [(Problem, slab dict)] when every slab brings its own sub-mesh and particles."""
    axis, cuts = engine.slab_cuts(P, world) if per_rank is None else (per_rank[0][1]["axis"], per_rank[0][1]["cuts"])
    comms = engine.ThreadComm.group(world)
    res, errs = [None] * world, []

    def work(r):
        try:
            if per_rank is None:
                eng = engine.Engine(P, device=device, slab=dict(rank=r, world=world, axis=axis, cuts=cuts,
                                                                comm=comms[r], migrate_every=migrate_every))
            else:
                eng = engine.Engine(per_rank[r][0], device=device,
                                    slab=dict(per_rank[r][1], comm=comms[r], migrate_every=migrate_every))
            n0 = eng.local_count()
            assert eng.initialize_lme() == 0, eng.error()
            assert eng.run(0, nsteps) == 0, eng.error()
            f, ids = eng.download_local()
            if per_rank is None:
                counts, lists = eng.lists()
            else:  # compact population: lists by local node ids -> global ids, rows scattered by particle id
                counts, lists = local_lists(eng, per_rank[r][1], ids, P.np_)
            res[r] = (f, ids, counts, lists, n0, eng.migrated_count(), eng.active())
            eng.close()
        except BaseException as ex:  # noqa: BLE001 -- release the other slabs, then report
            errs.append((r, ex))
            comms[r].sh.barrier.abort()

    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(600)
    for c in comms:
        c.close()
    assert not errs, errs
    return res, axis, cuts


def local_lists(eng, slab, ids, n_global):
    """neighbour lists of a sub-mesh slab in the global frame (node ids + offset, rows by particle id)"""
    import ctypes as C
    n = eng.local_count()
    cap = eng.L.nlps_b200_list_capacity(eng.h)
    # get_lists of a slab engine indexes rows by global id: hand it global-size arrays
    counts = np.zeros(n_global, np.int32)
    lists = np.full((n_global, cap), -7, np.int32)
    ip = C.POINTER(C.c_int)
    assert eng.L.nlps_b200_get_lists(eng.h, counts.ctypes.data_as(ip), lists.ctypes.data_as(ip), cap) == 0
    rows = lists[ids]
    rows[rows >= 0] += slab["node_offset"]
    lists[ids] = rows
    return counts, lists


def compare(P, single, res, check_active=True):
    f1, c1, l1, act1 = single
    m = merge([r[:4] for r in res], P.np_)
    assert np.array_equal(m["I0"], f1["I0"])
    assert np.array_equal(m["NumberNodes"], f1["NumberNodes"])
    assert np.array_equal(m["_counts"], c1) and np.array_equal(m["_lists"], l1)
    sc = field_scales(P)
    for k in COMPARE + (("Back_stress",) if "Back_stress" in f1 else ()):
        assert_close(m[k], f1[k], "slabs vs single: " + k, scale=sc.get(k))
    if check_active:
        act = np.zeros_like(act1)
        for r in res:
            act |= r[6]
        assert np.array_equal(act, act1)   # union of the slabs' ActiveNode views == Mesh.ActiveNode


@pytest.mark.parametrize("world", [2, 3])
def test_moving_block_slabs_match_single(world):
    nsteps = 60
    P = moving_block(nsteps=nsteps)
    single = run_single(P, nsteps)
    res, axis, cuts = run_slabs_threads(P, nsteps, world, migrate_every=4)
    assert axis == 0
    assert sum(r[5] for r in res) > 100, "the block must have crossed the cuts"
    compare(P, single, res)
    # the populations moved downstream
    assert res[-1][0]["x_GC"].shape[0] > res[-1][4]


def test_plastic_column_slabs_match_single():
    nsteps = 50
    P = sinking_column(nsteps=nsteps)
    single = run_single(P, nsteps)
    assert (single[0]["EPS_n"] > 0).sum() > 50
    res, axis, cuts = run_slabs_threads(P, nsteps, 2, migrate_every=3)
    assert axis == 1
    assert sum(r[5] for r in res) > 0
    compare(P, single, res)


def test_von_mises_column_slabs_match_single():
    """Von-Mises with kinematic hardening: the back stress (Phi.Back_stress, 3 extra columns) travels with the migrating
    particles and through the re-sorts."""
    from nlps_b200 import synthetic
    nsteps = 50
    P = synthetic.column_collapse_2d(scale=0.1, nsteps=nsteps, material=synthetic.VM_SOFT)
    P.fields["vel"][:, 1] = -0.12 * P.solver["cel"]
    single = run_single(P, nsteps)
    assert (single[0]["EPS_n"] > 0).sum() > 50 and np.abs(single[0]["Back_stress"]).max() > 0
    res, axis, cuts = run_slabs_threads(P, nsteps, 2, migrate_every=3)
    assert axis == 1 and sum(r[5] for r in res) > 0
    compare(P, single, res)


def test_submesh_slabs_match_global_engine():
    """The weak-scaling set-up of bench.py: every slab builds only its sub-mesh (cut +- band) and its own
    particles with global ids; three slabs reproduce the engine that holds the whole tall column."""
    from nlps_b200 import synthetic
    world, scale, nsteps = 3, 0.06, 40
    bx = max(4, int(round(354 * scale)))
    by = 2 * bx
    G = synthetic.structured_problem(2, (6 * bx, by * world + by // 4), 0.2 / bx, (bx, by * world), (0, 0),
                                     synthetic.DP_C2, nsteps, 0.5, (1e7 / 2000.0) ** 0.5 * 1.3, (0.0, -9.81))
    G.fields["vel"][:, 1] = -0.12 * G.solver["cel"]
    per_rank = []
    for r in range(world):
        Pr, sl = synthetic.column_slab_2d(r, world, scale=scale, nsteps=nsteps)
        Pr.fields["vel"][:, 1] = -0.12 * Pr.solver["cel"]
        per_rank.append((Pr, sl))
    single = run_single(G, nsteps)
    assert (single[0]["EPS_n"] > 0).sum() > 20
    res, axis, cuts = run_slabs_threads(G, nsteps, world, migrate_every=3, per_rank=per_rank)
    # I0 of a sub-mesh slab is a local node id
    for r, (Pr, sl) in zip(res, per_rank):
        r[0]["I0"] = r[0]["I0"] + sl["node_offset"]
    assert sum(r[5] for r in res) > 0
    compare(G, single, res, check_active=False)


def test_cube_3d_slabs_match_single():
    """3D (4 mask words, 125-node rings): Neo-Hookean cube drifting along x through two cuts."""
    from nlps_b200 import synthetic
    nsteps = 40
    P = synthetic.structured_problem(3, (44, 10, 10), 1.0 / 8, (32, 6, 6), (4, 2, 2), synthetic.NH_C1, nsteps, 0.5,
                                     (1e6 / 1000.0) ** 0.5 * 1.3, (0.0, 0.0, -9.81), gamma_lme=6.0, fixed=("bottom",),
                                     rollers=())
    z = P.fields["x_GC"][:, 2]
    P.fields["vel"][:, 0] = 0.3 * P.solver["cel"] * (1.0 + 0.2 * (z - z.min()))
    single = run_single(P, nsteps)
    res, axis, cuts = run_slabs_threads(P, nsteps, 2, migrate_every=4)
    assert axis == 0 and sum(r[5] for r in res) > 100
    compare(P, single, res)


def test_submesh_slabs_3d_match_global_engine():
    """bench.py --workload c3: z slabs of the cube, each on its own sub-mesh, against the engine holding the cube."""
    from nlps_b200 import synthetic
    world, cells, nsteps = 2, 20, 30
    G = synthetic.cube_3d(cells=cells, nsteps=nsteps)
    G.fields["vel"][:, 2] = -0.15 * G.solver["cel"]
    per_rank = []
    for r in range(world):
        Pr, sl = synthetic.cube_slab_3d(r, world, cells=cells, nsteps=nsteps)
        Pr.fields["vel"][:, 2] = -0.15 * Pr.solver["cel"]
        per_rank.append((Pr, sl))
    single = run_single(G, nsteps)
    res, axis, cuts = run_slabs_threads(G, nsteps, world, migrate_every=3, per_rank=per_rank)
    for r, (Pr, sl) in zip(res, per_rank):
        r[0]["I0"] = r[0]["I0"] + sl["node_offset"]
    assert axis == 2 and sum(r[5] for r in res) > 0
    compare(G, single, res, check_active=False)


@pytest.mark.parametrize("law", ["nh_drift", "mn_gravity"])
def test_slope_slabs_match_global_engine(law):
    """bench.py --workload c4 (BASELINE configs[3]): the slope, slabs along it with cuts at particle-count quantiles, each
    slab on its own sub-mesh, global particle ids with gaps (carved cells), migration every 3 steps -- against one engine
    holding the whole slope.  nh_drift: a Neo-Hookean slope pushed up-slope so that ~900 particles cross the cut (the
    stress-free Matsuoka-Nakai state sits next to the apex and does not survive a velocity kick);
    mn_gravity: the Matsuoka-Nakai slope of the bench under its gravity ramp."""
    from nlps_b200 import synthetic
    world, cells, width, nsteps = 2, 40, 5, 24
    mat = synthetic.NH_C1 if law == "nh_drift" else synthetic.MN_C4

    def make(r, w):
        P_, sl_ = synthetic.slope_slab_3d(r, w, cells=cells, width=width, nsteps=nsteps, ramp_steps=10, material=mat,
                                          band_cells=5)      # as bench.py --workload c4
        if law == "nh_drift":
            P_.fields["vel"][:, 2] = -0.15 * P_.solver["cel"] * np.clip((P_.fields["x_GC"][:, 2] - 0.2) / 0.4, 0.0, 1.0)
        return P_, sl_
    G, _ = make(0, 1)
    gid_G = (G.kept_cells.astype(np.int64)[:, None] * 8 + np.arange(8)[None, :]).ravel()
    per_rank = [make(r, world) for r in range(world)]
    per_rank = [(Pr, {k: v for k, v in sl.items() if k != "n_particles"}) for Pr, sl in per_rank]
    assert per_rank[0][1]["n_global"] > G.np_                 # the id space has gaps
    single = run_single(G, nsteps)

    class _Ids:                                               # run_slabs_threads sizes its list arrays by the id space
        np_ = per_rank[0][1]["n_global"]
    res, axis, cuts = run_slabs_threads(_Ids, nsteps, world, migrate_every=3, per_rank=per_rank)
    assert axis == 2
    if law == "nh_drift":
        assert sum(r[5] for r in res) > 100                   # particles did migrate
    compact = []
    for (f, ids, counts, lists, *rest), (Pr, sl) in zip(res, per_rank):
        rows = np.searchsorted(gid_G, ids)
        assert np.array_equal(gid_G[rows], ids)
        f["I0"] = f["I0"] + sl["node_offset"]
        cc = np.zeros(G.np_, np.int32)
        ll = np.full((G.np_, lists.shape[1]), -7, np.int32)
        cc[rows], ll[rows] = counts[ids], lists[ids]
        compact.append((f, rows, cc, ll, *rest))
    compare(G, single, compact, check_active=False)
    assert np.abs(single[0]["Stress"]).max() > 0.0
    # ... and against the ORACLE (the CPU restatement pinned to the reference's compiled laws), not only against a second
    # engine: the whole slope, same steps, the 1e-10 bar.
    import oracle
    o = oracle.Oracle(G)
    assert o.init_lme() == 0
    for k in range(nsteps):
        assert o.step(k) == 0, o.error()
    m = merge([c[:4] for c in compact], G.np_)
    assert np.array_equal(m["I0"], o.ints("I0")) and np.array_equal(m["_counts"], o.ints("NumberNodes"))
    assert np.array_equal(m["_lists"], o.lists())
    sc = field_scales(G)
    rtol = 1e-10  # (the Matsuoka-Nakai slope stays elastic over these steps: no ill-conditioned return mapping is involved)
    for kf in COMPARE:
        assert_close(m[kf], o.field(kf), f"slope slabs vs oracle ({law}): {kf}", rtol=rtol, scale=sc.get(kf))


def test_excursion_is_latched():
    """Without migration a particle eventually leaves the band its slab may roam in: error 9."""
    nsteps = 60
    P = moving_block(nsteps=nsteps)
    axis, cuts = engine.slab_cuts(P, 2)
    comms = engine.ThreadComm.group(2)
    codes = [None, None]

    def work(r):
        eng = engine.Engine(P, device=0, slab=dict(rank=r, world=2, axis=axis, cuts=cuts, comm=comms[r],
                                                   migrate_every=10 ** 6))
        eng.initialize_lme()
        rc = eng.run(0, nsteps)
        codes[r] = (rc, eng.error()[0])
        eng.close()

    th = [threading.Thread(target=work, args=(r,)) for r in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join(600)
    assert (1, 9) in codes, codes


@pytest.mark.parametrize("world", [2, 3])
def test_migration_capacity_fails_on_every_slab_without_hanging(world):
    """A slab without room for the particles that arrive must not leave its neighbours in a half-finished row exchange:
    with capacity_factor ~ 1 the block flying downstream overflows the last slab, EVERY slab returns EXIT_FAILURE with
    NLPS_ERR_SLAB_CAPACITY at the same migration, and nobody hangs."""
    nsteps = 60
    P = moving_block(nsteps=nsteps)
    axis, cuts = engine.slab_cuts(P, world)
    comms = engine.ThreadComm.group(world)
    out = [None] * world

    def work(r):
        try:
            eng = engine.Engine(P, device=0, slab=dict(rank=r, world=world, axis=axis, cuts=cuts, comm=comms[r],
                                                       migrate_every=4, capacity_factor=1.0 + 1e-9))
            assert eng.initialize_lme() == 0
            rc = 0
            for k0 in range(0, nsteps, 4):   # the slabs stop together, at the migration that fails
                rc = eng.run(k0, 4)
                if rc != 0:
                    break
            out[r] = (rc, eng.error()[0], k0)
            eng.close()
        except BaseException as ex:  # noqa: BLE001
            out[r] = ("exception", repr(ex))
            comms[r].sh.barrier.abort()

    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(300)
    assert not any(t.is_alive() for t in th), "a slab hangs in the exchange"
    for c in comms:
        c.close()
    assert all(o is not None and o[0] == 1 and o[1] == 10 for o in out), out   # NLPS_ERR_SLAB_CAPACITY everywhere
    assert len({o[2] for o in out}) == 1, out                                    # ... at the same step


def test_slabs_nccl_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29741", os.path.join(ROOT, "tests", "workers", "slab_nccl_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "NCCL slabs OK" in out.stdout


@pytest.mark.parametrize("case,world", [("block2d", 2), ("block2d", 3), ("cantilever3d", 2), ("cantilever3d", 3)])
def test_implicit_newmark_slabs_match_single(case, world):
    """SURVEY 8(e) "Implicit": every slab assembles the tangent of its own particles, the Krylov vectors are summed over
    the band nodes each iteration and the dot products over the slabs.  Converged Newmark-beta steps at several times
    the explicit time step on 2 / 3 slabs (particles migrate in between) against the engine that holds the whole cloud;
    tolerance = what a Newton loop stopped at |R| <= 1e-12 |R0| supports."""
    from nlps_b200 import synthetic
    nsteps = 6
    if case == "block2d":
        P = synthetic.structured_problem(2, (64, 16), 1.0 / 16, (48, 8), (8, 0), synthetic.NH_C1, nsteps, 0.5,
                                         (1e6 / 1000.0) ** 0.5 * 1.3, (0.0, -9.81), fixed=("bottom",), rollers=())
        P.fields["vel"][:, 0] = 0.05 * P.solver["cel"]
    elif case == "cantilever3d":  # BASELINE configs[4] shape: clamped beam with a tip traction (Neumann load on Area_0)
        P = synthetic.beam_3d(cells_per_unit=4 if world == 2 else 5, nsteps=nsteps)
    if case != "cantilever3d":
        P.solver["cfl"] = 4.0
    kw = dict(tol=1e-12, max_iter=25, pcg_rtol=1e-13)

    eng = engine.Engine(P, device=0)
    assert eng.initialize_lme() == 0 and eng.newmark_setup(**kw) == 0
    assert eng.newmark_run(0, nsteps) == 0, eng.error()
    f1 = eng.download()
    c1, l1 = eng.lists()
    eng.close()

    axis, cuts = engine.slab_cuts(P, world)
    comms = engine.ThreadComm.group(world)
    res, errs = [None] * world, []

    def work(r):
        try:
            e = engine.Engine(P, device=0, slab=dict(rank=r, world=world, axis=axis, cuts=cuts, comm=comms[r], migrate_every=2))
            assert e.initialize_lme() == 0, e.error()
            assert e.newmark_setup(**kw) == 0
            assert e.newmark_run(0, nsteps) == 0, e.error()
            f, ids = e.download_local()
            counts, lists = e.lists()
            res[r] = (f, ids, counts, lists, e.newmark_stats(), e.migrated_count())
            e.close()
        except BaseException as ex:  # noqa: BLE001
            errs.append((r, ex))
            comms[r].sh.barrier.abort()

    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(900)
    for c in comms:
        c.close()
    assert not errs, errs
    assert axis == 0
    m = merge([r[:4] for r in res], P.np_)
    assert np.array_equal(m["I0"], f1["I0"]) and np.array_equal(m["_counts"], c1) and np.array_equal(m["_lists"], l1)
    sc = field_scales(P)
    for k in ("x_GC", "dis", "vel", "acc", "F_n", "Stress", "rho", "J_n", "lambda"):
        assert_close(m[k], f1[k], f"implicit slabs vs single ({case}, {world} slabs): {k}", rtol=1e-8, scale=sc.get(k))
    # the slabs agreed on every solver decision: same Newton and Krylov iteration counts on all of them
    its = {(r[4]["newton_iters"], r[4]["pcg_iters_total"]) for r in res}
    assert len(its) == 1, its
    if case == "block2d":
        assert sum(r[5] for r in res) > 0, "particles must have crossed the cuts"
