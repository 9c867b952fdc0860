"""world_size-2 (gloo, CPU) emulation of the slab protocol of nlps_b200 with the ORACLE as
the per-slab stepper: same host planning (cuts, owners, halo node lists) as the engine, same three
exchanges per step (ActiveNode flags, M + momentum sums, force sums), checked against the
single-domain oracle.  Launched by tests/test_slabs_cpu.py through torch.distributed.run."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in ("nl-partsol_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))

import oracle  # noqa: E402
from nlps_b200 import engine, synthetic  # noqa: E402
from nlps_b200.problem import Problem  # noqa: E402
from util import assert_close, field_scales  # noqa: E402


def subset(P: Problem, rows):
    Q = Problem(ndim=P.ndim, coords=P.coords, r1p=P.r1p, r1i=P.r1i, r2p=P.r2p, r2i=P.r2i, h_avg=P.h_avg, dx=P.dx,
                solver=dict(P.solver), gravity=P.gravity, bounds=P.bounds, neumann=[], materials=P.materials)
    Q.fields = {k: np.ascontiguousarray(v[rows]) for k, v in P.fields.items()}
    Q.I0 = np.ascontiguousarray(P.I0[rows])
    Q.MatIdx = np.ascontiguousarray(P.MatIdx[rows])
    return Q


def exchange_sum(arr, halo, peer_rank):
    """add the peer's values on the halo nodes (both sides use the same ascending id list)"""
    mine = torch.from_numpy(np.ascontiguousarray(arr[halo]))
    other = torch.zeros_like(mine)
    reqs = [dist.isend(mine, peer_rank), dist.irecv(other, peer_rank)]
    for r in reqs:
        r.wait()
    out = arr.copy()
    out[halo] = arr[halo] + other.numpy()
    return out


def implicit_main(rank, world):
    """SURVEY 8(e) "Implicit": the protocol of csrc/nlps_implicit.inl over two slabs, with the oracle's Newmark-beta
    stages as the per-slab arithmetic -- every slab holds the tangent of ITS particles only (K = sum of K_s), the mass
    term a1 M and the unit rows of restricted / inactive dofs are the share of the slab that owns the node, Krylov
    vectors are consistent on the band nodes (band sums of y_s = K_s p after every product), p.Ap is summed over ALL
    local rows before the exchange, r.z and r.r over owned rows, then over the slabs.  Checked against the single-domain
    oracle (dense LU)."""
    nsteps = 4
    P = synthetic.structured_problem(2, (40, 12), 1.0 / 8, (28, 6), (6, 0), synthetic.NH_C1, nsteps, 0.5,
                                     (1e6 / 1000.0) ** 0.5 * 1.3, (0.0, -9.81), fixed=("bottom",), rollers=())
    P.fields["vel"][:, 0] = 0.05 * P.solver["cel"]
    P.solver["cfl"] = 4.0
    axis, cuts = engine.slab_cuts(P, world)
    assert axis == 0 and len(cuts) == 1
    owner = engine.slab_owner(P, axis, cuts)
    rows = np.nonzero(owner == rank)[0]
    halo = engine.slab_halo_nodes(P, axis, cuts[0], 6)
    assert 0 < len(halo) < P.nn
    peer, d, nn = 1 - rank, P.ndim, P.nn
    own = ((P.coords[:, axis] < cuts[0]) if rank == 0 else (P.coords[:, axis] >= cuts[0]))[:, None] & np.ones((1, d), bool)
    kw = dict(tol=1e-12, max_iter=25)

    o = oracle.Oracle(subset(P, rows))
    o.newmark_setup(**kw)
    a1, a2, a3, a4, a5, a6 = o.newmark_coeffs()

    def merge_active():
        a = o.active()
        t = torch.from_numpy(a[halo].copy())
        u = torch.zeros_like(t)
        for r in [dist.isend(t, peer), dist.irecv(u, peer)]:
            r.wait()
        a[halo] |= u.numpy()
        o.set_active(a)

    def allsum(x):
        t = torch.tensor([float(x)], dtype=torch.float64)
        dist.all_reduce(t)
        return float(t[0])

    assert o.search_closest() == 0
    merge_active()
    assert o.search_lists() == 0
    newton_total = 0
    for k in range(nsteps):
        assert o.search_closest() == 0
        merge_active()
        assert o.search_lists() == 0
        assert o.newmark_begin_after_search(k) == 0
        # lumped mass and the nodal projections of v_n, a_n: numerators and M are band-summed, then divided
        Ml, Vl, Al = o.nodal(0), o.newmark_get("Vn"), o.newmark_get("An")
        M = exchange_sum(Ml, halo, peer)
        # (a node that only the neighbour's particles reach is active with a local mass of zero: 0 / 0 in the stage)
        numV, numA = np.where(Ml != 0, Vl * Ml, 0.0), np.where(Ml != 0, Al * Ml, 0.0)
        with np.errstate(divide="ignore", invalid="ignore"):
            Vn = np.where(M != 0, exchange_sum(numV, halo, peer) / M, 0.0)
            An = np.where(M != 0, exchange_sum(numA, halo, peer) / M, 0.0)
        o.set_nodal(0, M)
        o.newmark_set("Vn", Vn)
        o.newmark_set("An", An)
        active = o.active().astype(bool)[:, None] & np.ones((1, d), bool)
        free = active & (o.fixed() == 0)
        b = np.broadcast_to(np.asarray(P.gravity)[:, k], (nn, d))
        dU = o.newmark_get("dU")                      # initial guess: the Dirichlet increments

        def residual(u):
            st, _ = o.newmark_residual(k, u)          # kinematics + stress of MY particles, gF = -f_int + f_trac
            bad = allsum(st != 0)
            if bad:
                return None
            gF = exchange_sum(o.nodal(2), halo, peer)
            return np.where(free, -gF + M * (a1 * u - a2 * Vn - a3 * An - b), 0.0)

        def norm_owned(v):
            return np.sqrt(allsum(float((v[own] ** 2).sum())))

        def solve(R):
            """Jacobi-PCG on K delta = -R with the slab protocol"""
            o.set_nodal(0, np.zeros_like(M))          # the particle part only: a1 M is added by the owner below
            st, Kl = o.newmark_tangent()
            o.set_nodal(0, M)
            assert st == 0
            notfree = (~free).ravel()
            Kl[notfree, notfree] -= 1.0               # the oracle's unit rows: owner's share, added below
            assert not Kl[notfree].any() and not Kl[:, notfree].any()
            share = np.where(free, a1 * M, 1.0) * own  # what the owner adds to the diagonal

            def band(v):
                return exchange_sum(v, halo, peer)

            def apply(p):
                y = (Kl @ p.ravel()).reshape(nn, d) + share * p
                pAp = allsum(float((p * y).sum()))     # all local rows, before the exchange
                return band(y), pAp

            diag = band(Kl.diagonal().reshape(nn, d) + share)
            # rows nobody touches (outside both slabs' reach) never enter: p = 0 there; keep the division defined
            diag = np.where(diag != 0, diag, 1.0)
            x = np.zeros_like(R)
            r = -R.copy()
            z = r / diag
            p = z.copy()
            rz = allsum(float((r * z)[own].sum()))
            r0 = np.sqrt(allsum(float((r * r)[own].sum())))
            it = 0
            while it < 2000:
                Ap, pAp = apply(p)
                alpha = rz / pAp
                x += alpha * p
                r -= alpha * Ap
                rr = np.sqrt(allsum(float((r * r)[own].sum())))
                it += 1
                if rr <= 1e-13 * r0:
                    break
                z = r / diag
                rz_new = allsum(float((r * z)[own].sum()))
                p = z + (rz_new / rz) * p
                rz = rz_new
            return x, it

        R = residual(dU)
        assert R is not None
        r0n = rn = norm_owned(R)
        iters = 0
        while iters < kw["max_iter"] and not (rn <= 100 * kw["tol"]) and not (rn <= kw["tol"] * r0n):
            delta, _ = solve(R)
            lam, ok = 1.0, False
            for _ in range(8):
                trial = dU + lam * delta
                Rt = residual(trial)
                if Rt is not None and norm_owned(Rt) < rn:
                    ok = True
                    break
                lam *= 0.5
            if not ok:
                trial = dU + delta
                Rt = residual(trial)
                assert Rt is not None
            dU, R = trial, Rt
            rnew = norm_owned(R)
            iters += 1
            if not ok and not (rnew < rn):
                rn = rnew
                break
            rn = rnew
        newton_total += iters
        # both slabs hold the same dU on the band (consistent vectors)
        t = torch.from_numpy(np.ascontiguousarray(dU[halo]))
        u = torch.zeros_like(t)
        for rq in [dist.isend(t, peer), dist.irecv(u, peer)]:
            rq.wait()
        dmax = np.abs(t.numpy() - u.numpy()).max()
        assert dmax <= 1e-12 * max(np.abs(dU).max(), 1e-300), (dmax, np.abs(dU).max(), iters, rn, r0n)
        o.newmark_set("dU", dU)
        assert o.newmark_finish() == 0

    full = oracle.Oracle(P)
    assert full.init_lme() == 0
    full.newmark_setup(**kw)
    for k in range(nsteps):
        assert full.newmark_step(k) == 0
    sc = field_scales(P)
    for name in ("x_GC", "dis", "vel", "acc", "F_n", "Stress", "rho", "J_n", "lambda"):
        assert_close(o.field(name), full.field(name)[rows], f"implicit slab {rank} {name}", rtol=1e-8, scale=sc.get(name))
    assert np.array_equal(o.ints("I0"), full.ints("I0")[rows])
    assert np.array_equal(o.lists(), full.lists()[rows])
    assert newton_total >= nsteps and np.abs(full.field("dis")).max() > 0
    dist.barrier()
    dist.destroy_process_group()
    print(f"slab {rank}: {len(rows)} particles, {len(halo)} halo nodes, {nsteps} implicit steps OK ({newton_total} Newton iterations)")


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    assert world == 2
    case = os.environ.get("SLAB_CASE", "column2d")
    if case == "implicit2d":
        return implicit_main(rank, world)
    if case == "column2d":
        nsteps = 12
        P = synthetic.column_collapse_2d(scale=0.04, nsteps=nsteps)       # 14 x 28 particle cells x 4
        P.fields["vel"][:, 1] = -0.05 * P.solver["cel"]                   # something to do besides gravity
        axis, cuts = engine.slab_cuts(P, world)
        assert axis == 1 and len(cuts) == 1
        tol_balance = 4 * 14 * 2
    else:
        # BASELINE configs[3] in small: the 3D Matsuoka-Nakai slope under its gravity ramp, cut along the slope (z) at
        # the particle-count median -- the geometry of bench.py --workload c4 on the global mesh
        nsteps = 6
        P, _ = synthetic.slope_slab_3d(0, 1, cells=30, width=2, nsteps=nsteps, ramp_steps=3)
        axis, cuts = engine.slab_cuts(P, world, axis=2)
        assert axis == 2 and len(cuts) == 1
        tol_balance = 8 * 2 * 30
    owner = engine.slab_owner(P, axis, cuts)
    counts = np.bincount(owner, minlength=world)
    assert counts.sum() == P.np_ and abs(int(counts[0]) - int(counts[1])) <= tol_balance, counts
    rows = np.nonzero(owner == rank)[0]
    halo = engine.slab_halo_nodes(P, axis, cuts[0], 6)
    # both ranks planned the same thing
    chk = torch.tensor([float(cuts[0]), float(len(halo)), float(halo.sum())], dtype=torch.float64)
    ref = chk.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(chk, ref)

    o = oracle.Oracle(subset(P, rows))
    peer = 1 - rank
    d = P.ndim

    def merge_active():
        a = o.active()
        t = torch.from_numpy(a[halo].copy())
        u = torch.zeros_like(t)
        for r in [dist.isend(t, peer), dist.irecv(u, peer)]:
            r.wait()
        a[halo] |= u.numpy()
        o.set_active(a)

    assert o.search_closest() == 0
    merge_active()
    assert o.search_lists() == 0
    for k in range(nsteps):
        assert o.search_closest() == 0
        merge_active()
        assert o.search_lists() == 0
        assert o.stage("p2g_mass_disp", k) == 0
        M, mom = o.nodal(0), o.nodal(1)
        # coverage of the band: outside it only one slab contributes
        mine = torch.from_numpy((M[:, 0] != 0).astype(np.uint8))
        both = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(both, mine)
        shared = np.nonzero(both[0].numpy() & both[1].numpy())[0]
        assert np.isin(shared, halo).all(), "a node outside the halo band receives sums from both slabs"
        o.set_nodal(0, exchange_sum(M, halo, peer))
        o.set_nodal(1, exchange_sum(mom, halo, peer))
        assert o.stage("grid_disp", k) == 0
        assert o.stage("kin_stress", k) == 0
        assert o.stage("force", k) == 0
        o.set_nodal(2, exchange_sum(o.nodal(2), halo, peer))
        assert o.stage("grid_acc", k) == 0
        assert o.stage("g2p", k) == 0

    # single-domain oracle on rank 0's side of the comparison: every rank checks its own rows
    full = oracle.Oracle(P)
    assert full.init_lme() == 0
    for k in range(nsteps):
        assert full.step(k) == 0
    sc = field_scales(P)
    for name in ("x_GC", "vel", "acc", "F_n", "Stress", "rho", "J_n", "lambda", "Beta", "EPS_n", "b_e_n"):
        assert_close(o.field(name), full.field(name)[rows], f"slab {rank} {name}", scale=sc.get(name))
    assert np.array_equal(o.ints("I0"), full.ints("I0")[rows])
    assert np.array_equal(o.lists(), full.lists()[rows])
    # plastic flow happened (2D column) / the slope is loaded (3D), so the comparison is not trivial
    if case == "column2d":
        assert float((full.field("EPS_n") > 0).sum()) > 0
    else:
        assert np.abs(full.field("Stress")).max() > 0
    dist.barrier()
    dist.destroy_process_group()
    print(f"slab {rank}: {len(rows)} particles, {len(halo)} halo nodes, {nsteps} steps OK")


if __name__ == "__main__":
    main()
