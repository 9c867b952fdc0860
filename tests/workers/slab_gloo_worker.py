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


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    assert world == 2
    case = os.environ.get("SLAB_CASE", "column2d")
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
