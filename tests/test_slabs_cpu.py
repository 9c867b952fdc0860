"""Host side of the multi-GPU slab path, no GPU needed: planning functions of the C ABI
and a world_size-2 gloo run of the exchange protocol with the oracle as the per-slab stepper."""
import os
import subprocess
import sys

import numpy as np
import pytest

from nlps_b200 import engine, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 3, 4])
def test_cuts_balance_and_ownership(world):
    P = synthetic.column_collapse_2d(scale=0.12, nsteps=2)  # 42 x 84 particle cells
    axis, cuts = engine.slab_cuts(P, world)
    assert axis == 1                                           # the column is taller than wide
    assert len(cuts) == world - 1 and np.all(np.diff(cuts) > 0)
    layers = np.unique(P.coords[:, axis])
    for c in cuts:                                             # strictly between two node layers
        assert not np.isclose(layers, c, atol=1e-9 * P.dx).any()
    owner = engine.slab_owner(P, axis, cuts)
    counts = np.bincount(owner, minlength=world)
    per_layer = 2 * 42 * 2                                     # particles whose I0 sits on one node layer
    assert counts.sum() == P.np_ and counts.max() - counts.min() <= 2 * per_layer
    L = engine.lib()
    m, keep = engine._mesh_struct(P)
    import ctypes as C
    cu = np.ascontiguousarray(cuts)
    for p in range(0, P.np_, 997):                             # the scalar C twin agrees
        assert L.nlps_b200_slab_owner(C.byref(m), axis, world, cu.ctypes.data_as(C.POINTER(C.c_double)),
                                      int(P.I0[p])) == owner[p]


def test_halo_nodes_cover_the_two_ring_reach():
    P = synthetic.cube_3d(cells=10, nsteps=2)
    axis, cuts = engine.slab_cuts(P, 2, axis=0)
    band = 6
    halo = engine.slab_halo_nodes(P, axis, cuts[0], band)
    assert np.all(np.diff(halo) > 0)
    assert np.all(np.abs(P.coords[halo, axis] - cuts[0]) <= band * P.dx * (1 + 1e-9))
    # every node a particle can reach (2-ring of I0), for particles up to band-3.5 cells beyond the cut,
    # lies in the band when it is on the far side of the cut
    owner = engine.slab_owner(P, axis, cuts)
    inhalo = np.zeros(P.nn, bool)
    inhalo[halo] = True
    for p in np.nonzero(owner == 0)[0][::17]:
        ring = P.r2i[P.r2p[P.I0[p]]:P.r2p[P.I0[p] + 1]]
        far = ring[P.coords[ring, axis] > cuts[0]]
        assert inhalo[far].all()


def test_more_slabs_than_layers_is_refused():
    P = synthetic.block_2d(cells=4, nsteps=2)
    with pytest.raises(RuntimeError):
        engine.slab_cuts(P, 64)


@pytest.mark.parametrize("case,port", [("column2d", "29731"), ("slope3d", "29733"), ("implicit2d", "29735")])
def test_slab_protocol_world2_gloo(case, port):
    """Two gloo ranks, each stepping its slab with the oracle and exchanging exactly what the engine
    exchanges; particle fields match the single-domain oracle to 1e-10, lists bit-exact.  column2d: the
    Drucker-Prager column of the bench line; slope3d: the Matsuoka-Nakai slope of bench.py --workload c4; implicit2d:
    the Newmark-beta scheme over two slabs (per-slab tangents, band sums of the Krylov vector, owner's share of the mass
    term, dot products over owned rows) against the single-domain oracle's dense-LU Newton, 1e-8."""
    env = dict(os.environ, OMP_NUM_THREADS="2", SLAB_CASE=case)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", port, os.path.join(ROOT, "tests", "workers", "slab_gloo_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("steps OK") == 2
