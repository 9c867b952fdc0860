"""Synthetic generators (bench / GPU-box inputs): sanity against the oracle port.  CPU only."""
import numpy as np
import pytest

import oracle
from nlps_b200 import engine, synthetic


def _brute_closest(P):
    d2 = ((P.fields["x_GC"][:, None, :] - P.coords[None, :, :]) ** 2).sum(-1)
    return d2.argmin(1).astype(np.int32)


@pytest.mark.parametrize("make", [lambda: synthetic.column_collapse_2d(scale=0.02, nsteps=10),
                                  lambda: synthetic.block_2d(cells=8, nsteps=10),
                                  lambda: synthetic.cube_3d(cells=4, nsteps=6)])
def test_synthetic_problem_runs_on_oracle(make):
    P = make()
    assert np.array_equal(P.I0, _brute_closest(P))
    # product locality builder == oracle restatement of the reference construction (bit-exact)
    r = oracle.build_locality(P.ndim, P.coords, P.conn)
    for a, b in zip(r[:5], (P.r1p, P.r1i, P.r2p, P.r2i, P.h_avg)):
        assert np.array_equal(a, b)
    assert r[5] == P.dx
    o = oracle.Oracle(P)
    assert o.init_lme() == 0, o.error()
    N, dN = o.shape(P.np_ // 2)   # converged LME weights: partition of unity, first-order consistency
    assert abs(N.sum() - 1) < 1e-13 and np.abs(dN.sum(0)).max() < 1e-6
    m0 = P.fields["mass"].sum()
    for k in range(3):
        assert o.step(k) == 0, o.error()
    # partition of unity => the lumped mass sums to the particle mass
    assert abs(o.nodal(0)[:, 0].sum() - m0) <= 1e-12 * m0


def test_abi_exports_every_declared_symbol():
    import re, os
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "nlps_b200.h")).read()
    names = set(re.findall(r"\b(nlps_b200_\w+)\s*\(", hdr))
    L = engine.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert len(names) >= 20


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        engine.Engine(synthetic.block_2d(cells=4, nsteps=2))


def test_slope_slabs_partition_the_global_slope():
    """bench.py --workload c4 host logic (no GPU): the per-rank slope problems of synthetic.slope_slab_3d tile the
    single-domain slope -- every particle of the global problem is owned by exactly one slab (by its closest node and the
    cuts), global ids match positions bit for bit, the quantile cuts balance the particle counts."""
    import numpy as np
    from nlps_b200 import synthetic
    cells, width, world = 48, 4, 4
    G, none = synthetic.slope_slab_3d(0, 1, cells=cells, width=width, nsteps=2, band_cells=5)
    assert none is None
    assert G.np_ == 8 * width * cells * (cells + 1) // 2
    gid_G = (G.kept_cells.astype(np.int64)[:, None] * 8 + np.arange(8)[None, :]).ravel()
    xg = G.fields["x_GC"]
    owned = np.zeros(G.np_, np.int32)
    counts = []
    for r in range(world):
        P, sl = synthetic.slope_slab_3d(r, world, cells=cells, width=width, nsteps=2, band_cells=5)
        assert sl["n_particles"] == G.np_ and sl["axis"] == 2 and len(sl["cuts"]) == world - 1
        rows = np.searchsorted(gid_G, sl["global_id"])
        assert np.array_equal(gid_G[rows], sl["global_id"])
        assert np.array_equal(P.fields["x_GC"], xg[rows])              # same bits as the global problem
        assert np.array_equal(P.coords[P.I0], G.coords[G.I0[rows]])    # same closest node, in sub-mesh numbering
        z = P.coords[P.I0][:, 2]                                       # the owner is decided by the closest node layer
        lo = sl["cuts"][r - 1] if r > 0 else -np.inf
        hi = sl["cuts"][r] if r < world - 1 else np.inf
        mine = (z > lo) & (z < hi)
        owned[rows[mine]] += 1
        counts.append(int(mine.sum()))
        assert (P.I0 + sl["node_offset"] == G.I0[rows]).all()          # z sub-mesh: node ids differ by a constant
    assert np.all(owned == 1)
    # quantile cuts: balanced wherever the slabs may be thinner than the clamp (two halo bands + 2 layers)
    c2 = []
    for r in range(2):
        P, sl = synthetic.slope_slab_3d(r, 2, cells=cells, width=width, nsteps=2, band_cells=5)
        z = P.coords[P.I0][:, 2]
        c2.append(int(((z > sl["cuts"][0]) if r else (z < sl["cuts"][0])).sum()))
    assert sum(c2) == G.np_ and max(c2) <= 1.1 * G.np_ / 2, c2
