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
