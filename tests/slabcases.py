"""Problems and helpers shared by the multi-slab tests (single-GPU thread loopback and NCCL)."""
from __future__ import annotations

import numpy as np

from nlps_b200 import synthetic

COMPARE = ("x_GC", "dis", "D_dis", "vel", "acc", "F_n", "DF", "Stress", "rho", "J_n", "W", "b_e_n", "EPS_n",
           "Kappa_n", "lambda", "Beta")


def moving_block(width=48, nsteps=60):
    """Neo-Hookean block flying along x with a sheared velocity profile under gravity: layers of
    particles keep crossing the cuts, so migration is exercised every few steps."""
    h = 1.0 / 16
    P = synthetic.structured_problem(2, (2 * width, 24), h, (width, 12), (6, 6), synthetic.NH_C1, nsteps, 0.5,
                                     (1e6 / 1000.0) ** 0.5 * 1.3, (0.0, -9.81), fixed=("bottom",), rollers=())
    y = P.fields["x_GC"][:, 1]
    cel = P.solver["cel"]
    P.fields["vel"][:, 0] = 0.3 * cel * (1.0 + 0.3 * np.sin(2 * np.pi * (y - y.min()) / (12 * h)))
    return P


def sinking_column(nsteps=60):
    """Drucker-Prager column pushed onto its fixed base: plastic history (b_e, EPS, kappa) travels
    with the migrating particles; the slab axis is y."""
    P = synthetic.column_collapse_2d(scale=0.1, nsteps=nsteps)
    P.fields["vel"][:, 1] = -0.12 * P.solver["cel"]
    return P


def merge(world_results, n_global):
    """[(fields, ids, counts, lists)] per slab -> global arrays by particle id"""
    out = {}
    seen = np.zeros(n_global, np.int32)
    for fields, ids, counts, lists in world_results:
        seen[ids] += 1
        for k, v in fields.items():
            if k not in out:
                out[k] = np.zeros((n_global,) + v.shape[1:], v.dtype)
            out[k][ids] = v
        out.setdefault("_counts", np.zeros(n_global, np.int32))[ids] = counts[ids]
        out.setdefault("_lists", np.full((n_global, lists.shape[1]), -7, np.int32))[ids] = lists[ids]
    assert np.all(seen == 1), "every particle must live in exactly one slab"
    return out
