"""Export the state held by the reference harness (oracle/_ref) as a `Problem`.  Test infra."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "nl-partsol_b200"))
from nlps_b200.problem import ALL_FIELDS, Problem  # noqa: E402


def lists_dense(ptr, idx, cap):
    n = len(ptr) - 1
    out = np.full((n, cap), -1, np.int32)
    for p in range(n):
        k = ptr[p + 1] - ptr[p]
        out[p, :k] = idx[ptr[p]:ptr[p + 1]]
    return out


def problem_from_ref(h, conn=None) -> Problem:
    s = h.scalars()
    r1p, r1i = h.table(2)
    r2p, r2i = h.table(3)
    solver = dict(cfl=s["cfl"], cel=s["cel"], nsteps=h.nsteps, gamma_lme=s["gamma_lme"],
                  tol_zero=s["tol_zero"], tol_wrapper=s["tol_wrapper"], max_iter_lme=s["max_iter_lme"],
                  tol_radial=s["tol_radial"], maxiter_radial=s["maxiter_radial"],
                  thickness=s["thickness"])
    p = Problem(ndim=h.ndim, coords=h.coords(), r1p=r1p, r1i=r1i, r2p=r2p, r2i=r2i, h_avg=h.h_avg(),
                dx=s["delta_x"], solver=solver, gravity=h.gravity())
    p.bounds = h.bounds()
    p.neumann = h.neumann() if hasattr(h.lib, "refh_neumann_nodes") else []
    p.materials = [h.material(m) for m in range(h.lib.refh_num_materials())]
    p.fields = {k: h.field(k) for k in ALL_FIELDS}
    try:  # GramsShapeFun (Type=aLME): Beta is the d x d metric (exported above), plus the cut-off ellipsoid
        p.fields["Cut_off_Ellipsoid"] = h.field("Cut_off_Ellipsoid")
        solver["alme"] = 1
    except KeyError:
        pass
    if any(t == "Von-Mises" for t, _ in p.materials):
        p.fields["Back_stress"] = h.field("Back_stress")
    p.I0 = h.ints("I0")
    p.MatIdx = h.ints("MatIdx")
    if conn is None:
        cp, ci = h.table(0)
        nne = cp[1] - cp[0]
        conn = ci.reshape(-1, nne)[:, ::-1].copy()  # chains are the file order reversed
    p.conn = conn
    return p
