"""Synthetic particle clouds of the shapes BASELINE.json names, as plain arrays (no files).

Structured Q4 (2D) / H8 (3D) background grids with the node / element numbering of
tests/deckgen.py, particles seeded at the Gauss points of a grid-aligned particle mesh
(element_to_particles__Q4__, Nodes/Q4.c:342-450; __H8__, Nodes/H8.c:389-587), volumes
Vol_element / GPxElement (Generate-One-Phase-Analysis.c:572-629).  Adjacency comes from the
engine's own scalable builder (nlps_b200_build_locality).
"""
from __future__ import annotations

import numpy as np

from . import engine
from .problem import Problem

_G = 1.0 / np.sqrt(3.0)


def _grid(nx, ny, nz, h, origin):
    if nz is None:
        ii, jj = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), indexing="xy")
        coords = np.stack([origin[0] + ii.ravel() * h, origin[1] + jj.ravel() * h], axis=1)
        e = np.arange(nx * ny)
        i, j = e % nx, e // nx
        n1 = j * (nx + 1) + i
        conn = np.stack([n1, n1 + 1, n1 + 1 + (nx + 1), n1 + (nx + 1)], axis=1).astype(np.int32)
    else:
        kk, jj, ii = np.meshgrid(np.arange(nz + 1), np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
        coords = np.stack([origin[0] + ii.ravel() * h, origin[1] + jj.ravel() * h, origin[2] + kk.ravel() * h], axis=1)
        e = np.arange(nx * ny * nz)
        i, j, k = e % nx, (e // nx) % ny, e // (nx * ny)
        sx, sy = 1, nx + 1
        sz = (nx + 1) * (ny + 1)
        n1 = k * sz + j * sy + i
        conn = np.stack([n1, n1 + sx, n1 + sx + sy, n1 + sy, n1 + sz, n1 + sx + sz, n1 + sx + sy + sz,
                         n1 + sy + sz], axis=1).astype(np.int32)
    return np.ascontiguousarray(coords), conn


def structured_problem(ndim, grid_cells, h, block_cells, block_origin_cell, material, nsteps, cfl, cel,
                       gravity, gamma_lme=3.0, fixed=("bottom",), rollers=("left", "right"), jitter=0.0,
                       seed=20261018, tol_radial=None, maxiter_radial=None, bc_scale=None):
    """Block of block_cells particle cells (one particle cell = one background cell, GPxElement 4 / 8)
    placed at block_origin_cell inside a grid of grid_cells cells of size h."""
    d = ndim
    nx, ny = grid_cells[0], grid_cells[1]
    nz = grid_cells[2] if d == 3 else None
    coords, conn = _grid(nx, ny, nz, h, (0.0,) * d)
    r1p, r1i, r2p, r2i, h_avg, dx = engine.build_locality(d, coords, conn)
    mtype, mpar = material
    if tol_radial is None:  # globals set by the last material parsed (F10-iv)
        tol_radial, maxiter_radial = {"Drucker-Prager": (1e-14, 10), "Matsuoka-Nakai": (1e-10, 20)}.get(
            mtype, (1e-14, 10))
    solver = dict(cfl=cfl, cel=cel, nsteps=nsteps, gamma_lme=gamma_lme, tol_zero=1e-6, tol_wrapper=1e-10,
                  max_iter_lme=10, tol_radial=tol_radial, maxiter_radial=maxiter_radial, thickness=1.0)
    g = np.zeros((d, nsteps))
    for k in range(d):
        g[k, :] = gravity[k]
    P = Problem(ndim=d, coords=coords, r1p=r1p, r1i=r1i, r2p=r2p, r2i=r2i, h_avg=h_avg, dx=dx, solver=solver,
                gravity=g)
    P.conn = conn
    P.materials = [(mtype, np.asarray(mpar, dtype=np.float64))]
    # ---- Dirichlet sets (zero curves), boundary order = deck order
    nxn, nyn = nx + 1, ny + 1
    nzn = (nz + 1) if d == 3 else 1
    idx = np.arange(nxn * nyn * nzn)
    ci, cj, ck = idx % nxn, (idx // nxn) % nyn, idx // (nxn * nyn)
    vert = cj if d == 2 else ck
    sel = dict(bottom=vert == 0, top=vert == (ny if d == 2 else nz), left=ci == 0, right=ci == nx)
    if d == 3:
        sel.update(front=cj == 0, back=cj == ny)
    zeros = np.zeros((d, nsteps))
    for name in fixed:
        P.bounds.append(dict(nodes=idx[sel[name]].astype(np.int32), dir=np.ones((d, nsteps), np.int32), val=zeros.copy()))
    for name in rollers:
        dr = np.zeros((d, nsteps), np.int32)
        dr[1 if name in ("front", "back") else 0, :] = 1
        P.bounds.append(dict(nodes=idx[sel[name]].astype(np.int32), dir=dr, val=zeros.copy()))
    # ---- particles
    bx, by = block_cells[0], block_cells[1]
    bz = block_cells[2] if d == 3 else 1
    ox, oy = block_origin_cell[0], block_origin_cell[1]
    oz = block_origin_cell[2] if d == 3 else 0
    e = np.arange(bx * by * bz)
    ei, ej, ek = e % bx + ox, (e // bx) % by + oy, e // (bx * by) + oz
    if d == 2:
        xi = np.array([[_G, _G], [_G, -_G], [-_G, _G], [-_G, -_G]])          # Q4.c:358-366
    else:
        xi = np.array([[sx_, sy_, sz_] for sz_ in (_G, -_G) for sy_ in (_G, -_G) for sx_ in (_G, -_G)])
    gp = xi.shape[0]
    cen = np.stack([(ei + 0.5) * h, (ej + 0.5) * h] + ([(ek + 0.5) * h] if d == 3 else []), axis=1)
    x = (cen[:, None, :] + 0.5 * h * xi[None, :, :]).reshape(-1, d)
    if jitter > 0.0:
        x = x + np.random.default_rng(seed).uniform(-jitter * h, jitter * h, x.shape)
    corner = (xi > 0).astype(np.int64)                                          # nearest corner of the cell
    ni = (ei[:, None] + corner[None, :, 0]).ravel()
    nj = (ej[:, None] + corner[None, :, 1]).ravel()
    nk = (ek[:, None] + corner[None, :, 2]).ravel() if d == 3 else 0
    I0 = (nk * (nxn * nyn) + nj * nxn + ni).astype(np.int32)
    if jitter > 0.0:
        # closest node of the containing cell
        c = np.floor(x / h).astype(np.int64)
        f = x / h - c
        rn = c + (f > 0.5)
        I0 = ((rn[:, 2] if d == 3 else 0) * (nxn * nyn) + rn[:, 1] * nxn + rn[:, 0]).astype(np.int32)
    vol = np.full(x.shape[0], h ** d / gp)
    P.init_fields(x, vol, np.zeros(x.shape[0], np.int32))
    P.I0 = I0
    return P


DP_C2 = ("Drucker-Prager", [2000.0, 1e7, 0.3, 0.0, 1e4, 1.0, 1e-2, 30.0, 0.0, 1.0, 0, 0, 0, 0, 0, 0])
NH_C1 = ("Neo-Hookean-Wriggers", [1000.0, 1e6, 0.3] + [0.0] * 13)
MN_C4 = ("Matsuoka-Nakai", [2000.0, 1e7, 0.3, 0.0, 8.0 / 3.0, 0.0, 0.0, 30.0, 0.0, 0.0, 1e3, 0.5, 20000.0, 0.005, 35.0, 0.0])


def column_collapse_2d(scale=1.0, nsteps=1000):
    """BASELINE configs[1]: 2D granular column (aspect 2, 0.2 m x 0.4 m) of Drucker-Prager material
    released in a box 6 column-widths wide with a fixed base and frictionless sides.  scale=1 ->
    354 x 708 particle cells x GPxElement 4 = 1,002,528 particles (the 10^6 of BASELINE.json; SURVEY
    8(d) C2) on a 2124 x 885 cell Q4 background grid."""
    bx = max(4, int(round(354 * scale)))
    by = 2 * bx
    nx, ny = 6 * bx, by + by // 4
    h = 0.2 / bx
    return structured_problem(2, (nx, ny), h, (bx, by), (0, 0), DP_C2, nsteps, 0.5, (1e7 / 2000.0) ** 0.5 * 1.3,
                              (0.0, -9.81))


def block_2d(cells=16, nsteps=200):
    """BASELINE configs[0]: elastic block under gravity."""
    return structured_problem(2, (cells + 4, cells + 4), 1.0 / cells, (cells, cells), (2, 0), NH_C1, nsteps, 0.5,
                              (1e6 / 1000.0) ** 0.5 * 1.3, (0.0, -9.81))


def cube_3d(cells=24, nsteps=100, material=NH_C1, gamma_lme=6.0):
    """BASELINE configs[2] shape: 3D cube on an H8 grid, GPxElement 8, gamma 6."""
    cel = (material[1][1] / material[1][0]) ** 0.5 * 1.3
    return structured_problem(3, (cells + 4, cells + 4, cells + 4), 1.0 / cells, (cells, cells, cells), (2, 2, 0),
                              material, nsteps, 0.5, cel, (0.0, 0.0, -9.81), gamma_lme=gamma_lme,
                              rollers=("left", "right", "front", "back"))
