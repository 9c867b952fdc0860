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


def _grid(nx, ny, nz, h, origin, cell_offset=(0, 0, 0)):
    """cell_offset: index of the grid's first node in a larger (global) grid; coordinates are
    (index + offset) * h so that a sub-grid reproduces the global coordinates bit for bit."""
    if nz is None:
        ii, jj = np.meshgrid(np.arange(nx + 1) + cell_offset[0], np.arange(ny + 1) + cell_offset[1], indexing="xy")
        coords = np.stack([origin[0] + ii.ravel() * h, origin[1] + jj.ravel() * h], axis=1)
        e = np.arange(nx * ny)
        i, j = e % nx, e // nx
        n1 = j * (nx + 1) + i
        conn = np.stack([n1, n1 + 1, n1 + 1 + (nx + 1), n1 + (nx + 1)], axis=1).astype(np.int32)
    else:
        kk, jj, ii = np.meshgrid(np.arange(nz + 1) + cell_offset[2], np.arange(ny + 1) + cell_offset[1],
                                 np.arange(nx + 1) + cell_offset[0], indexing="ij")
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
                       seed=20261018, tol_radial=None, maxiter_radial=None, bc_scale=None, cell_offset=(0, 0, 0),
                       keep_cell=None):
    """Block of block_cells particle cells (one particle cell = one background cell, GPxElement 4 / 8)
    placed at block_origin_cell inside a grid of grid_cells cells of size h.  cell_offset places the
    grid inside a larger global one (sub-mesh of one slab); block_origin_cell stays LOCAL."""
    d = ndim
    nx, ny = grid_cells[0], grid_cells[1]
    nz = grid_cells[2] if d == 3 else None
    coords, conn = _grid(nx, ny, nz, h, (0.0,) * d, cell_offset)
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
        if isinstance(name, tuple):     # (face, constrained direction)
            name, axis_ = name
            dr[axis_, :] = 1
        else:
            dr[1 if name in ("front", "back") else 0, :] = 1
        P.bounds.append(dict(nodes=idx[sel[name]].astype(np.int32), dir=dr, val=zeros.copy()))
    # ---- particles
    bx, by = block_cells[0], block_cells[1]
    bz = block_cells[2] if d == 3 else 1
    ox, oy = block_origin_cell[0], block_origin_cell[1]
    oz = block_origin_cell[2] if d == 3 else 0
    e = np.arange(bx * by * bz)
    ei, ej, ek = e % bx + ox, (e // bx) % by + oy, e // (bx * by) + oz
    if keep_cell is not None:   # carve the block: keep_cell(global cell indices i, j, k) -> bool mask
        m_ = keep_cell(ei + cell_offset[0], ej + cell_offset[1], ek + (cell_offset[2] if d == 3 else 0))
        ei, ej, ek = ei[m_], ej[m_], ek[m_]
        P.kept_cells = np.nonzero(m_)[0]
    if d == 2:
        # Gauss points (xi, eta) of Q4.c:358-366 in order; the element chain is the file order REVERSED
        # (Read-GID-Mesh.c:406-408), which maps eta to -y: physical offsets (+,-), (+,+), (-,-), (-,+)
        xi = np.array([[_G, -_G], [_G, _G], [-_G, -_G], [-_G, _G]])
    else:
        xi = np.array([[sx_, sy_, sz_] for sz_ in (_G, -_G) for sy_ in (_G, -_G) for sx_ in (_G, -_G)])
    gp = xi.shape[0]
    co = cell_offset
    cen = np.stack([(ei + co[0] + 0.5) * h, (ej + co[1] + 0.5) * h] + ([(ek + co[2] + 0.5) * h] if d == 3 else []), axis=1)
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
        rn = c + (f > 0.5) - np.asarray(co[:d])
        I0 = ((rn[:, 2] if d == 3 else 0) * (nxn * nyn) + rn[:, 1] * nxn + rn[:, 0]).astype(np.int32)
    vol = np.full(x.shape[0], h ** d / gp)
    P.init_fields(x, vol, np.zeros(x.shape[0], np.int32))
    P.I0 = I0
    return P


DP_C2 = ("Drucker-Prager", [2000.0, 1e7, 0.3, 0.0, 1e4, 1.0, 1e-2, 30.0, 0.0, 1.0, 0, 0, 0, 0, 0, 0])
NH_C1 = ("Neo-Hookean-Wriggers", [1000.0, 1e6, 0.3] + [0.0] * 13)
# SURVEY 8(f)-4 laws.  Von-Mises: slots 4 / 5 = Yield-stress / Hardening-Modulus, 16..19 = theta, K-0, K-inf, delta (Voce)
VM_SOFT = ("Von-Mises", [2000.0, 1e6, 0.3, 0.0, 400.0, 2e4] + [0.0] * 10 + [0.6, 20.0, 150.0, 40.0])
HENCKY_C1 = ("Hencky", [1000.0, 1e6, 0.3] + [0.0] * 13)
MN_C4 = ("Matsuoka-Nakai", [2000.0, 1e7, 0.3, 0.0, 8.0 / 3.0, 0.0, 0.0, 30.0, 0.0, 0.0, 1e3, 0.5, 20000.0, 0.005, 35.0, 0.0])


def column_collapse_2d(scale=1.0, nsteps=1000, material=None):
    """BASELINE configs[1]: 2D granular column (aspect 2, 0.2 m x 0.4 m) of Drucker-Prager material
    released in a box 6 column-widths wide with a fixed base and frictionless sides.  scale=1 ->
    354 x 708 particle cells x GPxElement 4 = 1,002,528 particles (the 10^6 of BASELINE.json; SURVEY
    8(d) C2) on a 2124 x 885 cell Q4 background grid."""
    bx = max(4, int(round(354 * scale)))
    by = 2 * bx
    nx, ny = 6 * bx, by + by // 4
    h = 0.2 / bx
    return structured_problem(2, (nx, ny), h, (bx, by), (0, 0), material or DP_C2, nsteps, 0.5, (1e7 / 2000.0) ** 0.5 * 1.3,
                              (0.0, -9.81))


def block_2d(cells=16, nsteps=200, material=NH_C1):
    """BASELINE configs[0]: elastic block under gravity."""
    cel = (material[1][1] / material[1][0]) ** 0.5 * 1.3
    return structured_problem(2, (cells + 4, cells + 4), 1.0 / cells, (cells, cells), (2, 0), material, nsteps, 0.5,
                              cel, (0.0, -9.81))


def _top_platen(P, top_layer, nx, ny, nsteps, per_step, layer_offset=0):
    """Dirichlet set of the cube compression: the node plane on top of the particle block moves down by
    `per_step` every step (the curve sample k is the nodal displacement INCREMENT of step k, U-Verlet.c:515-521)."""
    k = top_layer - layer_offset
    nplane = (nx + 1) * (ny + 1)
    if k < 0 or (k + 1) * nplane > P.nn:
        return
    nodes = (k * nplane + np.arange(nplane)).astype(np.int32)
    dr = np.zeros((3, nsteps), np.int32)
    dr[2, :] = 1
    val = np.zeros((3, nsteps))
    val[2, :] = -per_step
    P.bounds.append(dict(nodes=nodes, dir=dr, val=val))


def cube_3d(cells=24, nsteps=100, material=NH_C1, gamma_lme=6.0, xy=None, compress=True):
    """BASELINE configs[2]: 3D Neo-Hookean cube on an H8 grid, GPxElement 8, LME gamma 6, bottom fixed,
    lateral rollers, the top node plane pushed down by 10 % of the height in 500 steps (compress=True; no gravity), or
    the cube under gravity (compress=False).  xy: cross-section in cells when it is not a cube (cells = height)."""
    cel = (material[1][1] / material[1][0]) ** 0.5 * 1.3
    cxy = cells if xy is None else xy
    P = structured_problem(3, (cxy + 4, cxy + 4, cells + 4), 1.0 / cells, (cxy, cxy, cells), (2, 2, 0),
                           material, nsteps, 0.5, cel, (0.0, 0.0, 0.0 if compress else -9.81), gamma_lme=gamma_lme,
                           rollers=("left", "right", "front", "back"))
    if compress:
        _top_platen(P, cells, cxy + 4, cxy + 4, nsteps, 0.1 / 500.0)
    return P


def column_slab_2d(rank, world, scale=1.0, nsteps=1000, band_cells=6):
    """Weak-scaling version of column_collapse_2d for `world` slabs stacked along y: the column is
    `world` times taller (2 * bx * world particle-cell rows), every slab gets the rows
    [rank * by, (rank + 1) * by) plus one extra row either side (the engine keeps what the slab owns),
    on a SUB-MESH that reaches band_cells + 2 cells beyond its cuts -- per-rank memory and setup stay
    constant.  Node coordinates are (global index) * h: identical bits on every rank.
    Returns (Problem, slab dict for engine.Engine(..., slab=...))."""
    bx = max(4, int(round(354 * scale)))
    by = 2 * bx
    nx = 6 * bx
    h = 0.2 / bx
    tot_rows = by * world
    ny_glob = tot_rows + by // 4
    pad = band_cells + 2
    j0 = max(0, rank * by - pad)                        # first cell row of the sub-mesh
    j1 = min(ny_glob, (rank + 1) * by + pad) if rank < world - 1 else ny_glob
    c0 = max(0, rank * by - 1) if rank > 0 else 0      # particle-cell rows generated here
    c1 = min(tot_rows, (rank + 1) * by + 1) if rank < world - 1 else tot_rows
    P = structured_problem(2, (nx, j1 - j0), h, (bx, c1 - c0), (0, c0 - j0), DP_C2, nsteps, 0.5,
                           (1e7 / 2000.0) ** 0.5 * 1.3, (0.0, -9.81), fixed=("bottom",) if j0 == 0 else (),
                           cell_offset=(0, j0, 0))
    # global particle ids: (global cell index) * 4 + Gauss point; cells are numbered row-major
    e = np.arange(bx * (c1 - c0))
    gcell = (e // bx + c0) * bx + e % bx
    gid = (gcell[:, None] * 4 + np.arange(4)[None, :]).ravel().astype(np.int32)
    # a particle in cell row j has its closest node on row j or j + 1: slab r owns node rows (r*by, (r+1)*by]
    cuts = np.array([(r * by + 0.5) * h for r in range(1, world)])
    slab = dict(rank=rank, world=world, axis=1, cuts=cuts, band_cells=band_cells, global_id=gid,
                n_global=4 * bx * tot_rows, node_offset=j0 * (nx + 1))
    return P, slab


def slope_slab_3d(rank, world, cells=160, width=78, nsteps=100, gamma_lme=6.0, band_cells=6, material=MN_C4,
                  ramp_steps=100):
    """BASELINE configs[3]: a 45-degree slope of Matsuoka-Nakai soil, GPxElement 8, gravity ramped over
    `ramp_steps` steps, fixed base, rollers on the four sides, LME gamma 6; cells=160, width=78: 1,004,640 particle cells
    = 8,037,120 particles.  Axes: the VERTICAL is the grid's x axis (ground plane i = 0, gravity (-g, 0, 0)), the slope
    runs along z, the slab axis (node ids of a z sub-mesh are the global ids minus a constant, which is what the
    migration of closest-node ids needs): the particle cells (i, j, k) with i + k < cells of a cells x width x cells block.
    Cuts at particle-count quantiles (the wedge is taller at small z, so the slabs there are thinner); every rank builds
    only the sub-mesh within band_cells + 2 layers of its cuts and the particles of its layers (+1 layer each side; the
    engine keeps what the slab owns).  world == 1: (Problem, None).  Strong scaling of a fixed global problem."""
    c, w = cells, width
    h = 1.0 / c
    col = np.arange(c, 0, -1, dtype=np.int64)                 # particle cells in the z layer k: (c - k) * w
    cum = np.concatenate([[0], np.cumsum(col)])
    lay = [0] + [int(np.searchsorted(cum, cum[-1] * r / world)) for r in range(1, world)] + [c]
    for r in range(1, world):                                 # slabs must be wider than two halo bands
        lay[r] = max(lay[r], lay[r - 1] + 2 * band_cells + 2)
    nx, ny, nz_glob = c + 4, w + 4, c + 4
    pad = band_cells + 2
    k0 = max(0, lay[rank] + 2 - pad) if rank > 0 else 0                     # sub-mesh cell layers (global indices)
    k1 = min(nz_glob, lay[rank + 1] + 2 + pad) if rank < world - 1 else nz_glob
    c0 = max(0, lay[rank] - 1) if rank > 0 else 0                           # particle-cell layers generated here
    c1 = min(c, lay[rank + 1] + 1) if rank < world - 1 else c
    cel = (material[1][1] / material[1][0]) ** 0.5 * 1.3
    # the particle block starts at grid cell (0, 2, 2): global particle-cell layer of grid layer gk is gk - 2
    keep = lambda gi, gj, gk: gi + (gk - 2) < c
    rollers = [("front", 1), ("back", 1)]
    if k0 == 0:
        rollers.append(("bottom", 2))
    if k1 == nz_glob:
        rollers.append(("top", 2))
    P = structured_problem(3, (nx, ny, k1 - k0), h, (c, w, c1 - c0), (0, 2, c0 + 2 - k0), material, nsteps, 0.5, cel,
                           (-9.81, 0.0, 0.0), gamma_lme=gamma_lme, fixed=("left",), rollers=tuple(rollers),
                           cell_offset=(0, 0, k0), keep_cell=keep)
    ramp = np.minimum(1.0, (np.arange(nsteps) + 1.0) / max(1, ramp_steps))
    P.gravity = P.gravity * ramp[None, :]
    if world == 1:
        return P, None
    # global ids: (global particle-cell index, x fastest, z slowest) * 8 + Gauss point; carved-away cells leave gaps
    e = P.kept_cells.astype(np.int64)
    gcell = e + c0 * c * w
    gid = (gcell[:, None] * 8 + np.arange(8)[None, :]).ravel().astype(np.int32)
    cuts = np.array([(lay[r] + 2 + 0.5) * h for r in range(1, world)])
    slab = dict(rank=rank, world=world, axis=2, cuts=cuts, band_cells=band_cells, global_id=gid, n_global=8 * c * w * c,
                node_offset=k0 * (nx + 1) * (ny + 1), n_particles=int(8 * cum[-1] * w))
    return P, slab


def beam_3d(cells_per_unit=8, nsteps=20, gamma_lme=6.0, E=1e7, cfl=10.0, traction=-2.0e3):
    """BASELINE configs[4] shape: cantilever 8 x 1 x 1 on an H8 grid, GPxElement 8,
    Neo-Hookean, clamped at x = 0 (left face), tip traction on the last particle-cell layer ramped over
    the run; implicit Newmark-beta with dt = cfl x the explicit limit.  cells_per_unit = 32 -> 2.1e6
    particles."""
    c = cells_per_unit
    mat = ("Neo-Hookean-Wriggers", [1000.0, E, 0.3] + [0.0] * 13)
    cel = (E / 1000.0) ** 0.5 * 1.3
    P = structured_problem(3, (8 * c + 4, c + 4, c + 4), 1.0 / c, (8 * c, c, c), (0, 2, 2), mat, nsteps, cfl, cel,
                           (0.0, 0.0, 0.0), gamma_lme=gamma_lme, fixed=("left",), rollers=())
    x = P.fields["x_GC"]
    tip = np.nonzero(x[:, 0] > 8.0 - 1.0 / c)[0].astype(np.int32)
    dr = np.zeros((3, nsteps), np.int32)
    dr[2, :] = 1
    val = np.zeros((3, nsteps))
    val[2, :] = traction * np.minimum(1.0, (np.arange(nsteps) + 1) / max(nsteps, 1))
    P.neumann.append(dict(nodes=tip, dir=dr, val=val))
    # Phi.Area_0 of a 3D Neumann load (U-Verlet.c:847-849): the 8 particles of a tip cell share its face, V0 / h each
    P.fields["Area_0"] = P.fields["Vol_0"] * c
    return P


def cube_slab_3d(rank, world, cells=126, nsteps=100, gamma_lme=6.0, band_cells=6, material=NH_C1, xy=None,
                 compress=True):
    """BASELINE configs[2] split into `world` slabs along z: the global problem is cube_3d(cells)
    (cells^3 particle cells x GPxElement 8; 126 -> 16,003,008 particles); slab r owns the particle-cell layers
    [r*cells//world, (r+1)*cells//world) and builds only the sub-mesh within band_cells + 2 layers of them.
    Strong scaling of a fixed global problem.  Returns (Problem, slab dict)."""
    c = cells
    cxy = c if xy is None else xy
    h = 1.0 / c
    nx = ny = cxy + 4
    nz_glob = c + 4
    lay = [r * c // world for r in range(world + 1)]
    pad = band_cells + 2
    k0 = max(0, lay[rank] - pad) if rank > 0 else 0
    k1 = min(nz_glob, lay[rank + 1] + pad) if rank < world - 1 else nz_glob
    c0 = max(0, lay[rank] - 1) if rank > 0 else 0
    c1 = min(c, lay[rank + 1] + 1) if rank < world - 1 else c
    cel = (material[1][1] / material[1][0]) ** 0.5 * 1.3
    P = structured_problem(3, (nx, ny, k1 - k0), h, (cxy, cxy, c1 - c0), (2, 2, c0 - k0), material, nsteps, 0.5, cel,
                           (0.0, 0.0, 0.0 if compress else -9.81), gamma_lme=gamma_lme,
                           fixed=("bottom",) if k0 == 0 else (), rollers=("left", "right", "front", "back"),
                           cell_offset=(0, 0, k0))
    if compress:
        _top_platen(P, c, nx, ny, nsteps, 0.1 / 500.0, layer_offset=k0)
    e = np.arange(cxy * cxy * (c1 - c0), dtype=np.int64)
    gcell = e + c0 * cxy * cxy                           # cells are numbered x fastest, z slowest
    gid = (gcell[:, None] * 8 + np.arange(8)[None, :]).ravel().astype(np.int32)
    cuts = np.array([(lay[r] + 0.5) * h for r in range(1, world)])
    slab = dict(rank=rank, world=world, axis=2, cuts=cuts, band_cells=band_cells, global_id=gid,
                n_global=8 * cxy * cxy * c, node_offset=k0 * (nx + 1) * (ny + 1))
    return P, slab
