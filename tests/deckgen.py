"""Deck + GiD mesh generator for NL-PartSol's input grammar.

Test infrastructure: the reference ships no example decks, so the
fixtures that pin the oracle are produced by driving the reference's OWN parser
(`/root/reference/nl-partsol/src/InOutFun/**`) over decks written here.  The
same spec object also produces plain numpy arrays (``structured_arrays``) so
that the CUDA engine and the C oracle port can be fed the identical problem on
the GPU box, where /root/reference does not exist.

Grammar citations (reference file:line): NLPS-Solver `Read_GramsTime.c:96-203`,
gravity `Read_Generate_Gravity_Field.c:160-260`, GramsBox `Read_GramsBox.c:235-266`,
GramsBoundary `NLPS-Read-u-Dirichlet-Boundary-Conditions.c:46-318`,
One-Phase-Analysis `Generate-One-Phase-Analysis.c:417-444`, GramsShapeFun
`Read_GramsShapeFun.c:84-176`, Define-Material `Read_GramsMaterials2.c:103-205`,
GiD mesh `Read-GID-Mesh.c:225-430`, curves `ReadCurve.c:44-116`.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np


@dataclass
class Material:
    model: str = "Neo-Hookean-Wriggers"
    params: dict = field(default_factory=lambda: dict(rho=1000.0, E=1.0e6, nu=0.3))


@dataclass
class DeckSpec:
    """2D plane-strain block on a structured Q4 background grid."""
    nx: int = 12            # background cells in x
    ny: int = 12
    h: float = 0.0625       # background cell size
    origin: tuple = (0.0, 0.0)
    pnx: int = 8            # particle-mesh cells in x
    pny: int = 8
    ph: float = 0.0625      # particle-mesh cell size
    porigin: tuple = (0.125, 0.0)
    gpx: int = 4            # GPxElement
    scheme: str = "NPC-FS"
    cfl: float = 0.5
    cel: float = 31.622776601683793
    nsteps: int = 20
    gravity: tuple = (0.0, -9.81)
    gamma: float = 3.0
    shape: str = "LME"      # GramsShapeFun type: "LME" | "aLME" (Read_GramsShapeFun.c:84-176)
    tol_zero: float = 1e-6
    tol_wrapper: float = 1e-10
    max_iter: int = 10
    material: Material = field(default_factory=Material)
    # Dirichlet sets: list of (name, node-selector, {"V.x": scale or None, ...}, curve kind)
    # node selector: "bottom" | "left" | "right" | "top"
    dirichlet: list = field(default_factory=lambda: [
        ("Bottom", "bottom", {"V.x": 0.0, "V.y": 0.0}, "CONSTANT_CURVE"),
        ("Left", "left", {"V.x": 0.0, "V.y": None}, "CONSTANT_CURVE"),
        ("Right", "right", {"V.x": 0.0, "V.y": None}, "CONSTANT_CURVE"),
    ])
    # Neumann sets (NLPS-Read-u-Neumann-Boundary-Conditions.c:46-210): list of (name, particle-mesh element ids,
    # {"T.x": scale or None, "T.y": ...}, curve kind); the loaded particles are element * GPxElement + 0 .. GPxElement-1
    neumann: list = field(default_factory=list)
    # further materials: list of (Material, particle-mesh element ids); idx = 1, 2, ... in this order
    # (Generate-One-Phase-Analysis.c:497-560: each assignment sets MatIdx of element * GPxElement + 0 .. GPxElement-1)
    more_materials: list = field(default_factory=list)
    out_every: int = 1000000
    solver_extra: dict = field(default_factory=dict)   # e.g. Beta-Newmark-beta, TOL-Newmark-beta, Max-Iter


def background_nodes(spec: DeckSpec):
    nxn, nyn = spec.nx + 1, spec.ny + 1
    ii, jj = np.meshgrid(np.arange(nxn), np.arange(nyn), indexing="xy")
    x = spec.origin[0] + ii.ravel() * spec.h
    y = spec.origin[1] + jj.ravel() * spec.h
    return np.stack([x, y], axis=1)


def q4_connectivity(nx, ny):
    e = np.arange(nx * ny)
    i, j = e % nx, e // nx
    n1 = j * (nx + 1) + i
    return np.stack([n1, n1 + 1, n1 + 1 + (nx + 1), n1 + (nx + 1)], axis=1).astype(np.int32)


def boundary_nodes(spec: DeckSpec, which: str):
    nxn, nyn = spec.nx + 1, spec.ny + 1
    if which == "bottom":
        return np.arange(nxn)
    if which == "top":
        return (nyn - 1) * nxn + np.arange(nxn)
    if which == "left":
        return np.arange(nyn) * nxn
    if which == "right":
        return np.arange(nyn) * nxn + (nxn - 1)
    raise ValueError(which)


def _write_gid(path, coords, conn):
    with open(path, "w") as f:
        f.write("MESH dimension 2 ElemType Quadrilateral Nnode 4\n")
        f.write("Coordinates\n")
        for k, (x, y) in enumerate(coords):
            f.write(f"{k + 1} {float(x)!r} {float(y)!r} 0.0\n")
        f.write("End Coordinates\n")
        f.write("Elements\n")
        for k, row in enumerate(conn):
            f.write(f"{k + 1} " + " ".join(str(int(v) + 1) for v in row) + "\n")
        f.write("End Elements\n")


def _write_curve(path, kind, scale, num):
    with open(path, "w") as f:
        f.write(f"DAT_CURVE NUM#{num}\n")
        f.write(f"{kind} SCALE#{scale!r}\n")


def write_deck(spec: DeckSpec, outdir: str) -> str:
    """Write deck + meshes + lists + curves into outdir; return the deck path."""
    os.makedirs(outdir, exist_ok=True)
    os.makedirs(os.path.join(outdir, "Results"), exist_ok=True)
    _write_gid(os.path.join(outdir, "Background.msh"), background_nodes(spec),
               q4_connectivity(spec.nx, spec.ny))
    pxn = spec.pnx + 1
    pii, pjj = np.meshgrid(np.arange(pxn), np.arange(spec.pny + 1), indexing="xy")
    pcoords = np.stack([spec.porigin[0] + pii.ravel() * spec.ph,
                        spec.porigin[1] + pjj.ravel() * spec.ph], axis=1)
    _write_gid(os.path.join(outdir, "Particles.msh"), pcoords,
               q4_connectivity(spec.pnx, spec.pny))
    with open(os.path.join(outdir, "AllElems.txt"), "w") as f:
        for e in range(spec.pnx * spec.pny):
            f.write(f"{e}\n")
    lines = []
    lines.append(f"NLPS-Solver (Type={spec.scheme}) {{")
    lines.append(f"  CFL={spec.cfl!r}")
    lines.append(f"  Cel={spec.cel!r}")
    lines.append(f"  N={spec.nsteps}")
    for k, v in spec.solver_extra.items():
        lines.append(f"  {k}={v!r}")
    lines.append("}")
    lines.append("generate-gravity-field-constant")
    lines.append("{")
    lines.append(f" g.x {spec.gravity[0]!r}")
    lines.append(f" g.y {spec.gravity[1]!r}")
    lines.append("}")
    lines.append("GramsBox (Type=GID,File=Background.msh) {")
    for name, sel, dofs, kind in spec.dirichlet:
        with open(os.path.join(outdir, f"{name}.txt"), "w") as f:
            for n in boundary_nodes(spec, sel):
                f.write(f"{int(n)}\n")
        lines.append(f"  GramsBoundary (File={name}.txt) {{")
        for dof, scale in dofs.items():
            if scale is None:
                lines.append(f"     BcDirichlet {dof} NULL")
            else:
                cname = f"{name}_{dof.replace('.', '')}.curve"
                _write_curve(os.path.join(outdir, cname), kind, scale, spec.nsteps)
                lines.append(f"     BcDirichlet {dof} {cname}")
        lines.append("  }")
    lines.append("}")
    lines.append(f"One-Phase-Analysis (File=Particles.msh,GPxElement={spec.gpx}) {{")
    lines.append("}")
    for name, elems, comps, kind in spec.neumann:
        with open(os.path.join(outdir, f"{name}.txt"), "w") as f:
            for e in elems:
                f.write(f"{int(e)}\n")
        lines.append(f"Define-Neumann-Boundary(File={name}.txt)")
        lines.append("{")
        for comp, scale in comps.items():
            if scale is None:
                lines.append(f"  {comp} NULL")
            else:
                cname = f"{name}_{comp.replace('.', '')}.curve"
                _write_curve(os.path.join(outdir, cname), kind, scale, spec.nsteps)
                lines.append(f"  {comp} {cname}")
        lines.append("}")
    lines.append(f"GramsShapeFun (Type={spec.shape}) {{")
    lines.append(f"  gamma={spec.gamma!r}")
    lines.append(f"  TOL-Zero={spec.tol_zero!r}")
    lines.append(f"  TOL-Wrapper={spec.tol_wrapper!r}")
    lines.append(f"  MaxIter={spec.max_iter}")
    lines.append("  wrapper=Newton-Raphson")
    lines.append("}")
    lines.append(f"Define-Material(idx=0,Model={spec.material.model})")
    lines.append("{")
    for k, v in spec.material.params.items():
        lines.append(f"  {k}={v!r}")
    lines.append("}")
    lines.append("Assign-material-to-particles (MatIdx=0,Particles=AllElems.txt)")
    for idx, (mat, elems) in enumerate(spec.more_materials, start=1):
        lines.append(f"Define-Material(idx={idx},Model={mat.model})")
        lines.append("{")
        for k, v in mat.params.items():
            lines.append(f"  {k}={v!r}")
        lines.append("}")
        with open(os.path.join(outdir, f"Mat{idx}Elems.txt"), "w") as f:
            for e in elems:
                f.write(f"{int(e)}\n")
        lines.append(f"Assign-material-to-particles (MatIdx={idx},Particles=Mat{idx}Elems.txt)")
    lines.append(f"GramsOutputs (i={spec.out_every}) {{")
    lines.append("  DIR=Results")
    lines.append("  Out-velocity=true")
    lines.append("  Out-stress=true")
    lines.append("  Out-displacement=true")
    lines.append("}")
    deck = os.path.join(outdir, "deck.nlp")
    with open(deck, "w") as f:
        f.write("\n".join(lines) + "\n")
    return deck
