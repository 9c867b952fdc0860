"""Minimal readers of the legacy VTK files the particle writers produce (ASCII: the reference's
particle_results_vtk__InOutFun__; BINARY: nl-partsol_b200/host/b200_vtk_binary.h).  Test infrastructure."""
import numpy as np

_NCOMP = {"VECTORS": 3, "TENSORS": 9, "SCALARS": 1}


def read_ascii(path):
    tok = open(path).read().split("\n")
    out, i = {}, 0
    n = None
    while i < len(tok):
        w = tok[i].split()
        i += 1
        if not w:
            continue
        if w[0] == "POINTS":
            n = int(w[1])
            vals = []
            while len(vals) < 3 * n:
                vals += tok[i].split()
                i += 1
            out["POINTS"] = np.array(vals, float).reshape(n, 3)
        elif w[0] in ("CELLS", "CELL_TYPES"):
            cnt = int(w[2]) if w[0] == "CELLS" else int(w[1])
            vals = []
            while len(vals) < cnt:
                vals += tok[i].split()
                i += 1
            out[w[0]] = np.array(vals, int)
        elif w[0] in _NCOMP:
            name, typ, nc = w[1], w[2], _NCOMP[w[0]]
            if w[0] == "SCALARS":
                i += 1  # LOOKUP_TABLE
            vals = []
            while len(vals) < nc * n:
                vals += tok[i].split()
                i += 1
            out[name] = np.array(vals, float if typ == "double" else int).reshape(n, nc)
    return out


def read_binary(path):
    b = open(path, "rb").read()
    pos, out, n = 0, {}, None

    def line():
        nonlocal pos
        e = b.index(b"\n", pos)
        s = b[pos:e].decode()
        pos = e + 1
        return s
    assert line().startswith("# vtk DataFile")
    line()
    assert line().strip() == "BINARY"
    assert line().strip() == "DATASET UNSTRUCTURED_GRID"
    while pos < len(b):
        w = line().split()
        if not w:
            continue
        if w[0] == "POINTS":
            n = int(w[1])
            out["POINTS"] = np.frombuffer(b, ">f8", 3 * n, pos).reshape(n, 3).astype(float)
            pos += 24 * n
        elif w[0] == "CELLS":
            cnt = int(w[2])
            out["CELLS"] = np.frombuffer(b, ">i4", cnt, pos).astype(int)
            pos += 4 * cnt
        elif w[0] == "CELL_TYPES":
            cnt = int(w[1])
            out["CELL_TYPES"] = np.frombuffer(b, ">i4", cnt, pos).astype(int)
            pos += 4 * cnt
        elif w[0] == "POINT_DATA":
            assert int(w[1]) == n
        elif w[0] in _NCOMP:
            name, typ, nc = w[1], w[2], _NCOMP[w[0]]
            if w[0] == "SCALARS":
                assert line().startswith("LOOKUP_TABLE")
            dt, sz = (">f8", 8) if typ == "double" else (">i4", 4)
            out[name] = np.frombuffer(b, dt, nc * n, pos).reshape(n, nc).astype(float if typ == "double" else int)
            pos += sz * nc * n
    return out
