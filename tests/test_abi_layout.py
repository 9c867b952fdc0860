"""The ctypes mirrors of nl-partsol_b200/nlps_b200/engine.py against include/nlps_b200.h as the C compiler lays it out:
size of every struct and offset of every member.  CPU only (gcc).  A binding that drifts from the header reads garbage
past the end of a struct -- the C shims are protected by the host Makefile's dependency on the header, the Python view by
this test."""
import ctypes
import os
import subprocess
import tempfile

from nlps_b200 import engine

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
PAIRS = (("nlps_mesh", engine.Mesh), ("nlps_load", engine.Load), ("nlps_material", engine.Material),
         ("nlps_solver", engine.Solver), ("nlps_particles", engine.Particles), ("nlps_msg", engine.Msg),
         ("nlps_slab", engine.Slab), ("nlps_newmark", engine.Newmark), ("nlps_newmark_stats", engine.NewmarkStats))


def test_ctypes_mirrors_match_the_header():
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "nlps_b200.h"', 'int main(void) {']
    for cname, cls in PAIRS:
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    with tempfile.TemporaryDirectory() as tmp:
        src, exe = os.path.join(tmp, "layout.c"), os.path.join(tmp, "layout")
        with open(src, "w") as f:
            f.write("\n".join(lines) + "\n")
        # a member of a ctypes mirror that the header does not have is a compile error here
        subprocess.run(["gcc", "-std=gnu11", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    c = {k: int(v) for k, v in (ln.split() for ln in out.splitlines())}
    for cname, cls in PAIRS:
        assert ctypes.sizeof(cls) == c[cname], f"sizeof({cname}): ctypes {ctypes.sizeof(cls)} != C {c[cname]}"
        for fname, _ in cls._fields_:
            assert getattr(cls, fname).offset == c[f"{cname}.{fname}"], f"{cname}.{fname}"
    # every struct member of the header is mirrored (a member missing in a mirror shows up as a smaller ctypes struct
    # only when it is the last one: count them)
    import re
    hdr = open(os.path.join(ROOT, "include", "nlps_b200.h")).read()
    for cname, cls in PAIRS:
        body = re.search(r"typedef struct " + cname + r" \{(.*?)\} " + cname + ";", hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        members = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = re.sub(r"^(const\s+)?(unsigned\s+)?(long\s+long|double|int|void|char|nlps_\w+)\b", "", decl)
            members += [re.sub(r"[\*\s]|\[.*\]", "", n) for n in names.split(",")]
        assert [m for m in members if m] == [f for f, _ in cls._fields_], (cname, members)
