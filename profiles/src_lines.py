"""Per CUDA-source-line totals of an `ncu --page source --csv --print-source cuda,sass` dump:
samples, warp instructions, shared-memory wavefronts (ideal / actual) and the top stall reasons."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
which = sys.argv[2] if len(sys.argv) > 2 else None
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
secs = [i for i, r in enumerate(rows) if r and r[0] == "Function Name"]
secs.append(len(rows) + 1)
seen = set()
for a, b in zip(secs[:-1], secs[1:]):
    name = rows[a][1]
    if name in seen or (which and which not in name):
        continue
    seen.add(name)
    hdr = rows[a + 1]
    data = [r for r in rows[a + 2:b - 1] if len(r) == len(hdr)]
    il = hdr.index("Line No")
    isamp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")
    ith = hdr.index("Thread Instructions Executed")
    iw, iwi = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
    stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {}
    for r in data:
        try:
            ln = int(r[il])
        except ValueError:
            continue
        d = agg.setdefault(ln, dict(samp=0, ex=0, th=0, w=0, wi=0, st={}))
        d["samp"] += int(r[isamp] or 0); d["ex"] += int(r[iex] or 0); d["th"] += int(r[ith] or 0)
        d["w"] += int(r[iw] or 0); d["wi"] += int(r[iwi] or 0)
        for h in stall:
            d["st"][h] = d["st"].get(h, 0) + int(r[hdr.index(h)] or 0)
    tot = sum(d["samp"] for d in agg.values()) or 1
    tex = sum(d["ex"] for d in agg.values()) or 1
    tw = sum(d["w"] for d in agg.values())
    print("====", name[:90], "samples", tot, "warp-instr", tex, "smem wavefronts", tw)
    src = {}
    try:
        for k, line in enumerate(open(rows[a - 1][1]).read().split("\n")):
            src[k + 1] = line.strip()
    except Exception:
        pass
    for ln, d in sorted(agg.items(), key=lambda kv: -kv[1]["samp"])[:top_n]:
        st = sorted(d["st"].items(), key=lambda kv: -kv[1])[:2]
        print(f"  L{ln:4d} samp {d['samp'] / tot * 100:5.1f}% instr {d['ex'] / tex * 100:5.1f}% lanes {d['th'] / max(d['ex'], 1):4.1f} "
              f"wf {d['w'] / max(tw, 1) * 100:5.1f}% (x{d['w'] / max(d['wi'], 1):3.1f}) {[(k[6:], round(v / tot * 100, 1)) for k, v in st]}  {src.get(ln, '')[:70]}")

# optional 4th argument: comma-separated "name:lo-hi" line ranges -> totals per region of the LAST function printed
if len(sys.argv) > 4:
    print("  -- regions")
    for spec in sys.argv[4].split(","):
        nm, rg = spec.split(":")
        lo, hi = (int(x) for x in rg.split("-"))
        sel = [d for ln, d in agg.items() if lo <= ln <= hi]
        print(f"  {nm:14s} samp {sum(d['samp'] for d in sel) / tot * 100:5.1f}% instr {sum(d['ex'] for d in sel) / tex * 100:5.1f}% "
              f"wf {sum(d['w'] for d in sel) / max(tw, 1) * 100:5.1f}% lanes {sum(d['th'] for d in sel) / max(sum(d['ex'] for d in sel), 1):4.1f}")
