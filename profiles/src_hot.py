"""Per-phase stall samples of an `ncu --page source --csv` dump: phases are split at BAR.SYNC."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
which = sys.argv[2] if len(sys.argv) > 2 else None
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
secs.append(len(rows))
seen = set()
for a, b in zip(secs[:-1], secs[1:]):
    name = rows[a][1]
    if name in seen or (which and which not in name):
        continue
    seen.add(name)
    hdr = rows[a + 1]
    data = rows[a + 2:b]
    isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[isamp]) for r in data)
    print("====", name[:70], "samples", tot, "warp-instr", sum(int(r[iex]) for r in data))
    ph, acc, accs, n0 = 0, 0, {}, 0
    for i, r in enumerate(data):
        acc += int(r[isamp])
        for h in stall:
            accs[h] = accs.get(h, 0) + int(r[hdr.index(h)])
        if "BAR.SYNC" in r[isrc] or i == len(data) - 1:
            top = sorted(accs.items(), key=lambda kv: -kv[1])[:3]
            print(f"  phase {ph}: instr {n0}-{i}: {acc / tot * 100:5.1f}%  {[(k[6:], round(v / tot * 100, 1)) for k, v in top]}")
            ph, acc, accs, n0 = ph + 1, 0, {}, i + 1
    for i, r in enumerate(data):
        n = int(r[isamp])
        if n > 0.012 * tot:
            st = sorted(((int(r[hdr.index(h)]), h[6:]) for h in stall), reverse=True)[:2]
            print(f"   {i:5d} {n / tot * 100:5.1f}% ex={r[iex]:>8s} {r[isrc][:56]:56s} {st}")
