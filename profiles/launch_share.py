"""ncu launch list (--metrics gpu__time_duration.sum --csv --log-file X.csv) of `bench.py --steps K --warmup W ...` ->
per-kernel share of the K timed steps.

    python profiles/launch_share.py launches.csv K [bench_line.json] > profiles/rNN_launch_share_<workload>.txt

Window = the launches between the (W+1)-th and the (W+K+1)-th k_search after the initialisation's (the searches are
counted from the END of the list: the profile run and the parity / e2e legs are switched off in the profiled command, so
the last K+2 steps of the process are [K timed steps, 2 per-kernel profile steps])."""
import csv
import json
import re
import sys


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("<unnamed>::", "").replace("void ", "").strip()


def main():
    path, K = sys.argv[1], int(sys.argv[2])
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    launches = [(short(r[kn]), float(r[mv]) / 1e3) for r in rows[1:]]
    searches = [i for i, (n, _) in enumerate(launches) if n.startswith("k_search")]
    # the last 2 searches open the two per-kernel profile steps; the K before them are the timed steps
    lo, hi = searches[-(K + 2)], searches[-2]
    # the download between the timed steps and the profile steps (neighbour count for the roofline) is not part of a step
    NOT_STEP = ("k_soa_to_aos", "k_unpermute", "k_expand_lists", "k_export")
    win = [(n, t) for n, t in launches[lo:hi] if not n.startswith(NOT_STEP)]
    tot = sum(t for _, t in win)
    agg = {}
    for n, t in win:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += t
    print(f"# window: the {K} timed steps = launches [{lo}, {hi}) of {len(launches)} captured; per-launch times under ncu are")
    print("# cold-cache and serialised: the SHARE of the step is what compares with bench.py's live CUDA-event numbers")
    print(f"# total {tot / K:.1f} us per step under ncu")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n:60s} launches/step {c / K:5.2f}  us/step {t / K:10.1f}  share {100 * t / tot:5.1f}%")
    if len(sys.argv) > 3:
        l = json.loads(open(sys.argv[3]).read().strip().splitlines()[-1])
        pk = l["roofline"]["per_kernel"]
        live = {k: v["ms"] * v["launches_per_step"] * 1e3 for k, v in pk.items() if k != "reorder"}
        s = sum(live.values())
        print(f"# live (same command without ncu): ms_per_step {l['ms_per_step']:.4f}; per-kernel CUDA events (sum {s:.1f} us, re-sort excluded):")
        for k, v in sorted(live.items(), key=lambda kv: -kv[1]):
            print(f"#   {k:40s} us/step {v:10.1f}  share {100 * v / s:5.1f}%")


if __name__ == "__main__":
    main()
