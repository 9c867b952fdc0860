"""ncu raw page (ncu -i X.ncu-rep --page raw --csv) -> profiles/ncu_traffic.json entry for one workload.

    python profiles/make_traffic.py raw.csv c3 16003008 "profiles/r02_ncu_full_c3.txt"

The entry is keyed by the machine code of the captured kernels: `sass` = sha256 of the SASS of every kernel whose name
matches a captured one, taken from the stamp the build left next to the library (nl-partsol_b200/libnlps_b200.sass.json).
bench.py quotes `roofline.traffic` only when the library it runs holds exactly these kernels, else it prints null."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

NAMES = (("cw_kin", "kin_stress_p2g_force"), ("cw_lme_p2g", "lme_p2g_mass_disp"), ("cw_g2p", "g2p_update"),
         ("k_kin_force", "kin_stress_p2g_force"), ("k_lme_p2g", "lme_p2g_mass_disp"), ("k_g2p", "g2p_update"))


def main():
    raw, workload, particles, source = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    rows = list(csv.reader(open(raw)))
    hdr = rows[0]
    ki, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    ur, uw = unit[rows[1][ir]], unit[rows[1][iw]]
    acc = {}
    for r in rows[2:]:
        for key, name in NAMES:
            if key in r[ki]:
                acc.setdefault(name, []).append(float(r[ir]) * ur + float(r[iw]) * uw)
                break
    kernels = {k: sum(v) / len(v) for k, v in acc.items()}
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    stamp = json.load(open(os.path.join(ROOT, "nl-partsol_b200", "libnlps_b200.sass.json")))
    keys = {key for key, name in NAMES if name in kernels and any(key in r[ki] for r in rows[2:])}
    sass = {k: v for k, v in stamp.items() if any(k == key or k.startswith(key + "<") for key in keys)}
    data[workload] = {"hash": bench.library_hash(), "sass": sass, "particles": particles, "source": source,
                      "kernels": {k: int(v) for k, v in kernels.items()},
                      "launches_captured": {k: len(v) for k, v in acc.items()}}
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(data[workload], indent=1))


if __name__ == "__main__":
    main()
