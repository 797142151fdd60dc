#!/usr/bin/env python
"""Turn what tools/gpu/r02_evidence.sh brings back in gpurun_out/ into the summaries kept under profiles/.
usage: tools/profile_summaries.py TAG     (e.g. b  ->  profiles/r02_*_b.json, profiles/r02_push_traffic.json)"""
import collections
import csv
import gzip
import json
import re
import sys

from ncu_summary import KEYS

SETUP = ("k_inject", "k_fill", "k_init")


def short(name):
    return re.sub(r"\(.*", "", name).strip()


def launches(path, out):
    rows = list(csv.reader(gzip.open(path, "rt")))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[h + 1:]:
        if len(r) <= mv:
            continue
        ns = float(r[mv].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r[mu], 1.0)
        t = tot[short(r[kn])]
        t[0] += 1
        t[1] += ns / 1e3
    total = sum(v[1] for v in tot.values())
    setup = sum(v[1] for k, v in tot.items() if any(s in k for s in SETUP))
    ks = []
    for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        is_setup = any(s in k for s in SETUP)
        ks.append({"kernel": k, "launches": n, "total_us": round(us, 1), "share": round(us / total, 4),
                   "share_excl_setup": None if is_setup else round(us / (total - setup), 4), "avg_us": round(us / n, 2)})
    json.dump({"source": "ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 on `python bench.py --cells 256 "
                         "--steps 5 --warmup 5 --no-cpu-baseline --no-e2e` (first 4000 launches: set-up + ~8 laps; cold-cache, "
                         "serialised: compare shares, not absolute times)",
               "total_us": total, "total_us_excl_setup": total - setup, "kernels": ks}, open(out, "w"), indent=1)
    return ks


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": short(r[hdr.index("Kernel Name")]), "block": r[hdr.index("Block Size")], "grid": r[hdr.index("Grid Size")]}
        for i, hfull in enumerate(hdr):
            h = hfull.split(".", 2)[-1] if hfull.count(".") >= 2 and not hfull.startswith(("dram__", "gpu__", "sm", "l1tex", "lts", "launch")) else hfull
            for cand in (hfull, h):
                if cand in KEYS or cand.startswith("smsp__average_warps_issue_stalled") or cand == "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed":
                    try:
                        d[cand + (f" [{units[i]}]" if units[i] else "")] = float(r[i].replace(",", ""))
                    except ValueError:
                        pass
                    break
        out.append(d)
    return out


def main():
    tag = sys.argv[1]
    ks = launches(f"gpurun_out/r02_launches_{tag}.csv.gz", f"profiles/r02_launches_summary_{tag}.json")
    for k in ks[:8]:
        print(k)
    push = raw(f"gpurun_out/r02_{tag}_push_raw.csv")
    rest = raw(f"gpurun_out/r02_{tag}_rest_raw.csv")
    import os
    sort = raw(f"gpurun_out/r02_{tag}_sort_raw.csv") if os.path.exists(f"gpurun_out/r02_{tag}_sort_raw.csv") else []
    json.dump({"how": "ncu --set full --clock-control none on tools/microbench.py --cells 128 (8 tiles of 64^3: ONE k_push launch covers "
                      "16 containers of 4.19 M particles = 67.1 M alive particles), worker streams off; one launch per entry",
               "k_push (fused push+deposit, 16 containers per launch)": push,
               "counting sort (16 containers per launch, five laps after their last sort)": sort, "other kernels": rest},
              open(f"profiles/r02_ncu_summary_{tag}.json", "w"), indent=1)
    alive = 2 * 16 * 128 ** 3
    d = push[-1]
    rd = next(v for k, v in d.items() if k.startswith("dram__bytes_read.sum"))
    wr = next(v for k, v in d.items() if k.startswith("dram__bytes_write.sum"))
    unit = next(k for k in d if k.startswith("dram__bytes_read.sum"))
    scale = 1e6 if "Mbyte" in unit else (1e9 if "Gbyte" in unit else (1e3 if "Kbyte" in unit else 1.0))
    json.dump({"source": f"profiles/r02_ncu_summary_{tag}.json: dram__bytes_read.sum + dram__bytes_write.sum of one k_push launch "
                         "(16 containers, fused push+deposit)",
               "traffic_bytes_per_launch": (rd + wr) * scale, "alive_particles_per_launch": alive,
               "bytes_per_alive_particle": (rd + wr) * scale / alive}, open("profiles/r02_push_traffic.json", "w"), indent=1)
    for d in push:
        print(json.dumps(d, indent=1))


if __name__ == "__main__":
    sys.path.insert(0, "tools")
    main()
