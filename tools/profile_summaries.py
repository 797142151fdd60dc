#!/usr/bin/env python
"""Turn what tools/gpu/collect_evidence.sh brings back in gpurun_out/ into the summaries kept under profiles/.
usage: tools/profile_summaries.py TAG OUT_SUFFIX     (e.g. 30 i  ->  profiles/r01_*_i.json)"""
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
    tag, suf = sys.argv[1], sys.argv[2]
    ks = launches(f"gpurun_out/launches{tag}.csv.gz", f"profiles/r01_launches_summary_{suf}.json")
    for k in ks[:8]:
        print(k)
    push = raw(f"gpurun_out/r{tag}_push_raw.csv")
    rest = raw(f"gpurun_out/r{tag}_rest_raw.csv")
    json.dump({"how": "ncu --set full --clock-control none on tools/microbench.py --cells 128 (8 tiles of 64^3, 4.19 M particles per "
                      "container), worker streams off; one launch per entry", "k_push (fused push+deposit)": push,
               "other kernels": rest}, open(f"profiles/r01_ncu_summary_{suf}.json", "w"), indent=1)
    for d in push:
        print(json.dumps(d, indent=1))


if __name__ == "__main__":
    sys.path.insert(0, "tools")
    main()
