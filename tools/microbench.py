#!/usr/bin/env python
"""Kernel-variant microbenchmark: the config-5 plasma on a small block of 64^3 tiles, laps
timed per kernel class (CUDA events on the library stream) for a list of option settings.

  python tools/microbench.py --cells 128 --laps 10 "push_minb=5" "push_minb=6,deposit_agg=1" ...
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import runko_b200 as rb  # noqa: E402
from runko_b200._lib import check  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=128)
    ap.add_argument("--tile", type=int, default=64)
    ap.add_argument("--ppc", type=int, default=16)
    ap.add_argument("--laps", type=int, default=10)
    ap.add_argument("--out", default=None)
    ap.add_argument("settings", nargs="*", default=[""])
    args = ap.parse_args()
    L = rb.lib()
    check(L.b2p_init(0))
    conf, tpg, gb = bench.make_conf(args, 1)
    grid = rb.Grid(conf)
    tiles = []
    for i in range(tpg):
        for j in range(tpg):
            for k in range(tpg):
                t = rb.PicTile((i, j, k), conf)
                grid.add_tile(t)
                tiles.append(t)
    grid.set_uniform_B(0.0, 0.0, bench.binit(conf, args.ppc))
    grid.inject_thermal(args.ppc, 0.3, seed=42)
    for m in (rb.comm_mode.emf_E, rb.comm_mode.emf_B):
        grid.local_communication(m)
    lap = 0
    for _ in range(5):
        grid.step_pic(lap)
        lap += 1
    rb.sync()
    n_part = 2 * args.ppc * args.cells ** 3
    nk = L.b2p_profile_num_classes()
    names = [L.b2p_profile_class_name(k).decode() for k in range(nk)]
    results = []
    for setting in args.settings:
        for kv in filter(None, setting.split(",")):
            k, v = kv.split("=")
            check(L.b2p_set_option(k.encode(), int(v)))
        for _ in range(5):                      # settle (5 laps = one sort cycle)
            grid.step_pic(lap)
            lap += 1
        rb.sync()
        pms, pl = np.zeros(nk), np.zeros(nk, np.uint64)
        per_lap = []
        total_ms = 0.0
        for _ in range(args.laps):
            check(L.b2p_profile_enable(1))
            check(L.b2p_timer_start())
            grid.step_pic(lap)
            ms = C.c_float()
            check(L.b2p_timer_stop(C.byref(ms)))
            total_ms += ms.value
            a, b, c = np.zeros(nk), np.zeros(nk, np.uint64), np.zeros(nk)
            check(L.b2p_profile_report(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p)))
            check(L.b2p_profile_enable(0))
            pms += a
            pl += b
            per_lap.append({"lap_mod5": lap % 5, "ms": round(ms.value, 3),
                            "push_us": round(1e3 * a[names.index("push")] / max(int(b[names.index("push")]), 1), 1),
                            "deposit_us": round(1e3 * a[names.index("deposit")] / max(int(b[names.index("deposit")]), 1), 1)})
            lap += 1

        class _M:
            value = total_ms
        ms = _M()
        row = {"setting": setting, "ms_per_lap": ms.value / args.laps, "Gpush_per_s": n_part / (ms.value / args.laps) / 1e6,
               "us_per_launch": {names[k]: round(1e3 * pms[k] / int(pl[k]), 2) for k in range(nk) if pl[k]},
               "ms_per_lap_by_class": {names[k]: round(pms[k] / args.laps, 3) for k in range(nk) if pl[k]}, "per_lap": per_lap}
        results.append(row)
        u = row["us_per_launch"]
        print(f"{setting:44s} {row['ms_per_lap']:8.3f} ms/lap  {row['Gpush_per_s']:7.2f} Gp/s | push {u.get('push')}"
              f" deposit {u.get('deposit')} detect {u.get('detect_leavers')} place {u.get('sort_place')} gather {u.get('sort_gather')}"
              f" filter {u.get('filter')}", flush=True)
        print("      per lap (lap%5: ms push deposit): " + "  ".join(f"{q['lap_mod5']}:{q['ms']}/{q['push_us']}/{q['deposit_us']}" for q in per_lap[:5]), flush=True)
    if args.out:
        json.dump(results, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
