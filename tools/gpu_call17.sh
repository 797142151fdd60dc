#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export B2P_OPTS=push_streams=1,sort_streams=1
# full captures of the fused push: one lap after a sort, and 4 laps after
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_push -s 16 -c 2 -o gpurun_out/push_r17a -f python tools/microbench.py --cells 128 --laps 1 "" > gpurun_out/ncu17a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_push -s 70 -c 2 -o gpurun_out/push_r17b -f python tools/microbench.py --cells 128 --laps 1 "" > gpurun_out/ncu17b.log 2>&1
# the sort of lap 5 (nearly sorted input), all its kernels for the first 2 containers
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_sort|k_gather|DeviceScan|k_max|Radix" -s 80 -c 16 -o gpurun_out/sort_r17 -f python tools/microbench.py --cells 128 --laps 1 "" > gpurun_out/ncu17c.log 2>&1
# field kernels + the rest, once
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_filter|k_halo|k_edge|k_nodal|k_push_b|k_push_e|k_J_ex|k_zero|k_collect|k_append|k_gather_out" -s 200 -c 24 -o gpurun_out/fields_r17 -f python tools/microbench.py --cells 128 --laps 1 "" > gpurun_out/ncu17d.log 2>&1
tail -2 gpurun_out/ncu17a.log gpurun_out/ncu17b.log gpurun_out/ncu17c.log gpurun_out/ncu17d.log
# launch list of the bench command (reduced cube so the profiler finishes; same tiles, same launches per tile)
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches17.csv python bench.py --cells 256 --steps 5 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench17_under_ncu.log 2>&1
wc -l gpurun_out/launches17.csv
ls -la gpurun_out
