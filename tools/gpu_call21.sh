#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out /tmp/ncu
export B2P_OPTS=push_streams=1,sort_streams=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_sort_count|k_sort_scatter|k_sort_fix|k_gather$" -s 80 -c 5 -o /tmp/ncu/sort -f python tools/microbench.py --cells 128 --laps 1 "" > gpurun_out/ncu21.log 2>&1
ncu -i /tmp/ncu/sort.ncu-rep --page details > gpurun_out/r21_sort_details.txt 2>/dev/null
ncu -i /tmp/ncu/sort.ncu-rep --page raw --csv > gpurun_out/r21_sort_raw.csv 2>/dev/null
ncu -i /tmp/ncu/sort.ncu-rep --page source --csv > gpurun_out/r21_sort_source.csv 2>/dev/null
tail -3 gpurun_out/ncu21.log
