#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out /tmp/ncu
export B2P_OPTS=push_streams=1,sort_streams=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_push -s 50 -c 1 -o /tmp/ncu/push -f python tools/microbench.py --cells 128 --laps 1 "" > gpurun_out/ncu24.log 2>&1
ncu -i /tmp/ncu/push.ncu-rep --page details > gpurun_out/r24_push_details.txt 2>/dev/null
ncu -i /tmp/ncu/push.ncu-rep --page raw --csv > gpurun_out/r24_push_raw.csv 2>/dev/null
ncu -i /tmp/ncu/push.ncu-rep --page source --csv > gpurun_out/r24_push_source.csv 2>/dev/null
tail -2 gpurun_out/ncu24.log
