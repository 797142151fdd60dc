#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=${1:-33}
mkdir -p gpurun_out /tmp/ncu
timeout 1500 python bench.py --steps 10 --warmup 5 --profile > gpurun_out/bench${TAG}.json 2> gpurun_out/bench${TAG}.err
tail -c 1200 gpurun_out/bench${TAG}.json; tail -18 gpurun_out/bench${TAG}.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench${TAG}_ref.json 2> gpurun_out/bench${TAG}_ref.err
head -c 300 gpurun_out/bench${TAG}_ref.json; echo
export B2P_OPTS=push_streams=1,sort_streams=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_push -s 40 -c 1 -o /tmp/ncu/push -f python tools/microbench.py --cells 128 --laps 1 "" > gpurun_out/ncu${TAG}.log 2>&1
ncu -i /tmp/ncu/push.ncu-rep --page raw --csv > gpurun_out/r${TAG}_push_raw.csv 2>/dev/null
ncu -i /tmp/ncu/push.ncu-rep --page details > gpurun_out/r${TAG}_push_details.txt 2>/dev/null
ncu -i /tmp/ncu/push.ncu-rep --page source --csv > gpurun_out/r${TAG}_push_source.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:"k_sort_count|k_sort_scatter|k_sort_fix|k_sort_gather|k_nodal|k_edge_gather|k_filter|k_halo|k_push_b|k_push_e" -s 60 -c 30 -o /tmp/ncu/rest -f python tools/microbench.py --cells 128 --laps 1 "" >> gpurun_out/ncu${TAG}.log 2>&1
ncu -i /tmp/ncu/rest.ncu-rep --page raw --csv > gpurun_out/r${TAG}_rest_raw.csv 2>/dev/null
unset B2P_OPTS
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches${TAG}.csv python bench.py --cells 256 --steps 5 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench${TAG}_under_ncu.log 2>&1
gzip -f gpurun_out/launches${TAG}.csv
ls -la gpurun_out | tail -15
