#!/bin/bash
# Round-2 evidence: the default bench line + reference arm, `ncu --set full` of the push and of the other kernels,
# and the launch list of a 256^3 bench (shares only: cold-cache, serialised).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=${1:-b}
mkdir -p gpurun_out /tmp/ncu
timeout 1500 python bench.py --steps 20 --warmup 5 --profile > gpurun_out/r02_bench_${TAG}_n1_512.json 2> gpurun_out/r02_bench_${TAG}_n1_512.err
tail -c 600 gpurun_out/r02_bench_${TAG}_n1_512.json; echo; head -20 gpurun_out/r02_bench_${TAG}_n1_512.err
timeout 1200 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r02_bench_${TAG}_reference_arm.json 2> gpurun_out/r02_bench_${TAG}_ref.err
head -c 400 gpurun_out/r02_bench_${TAG}_reference_arm.json; echo
export B2P_OPTS=push_streams=1,sort_streams=0
timeout 900 ncu --set full --clock-control none --import-source on -k k_push -s 7 -c 1 -o /tmp/ncu/push -f python tools/microbench.py --cells 128 --laps 1 "" > gpurun_out/r02_ncu_${TAG}.log 2>&1
ncu -i /tmp/ncu/push.ncu-rep --page raw --csv > gpurun_out/r02_${TAG}_push_raw.csv 2>/dev/null
ncu -i /tmp/ncu/push.ncu-rep --page source --csv > gpurun_out/r02_${TAG}_push_source.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:"k_pack_|k_append|k_nodal|k_edge_gather|k_filter|k_halo|k_push_b|k_push_e|k_J_exchange" -s 40 -c 40 -o /tmp/ncu/rest -f python tools/microbench.py --cells 128 --laps 1 "" >> gpurun_out/r02_ncu_${TAG}.log 2>&1
ncu -i /tmp/ncu/rest.ncu-rep --page raw --csv > gpurun_out/r02_${TAG}_rest_raw.csv 2>/dev/null
# the six kernels of the third sort (containers five laps after their last sort)
timeout 600 ncu --set full --clock-control none -k regex:k_sort_ -s 12 -c 6 -o /tmp/ncu/sort -f python tools/microbench.py --cells 128 --laps 1 "" >> gpurun_out/r02_ncu_${TAG}.log 2>&1
ncu -i /tmp/ncu/sort.ncu-rep --page raw --csv > gpurun_out/r02_${TAG}_sort_raw.csv 2>/dev/null
unset B2P_OPTS
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_${TAG}.csv python bench.py --cells 256 --steps 5 --warmup 5 --no-cpu-baseline --no-e2e --no-emf > gpurun_out/r02_bench_${TAG}_under_ncu.log 2>&1
gzip -f gpurun_out/r02_launches_${TAG}.csv
ls -la gpurun_out | tail -8
