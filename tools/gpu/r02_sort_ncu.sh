#!/bin/bash
# ncu --set full of the counting-sort kernels (third sort of the microbench = a container five laps after its last sort)
# and of the migration kernels, plus a 256^3 bench with --profile for the host enqueue figures
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out /tmp/ncu
export B2P_OPTS=push_streams=1,sort_streams=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sort_ -s 10 -c 5 -o /tmp/ncu/sort -f python tools/microbench.py --cells 128 --laps 1 "" > gpurun_out/r02_ncu_sort.log 2>&1
ncu -i /tmp/ncu/sort.ncu-rep --page raw --csv > gpurun_out/r02_sort_raw.csv 2>/dev/null
ncu -i /tmp/ncu/sort.ncu-rep --page source --csv -k k_sort_place > gpurun_out/r02_sort_place_source.csv 2>/dev/null
unset B2P_OPTS
timeout 600 python bench.py --cells 256 --steps 10 --warmup 5 --profile --no-cpu-baseline --no-emf > gpurun_out/r02_bench_256_hostenq.json 2> gpurun_out/r02_bench_256_hostenq.err
head -5 gpurun_out/r02_bench_256_hostenq.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_256_hostenq.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['wall_ms_per_step'], d['host_enqueue_ms_per_step'], d['gpu_launches'])
PY
ls -la gpurun_out | tail -5
