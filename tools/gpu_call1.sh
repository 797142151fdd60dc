#!/bin/bash
# first GPU call of the session: tests, bench at 256^3 and 512^3, ncu launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
( timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log ) 
( timeout 600 python bench.py --profile > gpurun_out/bench_256.json 2> gpurun_out/bench_256.err; echo "rc=$?" >> gpurun_out/bench_256.err )
( timeout 600 python bench.py --cells 512 --steps 5 --no-cpu-baseline --profile > gpurun_out/bench_512.json 2> gpurun_out/bench_512.err; echo "rc=$?" >> gpurun_out/bench_512.err )
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_256.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "rc=$?" >> gpurun_out/bench_under_ncu.log )
tail -3 gpurun_out/pytest_gpu.log
cat gpurun_out/bench_256.json | cut -c1-600
tail -5 gpurun_out/bench_512.err
cat gpurun_out/bench_512.json | cut -c1-400
