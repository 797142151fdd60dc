#!/bin/bash
# GPU call 2: new gpu tests, 512^3 bench, ncu --set full of push/deposit at steady state
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log )
( timeout 900 python bench.py --cells 512 --steps 5 --no-cpu-baseline --profile > gpurun_out/bench_512.json 2> gpurun_out/bench_512.err; echo "rc=$?" >> gpurun_out/bench_512.err )
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_push|k_deposit_zigzag" -s 238 -c 4 -o gpurun_out/prof_push_deposit -f python bench.py --cells 128 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "rc=$?" >> gpurun_out/ncu_full.log )
tail -3 gpurun_out/pytest_gpu.log
tail -22 gpurun_out/bench_512.err
cut -c1-300 gpurun_out/bench_512.json
tail -3 gpurun_out/ncu_full.log
