#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_golden.py tests/test_shock_gpu.py tests/test_scale_properties_gpu.py -m gpu -x -q > gpurun_out/pytest_filter.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_filter.log )
tail -n 3 gpurun_out/pytest_filter.log
timeout 600 python tools/microbench.py --cells 256 --laps 5 --out gpurun_out/micro_filter.json "push_streams=1,sort_streams=1" 2>&1 | grep -v "^ *per lap" | tail -2
timeout 600 python bench.py --workload emf-wave --steps 5 --warmup 3 > gpurun_out/bench_emf2.json 2> gpurun_out/bench_emf2.err
tail -c 1500 gpurun_out/bench_emf2.json; tail -5 gpurun_out/bench_emf2.err
