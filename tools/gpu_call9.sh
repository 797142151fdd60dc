#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log )
tail -n 3 gpurun_out/pytest_gpu.log
( timeout 900 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log ); tail -n 2 gpurun_out/smoke.log
( timeout 1200 python bench.py --profile > gpurun_out/bench_512.json 2> gpurun_out/bench_512.err; echo "rc=$?" >> gpurun_out/bench_512.err )
tail -n 22 gpurun_out/bench_512.err
cat gpurun_out/bench_512.json | cut -c1-300
( timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?" >> gpurun_out/bench_ref.err )
cat gpurun_out/bench_ref.json | cut -c1-200
