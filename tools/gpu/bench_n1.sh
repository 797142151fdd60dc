#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 10 --warmup 5 --profile > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 700 gpurun_out/bench_n1.json; head -4 gpurun_out/bench_n1.err
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
