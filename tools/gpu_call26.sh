#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 10 --warmup 5 --profile > gpurun_out/bench26.json 2> gpurun_out/bench26.err
tail -c 2500 gpurun_out/bench26.json; tail -25 gpurun_out/bench26.err
