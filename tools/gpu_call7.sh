#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
B="python bench.py --cells 128 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:^k_push$' -s 112 -c 2 -o gpurun_out/prof_push_fused -f $B > gpurun_out/ncu_push.log 2>&1
tail -n 3 gpurun_out/ncu_push.log
