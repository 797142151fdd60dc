#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_push2 -s 7 -c 3 -o gpurun_out/r02_push2_a -f \
  python tools/microbench.py --cells 128 --laps 1 "push_kernel=2,push_streams=1" > gpurun_out/ncu_push2.log 2>&1
tail -3 gpurun_out/ncu_push2.log
