#!/bin/bash
# GPU parity suite, then the kernel-class microbenchmark: one-slot-per-thread k_push vs k_push2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log )
tail -n 15 gpurun_out/pytest_gpu.log
timeout 900 python tools/microbench.py --cells 256 --laps 5 --out gpurun_out/r02_micro_a.json "push_kernel=1,push_group=8,push_streams=4" "push_kernel=2,push_group=32,push_streams=2" "push_kernel=2,push_group=32,push_streams=1" 2>&1 | tail -8
