#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log )
tail -n 3 gpurun_out/pytest_gpu.log
timeout 900 python tools/microbench.py --cells 256 --laps 5 --out gpurun_out/micro10.json \
  "push_streams=1,filter_chunk=14" "push_streams=2" "push_streams=4" "push_streams=2,filter_chunk=7" "filter_chunk=24" "filter_chunk=35" 2>&1 | tee gpurun_out/micro10.log
