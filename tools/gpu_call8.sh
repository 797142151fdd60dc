#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_golden.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log )
tail -n 3 gpurun_out/pytest_gpu.log
timeout 900 python tools/microbench.py --cells 128 --out gpurun_out/micro8.json \
  "fuse_deposit=1,push_minb=5,deposit_agg=1" "fuse_deposit=1,push_minb=4,deposit_agg=1" "fuse_deposit=1,push_minb=6,deposit_agg=1" "fuse_deposit=0,push_minb=8,deposit_minb=6" 2>&1 | tee gpurun_out/micro8.log
