#!/bin/bash
# GPU parity suite + kernel-class microbenchmark (256^3 cells, 64 tiles of 64^3, 2 x 16 ppc)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log )
tail -n 5 gpurun_out/pytest_gpu.log
timeout 900 python tools/microbench.py --cells 256 --laps 5 --out gpurun_out/r02_micro.json "$@" 2>&1 | tail -8
