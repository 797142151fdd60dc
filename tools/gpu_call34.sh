#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_configs_gpu.py -m gpu -x -q -s > gpurun_out/pytest_cfg.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_cfg.log )
tail -n 30 gpurun_out/pytest_cfg.log
