#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L
( timeout 600 python -m pytest tests -m gpu -x -q -k "multigpu" > gpurun_out/pytest_mgpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_mgpu.log )
tail -n 25 gpurun_out/pytest_mgpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --cells 256 --steps 10 --warmup 5 --no-e2e > gpurun_out/bench25_n2_256.json 2> gpurun_out/bench25_n2_256.err
echo "rc=$?"; tail -c 1500 gpurun_out/bench25_n2_256.json; tail -5 gpurun_out/bench25_n2_256.err
timeout 900 python bench.py --gpus 1 --cells 256 --steps 10 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench25_n1_256.json 2> gpurun_out/bench25_n1_256.err
tail -c 600 gpurun_out/bench25_n1_256.json
