#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L | head -3
( timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k multigpu > gpurun_out/pytest_mgpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_mgpu.log )
tail -n 6 gpurun_out/pytest_mgpu.log
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --cells 256 --steps 10 --warmup 3 --no-cpu-baseline --profile > gpurun_out/bench_2gpu_256.json 2> gpurun_out/bench_2gpu_256.err; echo "rc=$?" >> gpurun_out/bench_2gpu_256.err )
tail -n 25 gpurun_out/bench_2gpu_256.err | cut -c1-200
cut -c1-250 gpurun_out/bench_2gpu_256.json
( timeout 600 python bench.py --gpus 1 --cells 256 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_1gpu_256.json 2> gpurun_out/bench_1gpu_256.err )
cut -c1-250 gpurun_out/bench_1gpu_256.json
