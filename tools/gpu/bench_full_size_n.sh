#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench34_n${N}_512.json 2> gpurun_out/bench34_n${N}_512.err
echo "bench rc=$?"; head -c 400 gpurun_out/bench34_n${N}_512.json; echo; tail -c 600 gpurun_out/bench34_n${N}_512.json; tail -5 gpurun_out/bench34_n${N}_512.err; nvidia-smi --query-gpu=memory.used,memory.total --format=csv
