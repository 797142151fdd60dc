#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-4}
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/run_multigpu_parity.py > gpurun_out/mgpu_parity_n$N.log 2>&1; echo "parity rc=$?"; tail -5 gpurun_out/mgpu_parity_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --cells 256 --steps 10 --warmup 5 --no-e2e > gpurun_out/bench29_n${N}_256.json 2> gpurun_out/bench29_n${N}_256.err
echo "bench rc=$?"; python - <<PY
import json
for l in open('gpurun_out/bench29_n${N}_256.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['ms_per_step'], d['value']/1e9, d['config']['gpu_blocks'])
PY
tail -3 gpurun_out/bench29_n${N}_256.err
