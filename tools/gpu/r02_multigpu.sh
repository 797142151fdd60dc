#!/bin/bash
# multi-GPU parity driver (vs the CPU oracle) + a 256^3-per-GPU bench with and without the comm overlap
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
CELLS=${2:-256}
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/run_multigpu_parity.py > gpurun_out/r02_mgpu_parity_n$N.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/r02_mgpu_parity_n$N.log
for OV in 2 1 0; do
B2P_OPTS="comm_overlap=$OV" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$OV bench.py --gpus $N --cells $CELLS --steps 10 --warmup 5 --no-e2e --no-emf > gpurun_out/r02_bench_n${N}_${CELLS}_ov$OV.json 2> gpurun_out/r02_bench_n${N}_${CELLS}_ov$OV.err
echo "bench overlap=$OV rc=$?"; python - <<PY
import json
for l in open('gpurun_out/r02_bench_n${N}_${CELLS}_ov$OV.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['ms_per_step'], d['value']/1e9, d['config']['gpu_blocks'], d['checks'].get('multi_gpu_equals_single'), {k:v for k,v in d['roofline']['share_of_step'].items()})
PY
tail -2 gpurun_out/r02_bench_n${N}_${CELLS}_ov$OV.err
done
