#!/bin/bash
# the default bench line at full size (512^3 x 32 ppc per GPU) on N GPUs, launched the way the driver does
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
TAG=${2:-c}
shift 2
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 10 --warmup 3 --profile "$@" > gpurun_out/r02_bench_${TAG}_n${N}_512.json 2> gpurun_out/r02_bench_${TAG}_n${N}_512.err
echo "bench rc=$?"
python - <<PY
import json
for l in open('gpurun_out/r02_bench_${TAG}_n${N}_512.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['ms_per_step'], d['value']/1e9, d.get('e2e',{}).get('value',0)/1e9, d['config']['gpu_blocks'], d['checks'].get('multi_gpu_equals_single'), d['roofline']['share_of_step'], d['clocks'])
PY
grep -v "^\[W\|^W0\|^\*\*\*" gpurun_out/r02_bench_${TAG}_n${N}_512.err | tail -40
nvidia-smi --query-gpu=memory.used,memory.total --format=csv | head -3
