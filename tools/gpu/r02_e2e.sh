#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 5 --profile --no-cpu-baseline --no-emf "$@" > gpurun_out/r02_bench_e2e.json 2> gpurun_out/r02_bench_e2e.err
grep -v "^  [a-z_]* *[0-9.]* ms/step" gpurun_out/r02_bench_e2e.err | tail -8
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_e2e.json').read().strip().splitlines()[-1])
e=d.get('e2e',{}).get('value',0.0); print(d['ms_per_step'], d['value']/1e9, e/1e9, e/d['value'], d['host_enqueue_ms_per_step'], d['clocks']['sm_mhz'])
PY
