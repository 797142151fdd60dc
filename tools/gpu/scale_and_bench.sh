#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_scale_properties_gpu.py tests/test_abi.py -m gpu -x -q > gpurun_out/pytest_scale.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_scale.log )
tail -n 12 gpurun_out/pytest_scale.log
timeout 1500 python bench.py --steps 10 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print(d['ms_per_step'], d['value']/1e9, d['e2e']['value']/1e9, d['clocks'])
print(d['checks'])
PY
tail -3 gpurun_out/bench_n1.err
