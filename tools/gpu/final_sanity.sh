#!/bin/bash
# what the driver runs at round end, shortened: smoke(), a small bench line of each arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --cells 256 --steps 5 --warmup 3 > gpurun_out/bench_sanity.json 2> gpurun_out/bench_sanity.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_sanity.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','timed_laps','sort_laps_timed','gpu_launches')}, d['e2e']['value'], d['cpu_baseline']['value'], d['roofline']['frac'], d['checks'])
PY
