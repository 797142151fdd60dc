#!/bin/bash
# bench.py --workload beam / shock / emf-wave: $1 = "small" (quick functional pass) or "full" (BASELINE sizes)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
if [ "$1" = "small" ]; then BC="--cells 256"; SC="--cells 512"; EC="--cells 256 --tile 64"; TAG=small; else BC=""; SC=""; EC=""; TAG=full; fi
for W in beam shock emf-wave; do
  case $W in beam) X="$BC";; shock) X="$SC";; *) X="$EC";; esac
  timeout 1500 python bench.py --workload $W $X --steps 10 --warmup 5 --profile > gpurun_out/r02_bench_${W}_$TAG.json 2> gpurun_out/r02_bench_${W}_$TAG.err
  echo "$W rc=$?"; tail -12 gpurun_out/r02_bench_${W}_$TAG.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_bench_${W}_$TAG.json'))
    print('$W', d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['step']['frac_of_peak'], d.get('e2e',{}).get('value'), d.get('cpu_baseline',{}).get('value'), d.get('checks'))
except Exception as e: print('no json', e)
PY
done
