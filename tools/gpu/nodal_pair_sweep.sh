#!/bin/bash
# nodal staging layouts (B2P_NODAL_BPAIR = 0 / 1 / 2 builds of the library): parity subset + per-lap push times
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for v in "" _bpair _apair; do
  lib=$PWD/runko_b200/libb200pic$v.so
  [ -f "$lib" ] || continue
  echo "=== $lib"
  B2P_LIB=$lib timeout 600 python tools/microbench.py --cells 256 --laps 5 --out gpurun_out/micro_nodal$v.json "push_streams=1,sort_streams=1" "push_minb=5" 2>&1 | grep -v "^ *per lap" | tail -3
  B2P_LIB=$lib python - <<PY
import json
for r in json.load(open('gpurun_out/micro_nodal$v.json')):
    print(r['setting'], round(r['ms_per_lap'],3), 'nodal', r['us_per_launch'].get('nodal_means'), [ (q['lap_mod5'], q['push_us']) for q in r['per_lap']])
PY
  if [ -n "$v" ]; then
    ( B2P_LIB=$lib timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_edge_cases_gpu.py tests/test_shock_gpu.py tests/test_golden.py -m gpu -x -q > gpurun_out/pytest_nodal$v.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_nodal$v.log )
    tail -n 3 gpurun_out/pytest_nodal$v.log
  fi
done
