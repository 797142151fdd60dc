#!/bin/bash
# compile-time kernel variants built as runko_b200/libb200pic_<tag>.so: per-lap push times of each, parity subset for the non-default ones
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for lib in runko_b200/libb200pic.so runko_b200/libb200pic_*.so; do
  [ -f "$lib" ] || continue
  tag=$(basename $lib .so | sed 's/libb200pic//')
  echo "=== $lib"
  B2P_LIB=$PWD/$lib timeout 600 python tools/microbench.py --cells 256 --laps 10 --out gpurun_out/micro_var$tag.json "push_streams=1,sort_streams=1" 2>&1 | grep -v "^ *per lap" | tail -2
  python - <<PY
import json
for r in json.load(open('gpurun_out/micro_var$tag.json')):
    print(r['setting'], round(r['ms_per_lap'],3), 'push', r['us_per_launch'].get('push'), [ (q['lap_mod5'], q['push_us']) for q in r['per_lap']])
PY
  if [ -n "$tag" ]; then
    ( B2P_LIB=$PWD/$lib timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_edge_cases_gpu.py tests/test_shock_gpu.py tests/test_golden.py -m gpu -x -q > gpurun_out/pytest_var$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_var$tag.log )
    tail -n 2 gpurun_out/pytest_var$tag.log
  fi
done
