#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/microbench.py --cells 256 --laps 5 --out gpurun_out/micro14.json "push_streams=1,sort_streams=1" 2>&1 | tail -5
python - <<'PY'
import json
for r in json.load(open('gpurun_out/micro14.json')):
    print(r['setting'], {k:round(v,3) for k,v in r['ms_per_lap_by_class'].items()})
    print(r['us_per_launch'])
PY
# one full ncu capture of the fused push (after a sort lap and 3 laps later)
B2P_OPTS=push_streams=1,sort_streams=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_push -s 16 -c 2 -o gpurun_out/push_r14a -f python tools/microbench.py --cells 128 --laps 1 "" > gpurun_out/ncu14a.log 2>&1
B2P_OPTS=push_streams=1,sort_streams=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_push -s 70 -c 2 -o gpurun_out/push_r14b -f python tools/microbench.py --cells 128 --laps 1 "" > gpurun_out/ncu14b.log 2>&1
tail -3 gpurun_out/ncu14a.log gpurun_out/ncu14b.log
ls -la gpurun_out
