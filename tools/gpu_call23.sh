#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
B2P_TRACE=1 timeout 900 python bench.py --cells 256 --steps 10 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench23.json 2> gpurun_out/bench23.err
grep -c trace gpurun_out/bench23.err; grep "trace" gpurun_out/bench23.err | tail -40
python -c "
import json;d=json.load(open('gpurun_out/bench23.json'));print(d['ms_per_step'],d['value'],d['roofline']['share_of_step'])"
