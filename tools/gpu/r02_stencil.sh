#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_shock_gpu.py tests/test_golden.py -m gpu -x -q -k "stencil or shock" 2>&1 | tail -3
for IT in 1 2 4; do
B2P_OPTS="stencil_it=$IT" timeout 600 python bench.py --workload shock --steps 5 --warmup 3 --profile --no-cpu-baseline --no-e2e > gpurun_out/st_$IT.json 2> gpurun_out/st_$IT.err
echo "IT=$IT $(grep push_b gpurun_out/st_$IT.err) $(python -c "import json; print(json.load(open('gpurun_out/st_$IT.json'))['ms_per_step'])")"
done
