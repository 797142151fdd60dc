#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py --workload emf-wave --steps 10 --warmup 3 > gpurun_out/bench31_emf.json 2> gpurun_out/bench31_emf.err
echo rc=$?; tail -c 2500 gpurun_out/bench31_emf.json; tail -5 gpurun_out/bench31_emf.err
