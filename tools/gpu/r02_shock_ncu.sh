#!/bin/bash
# ncu --set full of the two field kernels that dominate the shock lap (extended-stencil push_b, separable filter)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out /tmp/ncu
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_push_b_stencil|k_filter_binomial2" -s 4 -c 2 -o /tmp/ncu/shock -f python bench.py --workload shock --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_ncu_shock.log 2>&1
ncu -i /tmp/ncu/shock.ncu-rep --page raw --csv > gpurun_out/r02_shock_raw.csv 2>/dev/null
ncu -i /tmp/ncu/shock.ncu-rep --page source --csv -k regex:k_push_b_stencil > gpurun_out/r02_stencil_source.csv 2>/dev/null
tail -3 gpurun_out/r02_ncu_shock.log; ls -la gpurun_out/r02_shock_raw.csv gpurun_out/r02_stencil_source.csv
