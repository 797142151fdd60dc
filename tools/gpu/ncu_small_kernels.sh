#!/bin/bash
# ncu --set full of the field-phase and migration kernels (the sort kernels fill collect_evidence.sh's capture)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out /tmp/ncu
export B2P_OPTS=push_streams=1,sort_streams=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_filter|k_halo_fill|k_J_exchange|k_append|k_gather_outgoing|k_edge_gather|k_nodal|k_push_b|k_push_e|k_collect" -s 40 -c 16 -o /tmp/ncu/small -f python tools/microbench.py --cells 256 --laps 1 "" > gpurun_out/ncu_small.log 2>&1
tail -3 gpurun_out/ncu_small.log
ncu -i /tmp/ncu/small.ncu-rep --page raw --csv > gpurun_out/r31_small_raw.csv 2>/dev/null
ncu -i /tmp/ncu/small.ncu-rep --page source --csv -k regex:k_filter > gpurun_out/r31_filter_source.csv 2>/dev/null
ls -la gpurun_out/r31*
