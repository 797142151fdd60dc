#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out /tmp/ncu
export B2P_OPTS=push_streams=1,sort_streams=1
R=/tmp/ncu
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_push -s 16 -c 1 -o $R/push_a -f python tools/microbench.py --cells 128 --laps 1 "" > gpurun_out/ncu18a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_push -s 70 -c 1 -o $R/push_b -f python tools/microbench.py --cells 128 --laps 1 "" > gpurun_out/ncu18b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_sort|k_gather|DeviceScan|k_max|Radix" -s 80 -c 8 -o $R/sort -f python tools/microbench.py --cells 128 --laps 1 "" > gpurun_out/ncu18c.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_filter|k_halo|k_edge|k_nodal|k_push_b|k_push_e|k_J_ex|k_zero|k_collect|k_append|k_gather_out" -s 200 -c 14 -o $R/fields -f python tools/microbench.py --cells 128 --laps 1 "" > gpurun_out/ncu18d.log 2>&1
for f in push_a push_b sort fields; do
  ncu -i $R/$f.ncu-rep --page raw --csv > gpurun_out/r18_${f}_raw.csv 2>/dev/null
done
ncu -i $R/push_a.ncu-rep --page source --csv > gpurun_out/r18_push_a_source.csv 2>/dev/null
ncu -i $R/push_b.ncu-rep --page source --csv > gpurun_out/r18_push_b_source.csv 2>/dev/null
ncu -i $R/push_a.ncu-rep --page details > gpurun_out/r18_push_a_details.txt 2>/dev/null
ncu -i $R/push_b.ncu-rep --page details > gpurun_out/r18_push_b_details.txt 2>/dev/null
ncu -i $R/sort.ncu-rep --page details > gpurun_out/r18_sort_details.txt 2>/dev/null
cp $R/push_a.ncu-rep gpurun_out/r18_push_a.ncu-rep
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3500 --csv --log-file gpurun_out/launches18.csv python bench.py --cells 256 --steps 5 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench18_under_ncu.log 2>&1
gzip -f gpurun_out/launches18.csv
du -sh gpurun_out; ls -la gpurun_out
