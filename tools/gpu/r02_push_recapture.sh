#!/bin/bash
# re-capture of the push alone (`ncu -k k_push`: the exact kernel name — a regex also matches k_push_e_fdtd2) + the headline bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=${1:-e}
mkdir -p gpurun_out /tmp/ncu
export B2P_OPTS=push_streams=1,sort_streams=0
timeout 900 ncu --set full --clock-control none --import-source on -k k_push -s 7 -c 1 -o /tmp/ncu/push -f python tools/microbench.py --cells 128 --laps 1 "" > gpurun_out/r02_ncu_${TAG}_push.log 2>&1
ncu -i /tmp/ncu/push.ncu-rep --page raw --csv > gpurun_out/r02_${TAG}_push_raw.csv 2>/dev/null
ncu -i /tmp/ncu/push.ncu-rep --page source --csv > gpurun_out/r02_${TAG}_push_source.csv 2>/dev/null
unset B2P_OPTS
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/r02_${TAG}_push_raw.csv')))
h=rows[0]; r=rows[2]
for k in ('Kernel Name','Grid Size','dram__bytes_read.sum','dram__bytes_write.sum','gpu__time_duration.sum'):
    print(k, r[h.index(k)], rows[1][h.index(k)])
PY
