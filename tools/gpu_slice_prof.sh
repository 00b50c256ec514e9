#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:match_fast_kernel -s 3 -c 1 -f -o gpurun_out/prof_k3_slice python tools/slice_match.py 8 > gpurun_out/ncu_k3_slice.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/prof_k3_slice.ncu-rep --page raw --csv > gpurun_out/ncu_k3_slice_raw.csv
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/ncu_k3_slice_raw.csv')))
hdr,units=rows[0],rows[1]
d=dict(zip(hdr,rows[2]))
for k in hdr:
    if any(x in k for x in ("gpu__time_duration.sum","sm__cycles_active.avg","sm__cycles_active.max","sm__cycles_active.min","sm__cycles_elapsed.avg ","sm__cycles_elapsed.max","launch__waves","launch__grid_size","sm__warps_active.avg.pct","smsp__cycles_active.avg","smsp__cycles_active.min","smsp__cycles_active.max","gpc__cycles_elapsed.max","sm__inst_executed_pipe_xu.avg.pct","smsp__inst_executed.sum ")):
        print(k, d[k], units[hdr.index(k)])
PY
