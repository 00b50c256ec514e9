#!/bin/bash
# full ncu capture of the adaptive match kernel on a configs[3] slice + per-line source/SASS pages
mkdir -p gpurun_out
timeout 300 python tools/profile_k3.py 3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:match_fast_kernel -s 2 -c 1 -f -o gpurun_out/prof_k3 python tools/profile_k3.py 3 > gpurun_out/ncu_k3.log 2>&1
echo "ncu k3 exit $?"
python tools/ncu_summary.py gpurun_out/prof_k3.ncu-rep > gpurun_out/ncu_k3_summary.txt; cat gpurun_out/ncu_k3_summary.txt
ncu -i gpurun_out/prof_k3.ncu-rep --page source --csv > gpurun_out/ncu_k3_source.csv 2>/dev/null
ncu -i gpurun_out/prof_k3.ncu-rep --page source --csv --print-source sass > gpurun_out/ncu_k3_sass.csv 2>/dev/null
wc -l gpurun_out/ncu_k3_source.csv gpurun_out/ncu_k3_sass.csv
