#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:match_fast -s 1 -c 1 -f -o gpurun_out/prof_k3_real python tools/profile_match_real.py 2 > gpurun_out/ncu_k3_real.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_k3_real.ncu-rep > gpurun_out/ncu_k3_real_summary.txt 2>&1
cat gpurun_out/ncu_k3_real_summary.txt; tail -3 gpurun_out/ncu_k3_real.log
