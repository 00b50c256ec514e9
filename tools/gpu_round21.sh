#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fp_fft_chroma_h32 -s 1 -c 1 -f -o gpurun_out/prof_k1_h32 python tools/profile_target.py 2 12 > gpurun_out/ncu_k1_h32.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_k1_h32.ncu-rep > gpurun_out/ncu_k1_h32_summary.txt 2>&1
cat gpurun_out/ncu_k1_h32_summary.txt; tail -3 gpurun_out/ncu_k1_h32.log
