#!/bin/bash
# the ncu artefacts of the round: launch list of a short bench run, full captures of K1 (default kernel, bench
# season shapes) and of K3 (adaptive kernel, 60-episode slice of configs[3])
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --fp-hours 40 --cpu-baseline 0 > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fp_fft_chroma -s 2 -c 1 -f -o gpurun_out/prof_k1_final python tools/profile_target.py 3 > gpurun_out/ncu_k1_final.log 2>&1
echo "ncu k1 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:match_fast_kernel -s 2 -c 1 -f -o gpurun_out/prof_k3_final python tools/profile_k3.py 3 > gpurun_out/ncu_k3_final.log 2>&1
echo "ncu k3 exit $?"
python tools/ncu_summary.py gpurun_out/prof_k1_final.ncu-rep > gpurun_out/ncu_k1_final.txt
python tools/ncu_summary.py gpurun_out/prof_k3_final.ncu-rep > gpurun_out/ncu_k3_final.txt; cat gpurun_out/ncu_k3_final.txt
ncu -i gpurun_out/prof_k3_final.ncu-rep --page source --csv --print-source sass > gpurun_out/ncu_k3_final_sass.csv 2>/dev/null
ncu -i gpurun_out/prof_k1_final.ncu-rep --page source --csv --print-source sass > gpurun_out/ncu_k1_final_sass.csv 2>/dev/null
