#!/bin/bash
# the ncu artefacts of the round: launch list of a short bench run, full captures of K1 (default kernel, bench
# season shapes) and of K3 (adaptive kernel, configs[3] slice)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --fp-hours 40 --cpu-baseline 0 > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fp_fft_chroma -s 2 -c 1 -f -o gpurun_out/prof_k1_final python tools/profile_target.py 3 > gpurun_out/ncu_k1_final.log 2>&1
echo "ncu k1 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:match_fast_kernel -s 2 -c 1 -f -o gpurun_out/prof_k3_final python tools/profile_target.py 3 > gpurun_out/ncu_k3_final.log 2>&1
echo "ncu k3 exit $?"
