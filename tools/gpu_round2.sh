#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
./tools/_bin/pipe_peak > gpurun_out/pipe_peak_warm.jsonl 2>&1
NB200_BENCH_SPIN_S=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --cpu-baseline 0 > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fp_fft_chroma -s 1 -c 1 -f -o gpurun_out/prof_k1 python tools/profile_target.py 2 > gpurun_out/ncu_k1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:match_kernel -s 1 -c 1 -f -o gpurun_out/prof_k3 python tools/profile_target.py 2 > gpurun_out/ncu_k3.log 2>&1
tail -3 gpurun_out/smoke.log; tail -25 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err; cat gpurun_out/bench_ref.json; cat gpurun_out/pipe_peak_warm.jsonl; ls -la gpurun_out
