#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=5 --timeout=300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python tools/quick_bench.py > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fp_fft_chroma_g4 -s 1 -c 1 -f -o gpurun_out/prof_k1g4 python tools/profile_target.py 2 4 > gpurun_out/ncu_k1.log 2>&1
tail -12 gpurun_out/pytest_gpu.log; cat gpurun_out/quick_bench.json; tail -3 gpurun_out/quick_bench.err; cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
