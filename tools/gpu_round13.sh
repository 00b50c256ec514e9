#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=3 --timeout=300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python tools/quick_bench.py > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:match_fast -s 1 -c 1 -f -o gpurun_out/prof_k3s python tools/profile_target.py 2 > gpurun_out/ncu_k3.log 2>&1
tail -8 gpurun_out/pytest_gpu.log; cat gpurun_out/quick_bench.json; tail -3 gpurun_out/quick_bench.err; cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
