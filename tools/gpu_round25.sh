#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 --timeout=300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -12 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_n1.json | cut -c1-1800; tail -3 gpurun_out/bench_n1.err
