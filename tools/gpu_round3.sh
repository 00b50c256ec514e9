#!/bin/bash
mkdir -p gpurun_out
python tools/h2d_bw.py > gpurun_out/h2d_bw.json 2>&1
./tools/_bin/pipe_peak > gpurun_out/pipe_peak_v2.jsonl 2>&1
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 1500 python -m pytest tests -m gpu -v --durations=0 --timeout=400 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
cat gpurun_out/h2d_bw.json; cat gpurun_out/pipe_peak_v2.jsonl; cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err; tail -70 gpurun_out/pytest_gpu.log
