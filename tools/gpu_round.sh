#!/bin/bash
# One GPU-box visit: microbenchmark, parity tests, quick timings.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
./tools/_bin/pipe_peak > gpurun_out/pipe_peak.jsonl 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python tools/quick_bench.py > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err
tail -5 gpurun_out/pytest_gpu.log
cat gpurun_out/pipe_peak.jsonl
cat gpurun_out/quick_bench.json
tail -3 gpurun_out/quick_bench.err
