#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 --timeout=300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1
NB200_BENCH_SPIN_STEPS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --cpu-baseline 0 > gpurun_out/bench_under_ncu.log 2>&1
tail -12 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err; tail -2 gpurun_out/smoke.log; tail -3 gpurun_out/bench_under_ncu.log | cut -c1-300; wc -l gpurun_out/launches.csv
