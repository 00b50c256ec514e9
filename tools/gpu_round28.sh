#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_properties.py tests/test_fingerprint_gpu.py -m gpu -q -x --timeout=500 > gpurun_out/pytest_prop.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_prop.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -15 gpurun_out/pytest_prop.log; cat gpurun_out/bench_n1.json | cut -c1-1300; tail -3 gpurun_out/bench_n1.err
