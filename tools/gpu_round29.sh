#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_match_gpu.py tests/test_properties.py tests/test_vote_gpu.py -m gpu -q -x --timeout=300 > gpurun_out/pytest_match.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_match.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -6 gpurun_out/pytest_match.log; cat gpurun_out/bench_n1.json | cut -c1-1400; tail -3 gpurun_out/bench_n1.err
