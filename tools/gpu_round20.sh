#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_fingerprint_gpu.py -m gpu -q -x --timeout=200 > gpurun_out/pytest_fp.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_fp.log
timeout 200 python tools/k1_variants.py > gpurun_out/k1_variants.json 2> gpurun_out/k1_variants.err
timeout 400 python tools/quick_bench.py > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err
tail -8 gpurun_out/pytest_fp.log; cat gpurun_out/k1_variants.json; tail -3 gpurun_out/k1_variants.err; cat gpurun_out/quick_bench.json; tail -3 gpurun_out/quick_bench.err
