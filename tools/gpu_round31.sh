#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_fingerprint_gpu.py tests/test_match_gpu.py -m gpu -q -x --timeout=200 > gpurun_out/pytest_fp.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_fp.log
timeout 200 python tools/k1_variants.py 1 8 12 > gpurun_out/k1_variants.json 2> gpurun_out/k1_variants.err
tail -4 gpurun_out/pytest_fp.log; cat gpurun_out/k1_variants.json; tail -3 gpurun_out/k1_variants.err
