#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_dist_gpu.py -m gpu -q -x --timeout=400 > gpurun_out/pytest_dist.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_dist.log
tail -15 gpurun_out/pytest_dist.log
