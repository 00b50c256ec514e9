#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_capi.py -m gpu -q -x --timeout=200 > gpurun_out/pytest_capi.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_capi.log
tail -30 gpurun_out/pytest_capi.log
