#!/bin/bash
# everything the driver runs at round end on one GPU: the GPU test suite, smoke(), both bench arms
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 300 python tools/h2d_bw.py --gpus 1 > gpurun_out/h2d_bw_1.json 2>&1; cat gpurun_out/h2d_bw_1.json | tr -d '\n'; echo
WITH_REF=${WITH_REF:-} bash tools/gpu_bench.sh 1
