#!/bin/bash
# multi-GPU jobs behind the C ABI on N GPUs (default 2): single-process and one-process-per-GPU shapes,
# then the whole GPU test suite.
N=${1:-2}
mkdir -p gpurun_out
timeout 150 python tools/check_multi_gpu.py $N > gpurun_out/multi_single.log 2>&1
echo "single-process exit $?"; tail -4 gpurun_out/multi_single.log
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 tools/check_multi_gpu.py > gpurun_out/multi_procs.log 2>&1
echo "torchrun exit $?"; tail -6 gpurun_out/multi_procs.log
if [ -z "$SKIP_PYTEST" ]; then
timeout 1200 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu.log
fi
timeout 200 python tools/k1_variants.py 16 17 > gpurun_out/k1_variants.json 2> gpurun_out/k1_variants.err; cat gpurun_out/k1_variants.json
