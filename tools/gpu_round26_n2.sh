#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dist_gpu.py tests/test_vote_gpu.py -m gpu -q -x --timeout=600 > gpurun_out/pytest_dist.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_dist.log
NB200_BENCH_WATCHDOG_S=90 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -6 gpurun_out/pytest_dist.log; cat gpurun_out/bench_n2.json | cut -c1-900; grep -v "^\[W\|^W1\|^\*\*\*\|OMP_NUM" gpurun_out/bench_n2.err | tail
