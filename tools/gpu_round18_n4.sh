#!/bin/bash
mkdir -p gpurun_out
NB200_BENCH_WATCHDOG_S=90 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
cat gpurun_out/bench_n4.json | cut -c1-1200; grep -v "^\[W\|^W1\|^\*\*\*\|OMP_NUM" gpurun_out/bench_n4.err | tail -20
