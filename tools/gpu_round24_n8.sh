#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
NB200_BENCH_WATCHDOG_S=120 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
cat gpurun_out/bench_n$N.json | cut -c1-1500; grep -v "^\[W\|^W1\|^\*\*\*\|OMP_NUM" gpurun_out/bench_n$N.err | tail -20
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; head -30 gpurun_out/topo.txt
