#!/bin/bash
# bench.py on N GPUs (default 1), as the driver launches it, then the reference arm.
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  ( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
else
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
fi
echo "bench exit $?"; tail -c 1500 gpurun_out/bench_n$N.err; tail -c 6000 gpurun_out/bench_n$N.json
if [ -n "$WITH_REF" ]; then
  timeout 600 python bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
  echo "ref exit $?"; tail -c 1200 gpurun_out/bench_ref.json
fi
