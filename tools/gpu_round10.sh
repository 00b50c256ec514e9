#!/bin/bash
mkdir -p gpurun_out; NG=${NG:-2}


timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $NG --steps 30 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "exit $?" >> gpurun_out/bench_n2.err
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cat gpurun_out/bench_n2.json; grep -v "^\*\|OMP" gpurun_out/bench_n2.err | tail -12; cat gpurun_out/bench_n1.json
