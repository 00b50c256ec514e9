#!/bin/bash
# the scaling run on an 8-GPU box: N-way H2D ceiling, multi-GPU checks, bench at N = 8, 4, 2, 1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 300 python tools/h2d_bw.py --gpus 1,2,4,8 > gpurun_out/h2d_bw_n.json 2> gpurun_out/h2d_bw_n.err; echo "h2d exit $?"; tr -d '\n ' < gpurun_out/h2d_bw_n.json; echo
timeout 300 python tools/check_multi_gpu.py 8 > gpurun_out/multi_single8.log 2>&1; echo "single-process 8 exit $?"; tail -2 gpurun_out/multi_single8.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 tools/check_multi_gpu.py > gpurun_out/multi_procs4.log 2>&1; echo "torchrun 4 exit $?"; tail -3 gpurun_out/multi_procs4.log
for n in 8 4 2 1; do bash tools/gpu_bench.sh $n > gpurun_out/bench_n$n.log 2>&1; echo "bench $n: $(head -1 gpurun_out/bench_n$n.log)"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n$n.json').read().strip().splitlines()[-1])
    print($n, 'search', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'fp', round(d['fingerprint']['value']), 'season_e2e', round(d['season_e2e']['value']), 'resident', round(d['season_e2e']['resident_value']), d['collective_ms'], d['parity'])
except Exception as e:
    print('no line', e)
PY
done
