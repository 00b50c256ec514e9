#!/bin/bash
# 8-GPU box, short form: single-process 8-rank job check, then bench at N = 8 and 4 (N = 2 and 1 run on a 2-GPU box)
mkdir -p gpurun_out
timeout 300 python tools/check_multi_gpu.py 8 > gpurun_out/multi_single8.log 2>&1; echo "single-process 8 exit $?"; tail -2 gpurun_out/multi_single8.log
for n in 8 4; do bash tools/gpu_bench.sh $n > gpurun_out/bench_n$n.log 2>&1; echo "bench $n: $(head -1 gpurun_out/bench_n$n.log)"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n$n.json').read().strip().splitlines()[-1])
    print($n, 'search', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'fp', round(d['fingerprint']['value']), 'season_e2e', round(d['season_e2e']['value']), 'resident', round(d['season_e2e']['resident_value']), d['collective_ms'], d['parity'])
except Exception as e:
    print('no line', e)
PY
done
