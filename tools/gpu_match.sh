#!/bin/bash
# match stage on the GPU: parity tests, adversarial seasons, the search leg of the bench, sanitizer
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_match_gpu.py tests/test_vote_gpu.py tests/test_properties.py tests/test_full_size_gpu.py tests/test_multi_gpu.py -m gpu -q -x --timeout=600 > gpurun_out/pytest_match.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_match.log
timeout 300 python tools/adversarial_match.py --oracle > gpurun_out/adversarial_match.json 2> gpurun_out/adversarial_match.err
echo "adversarial exit $?"; tail -3 gpurun_out/adversarial_match.err; cat gpurun_out/adversarial_match.json
timeout 300 python bench.py --legs search --cpu-baseline 0 > gpurun_out/bench_search.json 2> gpurun_out/bench_search.err
echo "bench exit $?"; tail -3 gpurun_out/bench_search.err; python -c "
import json; d=json.load(open('gpurun_out/bench_search.json')); print({k:d[k] for k in ('value','ms_per_step','kernel_ms_per_step_rank0','collective_ms')}); print(d['e2e']); print(d['roofline_popc']['frac'], d['roofline_popc']['kernel_ms'])"
if [ -n "$SANITIZE" ]; then
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize_target.py > gpurun_out/sanitize_$tool.log 2>&1; echo "$tool exit $?"; tail -3 gpurun_out/sanitize_$tool.log
done
fi
