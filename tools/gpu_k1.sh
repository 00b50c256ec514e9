#!/bin/bash
# K1 variants on the GPU: microbenchmarks, timing + equality, parity tests, ncu capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/k1_smi.txt
[ -n "$K1_SKIP_MIX" ] || timeout 120 tools/_bin/k1_mix_peak ${K1_MIX_ARGS} > gpurun_out/k1_mix_peak.jsonl 2> gpurun_out/k1_mix_peak.err
[ -x tools/_bin/tmem_bw ] && timeout 120 tools/_bin/tmem_bw > gpurun_out/tmem_bw.jsonl 2>&1
cat gpurun_out/tmem_bw.jsonl
timeout 300 python tools/k1_variants.py ${K1_VARIANTS:-17 18} > gpurun_out/k1_variants.json 2> gpurun_out/k1_variants.err
echo "variants exit $?"; tail -3 gpurun_out/k1_variants.err; cat gpurun_out/k1_variants.json
NB200_K1_VARIANT=${K1_TEST_VARIANT:-18} timeout 600 python -m pytest tests/test_fingerprint_gpu.py tests/test_capi.py -m gpu -q -x --timeout=300 > gpurun_out/pytest_fp_variant.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_fp_variant.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fp_fft_chroma -s 2 -c 1 -f -o gpurun_out/prof_k1_tm python tools/profile_target.py 3 ${K1_TEST_VARIANT:-18} > gpurun_out/ncu_k1_tm.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_k1_tm.log
[ -n "$K1_SKIP_MIX" ] || cat gpurun_out/k1_mix_peak.jsonl
