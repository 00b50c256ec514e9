#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_fingerprint_gpu.py -m gpu -q -x --timeout=200 > gpurun_out/pytest_fp.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_fp.log
timeout 200 python tools/k1_variants.py 0 8 10 12 > gpurun_out/k1_variants.json 2> gpurun_out/k1_variants.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fp_fft_chroma_h32 -s 1 -c 1 -f -o gpurun_out/prof_k1_h32 python tools/profile_target.py 2 12 > gpurun_out/ncu_k1_h32.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_k1_h32.ncu-rep > gpurun_out/ncu_k1_h32_summary.txt 2>&1
tail -4 gpurun_out/pytest_fp.log; cat gpurun_out/k1_variants.json; tail -3 gpurun_out/k1_variants.err; cat gpurun_out/ncu_k1_h32_summary.txt
