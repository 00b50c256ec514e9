#!/bin/bash
# round-end validation on one GPU: microbenchmark of the current K1 mix, the whole GPU suite, smoke, both bench arms, ncu artefacts
mkdir -p gpurun_out
timeout 120 tools/_bin/k1_mix_peak > gpurun_out/k1_mix_peak_v4.jsonl 2>&1; cat gpurun_out/k1_mix_peak_v4.jsonl; cp gpurun_out/k1_mix_peak_v4.jsonl profiles/r02_k1_fp_mix_peak_v4.jsonl
WITH_REF=1 bash tools/gpu_all.sh
bash tools/gpu_profile.sh
