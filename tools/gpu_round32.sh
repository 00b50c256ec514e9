#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fp_fft_chroma_h32 -s 1 -c 1 -f -o gpurun_out/prof_k1_h32 python tools/profile_target.py 2 12 > gpurun_out/ncu_k1_h32.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_k1_h32.ncu-rep > gpurun_out/ncu_k1_h32_summary.txt 2>&1
ncu -i gpurun_out/prof_k1_h32.ncu-rep --page raw --csv --metrics smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_lsu.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active 2>/dev/null | tail -2 > gpurun_out/ncu_k1_fp_ops.csv
cat gpurun_out/ncu_k1_h32_summary.txt; cat gpurun_out/ncu_k1_fp_ops.csv | cut -c1-1500
