#!/bin/bash
# Round-2 evidence: launch list of the bench command, ncu --set full of the attention kernel and of the main-layer GEMMs
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc48 -s 2 -c 2 -o gpurun_out/r2_attn_full python tools/prof_step.py 64 1 > gpurun_out/r2_attn.log 2>&1
ncu -i gpurun_out/r2_attn_full.ncu-rep --page raw --csv > gpurun_out/r2_attn_full_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 14 -c 5 -o gpurun_out/r2_gemm_full python tools/prof_step.py 64 1 > gpurun_out/r2_gemm.log 2>&1
ncu -i gpurun_out/r2_gemm_full.ncu-rep --page raw --csv > gpurun_out/r2_gemm_full_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:"fused_ff_kernel|fused_qkv_kernel|attn_freq|logmel|norm_kernel" -c 12 -o gpurun_out/r2_front_full python tools/prof_step.py 64 1 > gpurun_out/r2_front.log 2>&1
ncu -i gpurun_out/r2_front_full.ncu-rep --page raw --csv > gpurun_out/r2_front_full_raw.csv 2>/dev/null
ls -la gpurun_out/r2_*
