#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 15 -c 4 -o gpurun_out/c5_gemm_full python tools/prof_step.py 64 1 > gpurun_out/c5_gemm.log 2>&1
ncu -i gpurun_out/c5_gemm_full.ncu-rep --page raw --csv > gpurun_out/c5_gemm_full_raw.csv 2>/dev/null
ncu -i gpurun_out/c5_gemm_full.ncu-rep --page source --csv --launch-skip 2 --launch-count 1 > gpurun_out/c5_ff1_source.csv 2>/dev/null
ls -la gpurun_out/c5_*
