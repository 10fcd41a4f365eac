#!/bin/bash
# ncu --set full captures of the main-layer GEMMs and of the attention kernel (64 x 30 s clips, one 128-chunk wave)
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 14 -c 5 -o gpurun_out/c2_gemm_full python tools/prof_step.py 64 1 > gpurun_out/c2_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc64 -s 3 -c 1 -o gpurun_out/c2_attn_full python tools/prof_step.py 64 1 > gpurun_out/c2_attn.log 2>&1
ncu -i gpurun_out/c2_gemm_full.ncu-rep --page raw --csv > gpurun_out/c2_gemm_full_raw.csv 2>/dev/null
ncu -i gpurun_out/c2_attn_full.ncu-rep --page raw --csv > gpurun_out/c2_attn_full_raw.csv 2>/dev/null
BT_ATTN_FREQ_SIMT=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "stage_parity_h16" > gpurun_out/c2_freqsimt.log 2>&1
tail -3 gpurun_out/c2_freqsimt.log
ls -la gpurun_out/c2_*
