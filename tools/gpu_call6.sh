#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fused_ff_kernel|fused_qkv_kernel" -c 4 -o gpurun_out/c6_fused python tools/prof_step.py 64 1 > gpurun_out/c6.log 2>&1
ncu -i gpurun_out/c6_fused.ncu-rep --page raw --csv > gpurun_out/c6_fused_raw.csv 2>/dev/null
for i in 0 1 2 3; do ncu -i gpurun_out/c6_fused.ncu-rep --page source --csv --launch-skip $i --launch-count 1 > gpurun_out/c6_src$i.csv 2>/dev/null; done
ls -la gpurun_out/c6_*
