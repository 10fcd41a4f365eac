#!/bin/bash
mkdir -p gpurun_out
for a in 1 9 16 32 48 24; do
  BT_GEMM_ABL=$a timeout 200 python bench.py --steps 6 --warmup 2 --no-cpu-baseline > gpurun_out/c4_abl$a.json 2> gpurun_out/c4_abl$a.err
done
ls gpurun_out/c4_*
