#!/bin/bash
# first GPU call of round 2: parity of the fp16 path, pipe microbenchmarks, attention exp2-offload sweep, bench
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/c1_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/c1_tests.log
timeout 120 tools/_build/ubench_pipes > gpurun_out/c1_ubench.txt 2>&1
for pp in 0 2 3 4; do
  BT_ATTN_POLY=$pp timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c1_bench_pp$pp.json 2> gpurun_out/c1_bench_pp$pp.err
done
timeout 600 python bench.py > gpurun_out/c1_bench_full.json 2> gpurun_out/c1_bench_full.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c1_bench_ref.json 2> gpurun_out/c1_bench_ref.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c1_launches.csv python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/c1_bench_under_ncu.log 2>&1
tail -5 gpurun_out/c1_tests.log
