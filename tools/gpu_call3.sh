#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -s -x > gpurun_out/c3_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/c3_tests.log
tail -5 gpurun_out/c3_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c3_bench_te1.json 2> gpurun_out/c3_bench_te1.err
BT_GEMM_TMA_EPI=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c3_bench_te0.json 2> gpurun_out/c3_bench_te0.err
tail -3 gpurun_out/c3_bench_te1.err
