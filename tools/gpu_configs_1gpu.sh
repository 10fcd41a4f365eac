#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --config 3 --clips 1024 > gpurun_out/cfg3_1gpu.json 2> gpurun_out/cfg3_1gpu.err; tail -3 gpurun_out/cfg3_1gpu.err
timeout 600 python bench.py --config 4 --clips 512 > gpurun_out/cfg4_1gpu.json 2> gpurun_out/cfg4_1gpu.err; tail -3 gpurun_out/cfg4_1gpu.err
timeout 600 python bench.py --config 5 --clips 96 > gpurun_out/cfg5_1gpu.json 2> gpurun_out/cfg5_1gpu.err; tail -3 gpurun_out/cfg5_1gpu.err
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q -p no:cacheprovider -x > gpurun_out/cfg_e2e_tests.log 2>&1; tail -5 gpurun_out/cfg_e2e_tests.log
