#!/bin/bash
# BASELINE config 5 (ragged clips) and the headline config 2 on 8 GPUs of one box; "all" adds config 3 (10 000 WAV files)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus 8 --config 5 --clips 1024 > gpurun_out/cfg5_8gpu.json 2> gpurun_out/cfg5_8gpu.err; tail -2 gpurun_out/cfg5_8gpu.err
timeout 600 $TR bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/cfg2_8gpu.json 2> gpurun_out/cfg2_8gpu.err; tail -2 gpurun_out/cfg2_8gpu.err
if [ "$1" = all ]; then
  timeout 900 $TR bench.py --gpus 8 --config 3 > gpurun_out/cfg3_8gpu.json 2> gpurun_out/cfg3_8gpu.err; tail -2 gpurun_out/cfg3_8gpu.err
fi
