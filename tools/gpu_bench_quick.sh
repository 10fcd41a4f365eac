#!/bin/bash
# usage: bash tools/gpu_bench_quick.sh <tag> [ENV=VAL ...]   -> gpurun_out/q_<tag>.json
tag=$1; shift
mkdir -p gpurun_out
env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/q_$tag.json 2> gpurun_out/q_$tag.err
tail -2 gpurun_out/q_$tag.err
