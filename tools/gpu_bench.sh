#!/bin/bash
# bench line(s) + fp64 pipe probe on the GPU box:  gpurun -- 'bash tools/gpu_bench.sh r02b [extra bench args]'
tag=${1:-r02}; shift
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/fp64_probe tools/fp64_probe.cu && gpurun_out/fp64_probe > gpurun_out/fp64_probe_$tag.jsonl 2>&1
python bench.py "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench exit $?"
tail -c 600 gpurun_out/bench_$tag.err
cat gpurun_out/fp64_probe_$tag.jsonl
