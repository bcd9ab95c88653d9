#!/bin/bash
# ncu full captures of the two grid kernels + the bench line:  gpurun -- 'bash tools/gpu_prof.sh r02e'
tag=${1:-r02}
mkdir -p gpurun_out
python bench.py --steps 200 --warmup 10 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench exit $?"
ncu --set full --clock-control none --import-source on -k regex:sgp4_grid_kernel --launch-skip 4 -c 1 -f -o gpurun_out/prof_k1_$tag \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-subrecords > gpurun_out/ncu_k1_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sdp4_grid_kernel --launch-skip 4 -c 1 -f -o gpurun_out/prof_k2_$tag \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload config3 > gpurun_out/ncu_k2_$tag.log 2>&1
ls -la gpurun_out/*$tag*
