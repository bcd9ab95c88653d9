#!/bin/bash
# One GPU-box session: parity tests, the bench line, a launch list and one full ncu capture per dominant kernel.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r02a'
tag=${1:-r02}
mkdir -p gpurun_out
rm -f gpurun_out/parity_full.jsonl
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu_$tag.txt 2>&1
python -m pytest tests -m gpu -x -q -rP > gpurun_out/tests_$tag.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests_$tag.log
tail -3 gpurun_out/tests_$tag.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench exit $?"
python bench.py --steps 20 --warmup 5 --workload config3 --no-cpu-baseline > gpurun_out/bench_c3_$tag.json 2>> gpurun_out/bench_$tag.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/launch_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sgp4_grid_kernel --launch-skip 4 -c 1 -f -o gpurun_out/prof_k1_$tag \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k1_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sdp4_grid_kernel --launch-skip 4 -c 1 -f -o gpurun_out/prof_k2_$tag \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload config3 > gpurun_out/ncu_k2_$tag.log 2>&1
ls -la gpurun_out | tail -12
