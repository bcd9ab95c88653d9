#!/bin/bash
# quick GPU session: selected tests, then the bench line + fp64 probe (+ optional extra tools)
#   gpurun -- 'bash tools/gpu_quick.sh r02c "-k multi_device" --steps 200'
tag=${1:-r02}; sel=${2:-}; shift; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -rP $sel > gpurun_out/tests_$tag.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/tests_$tag.log
bash tools/gpu_bench.sh $tag "$@"
if [ -n "$AZ_EXTRA" ]; then for t in $AZ_EXTRA; do python tools/$t.py > gpurun_out/${t}_$tag.jsonl 2> gpurun_out/${t}_$tag.err; echo "$t exit $?"; tail -20 gpurun_out/${t}_$tag.jsonl; done; fi
