#!/bin/bash
# Build library variants ON THE GPU BOX (same sources, extra -D flags) and time each:
#   gpurun -- 'bash tools/variant_sweep.sh r02e "AZ_K2_LANES=1 AZ_DEFAULT_K2_BLOCKS=5" "AZ_K2_LANES=2 AZ_DEFAULT_K2_BLOCKS=4"'
tag=$1; shift
mkdir -p gpurun_out
out=gpurun_out/variant_sweep_$tag.jsonl; : > $out
AZ_TAG=default python tools/k2_variant_timing.py >> $out 2>> gpurun_out/variant_sweep_$tag.err
i=0
for v in "$@"; do
  i=$((i+1)); lib=/tmp/libaz_v$i.so
  python -m astroz_b200.build --variant $lib $v > /dev/null 2>> gpurun_out/variant_sweep_$tag.err || continue
  ASTROZ_B200_LIB=$lib AZ_TAG="$v" python tools/k2_variant_timing.py >> $out 2>> gpurun_out/variant_sweep_$tag.err
done
cat $out
