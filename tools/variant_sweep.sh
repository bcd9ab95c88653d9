#!/bin/bash
# Build library variants ON THE GPU BOX (same sources, extra -D flags) and run a timing tool against each:
#   gpurun -- 'bash tools/variant_sweep.sh r02e k2_variant_timing "AZ_K2_LANES=1 AZ_DEFAULT_K2_BLOCKS=5" "AZ_K2_LANES=2"'
tag=$1; tool=$2; shift; shift
mkdir -p gpurun_out
out=gpurun_out/variant_${tool}_$tag.jsonl; : > $out
echo '{"variant": "default"}' >> $out
AZ_TAG=default python tools/$tool.py >> $out 2>> gpurun_out/variant_${tool}_$tag.err
i=0
for v in "$@"; do
  i=$((i+1)); lib=/tmp/libaz_v$i.so
  python -m astroz_b200.build --variant $lib $v > /dev/null 2>> gpurun_out/variant_${tool}_$tag.err || continue
  echo "{\"variant\": \"$v\"}" >> $out
  ASTROZ_B200_LIB=$lib AZ_TAG="$v" python tools/$tool.py >> $out 2>> gpurun_out/variant_${tool}_$tag.err
done
cat $out
