#!/bin/bash
# N-GPU session: multi-GPU tests + the torchrun bench line (+ the reference arm):
#   gpurun --gpus 2 -- 'bash tools/gpu_multi.sh 2 r02g'
n=${1:-2}; tag=${2:-r02}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_$tag.txt 2>&1
nvidia-smi topo -m >> gpurun_out/gpus_$tag.txt 2>&1
python -m pytest tests/test_gpu_multi.py -m gpu -x -q -rP > gpurun_out/tests_multi_$tag.log 2>&1; echo "pytest multi exit $?"; tail -4 gpurun_out/tests_multi_$tag.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $n --steps 200 --warmup 10 > gpurun_out/bench_n${n}_$tag.json 2> gpurun_out/bench_n${n}_$tag.err
echo "bench N=$n exit $?"; tail -c 1500 gpurun_out/bench_n${n}_$tag.err | tail -15
head -c 600 gpurun_out/bench_n${n}_$tag.json
