#!/bin/bash
# Weak-scaling run on one box: bench.py at N = 2, 4, 8 ranks (torchrun, NCCL only for the barrier / max reduction).
# usage: gpurun --gpus 8 --timeout 600 -- 'bash profiles/gpu_scale.sh <tag>'
TAG=${1:-scale}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.max.sm,power.limit --format=csv > $OUT/gpus.txt 2>&1
for N in 8 4 2; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
      bench.py --gpus $N --steps 100 --warmup 3 --no-cpu-baseline > $OUT/bench_${N}gpu.json 2> $OUT/err_${N}gpu.txt
  tail -c 400 $OUT/err_${N}gpu.txt
  timeout 10 python profiles/benchsum.py < $OUT/bench_${N}gpu.json
done
