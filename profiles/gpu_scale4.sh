#!/bin/bash
# e2e scaling check at N = 4 ranks (NUMA-bound pinned buffers).  usage: gpurun --gpus 4 --timeout 600 -- 'bash profiles/gpu_scale4.sh <tag>'
TAG=${1:-scale4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 4 --steps 100 --warmup 3 --no-cpu-baseline > $OUT/bench_4gpu.json 2> $OUT/err_4gpu.txt
tail -c 300 $OUT/err_4gpu.txt
timeout 10 python profiles/benchsum.py < $OUT/bench_4gpu.json
