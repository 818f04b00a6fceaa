#!/bin/bash
# ncu --set full capture of the fast-path kernels (3 launches after warm-up) + launch list.
# usage: gpurun --timeout 900 -- 'bash profiles/gpu_ncu.sh <tag> [kernel regex] [extra bench args]'
TAG=${1:-n}; RE=${2:-fbk_fast}; shift; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
export FB_PIPELINE=1   # serial schedule: whole batch per launch, as in bench.py's per-kernel roofline pass
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$RE -s 9 -c 3 -f \
    -o $OUT/prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e "$@" > $OUT/ncu_full.log 2>&1
tail -3 $OUT/ncu_full.log
