#!/bin/bash
# ncu --set full of the image kernels (crop, Gaussians) on the align512_blocks workload
TAG=${1:-img}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python bench.py --workload align512_blocks --steps 30 --no-cpu-baseline > $OUT/bench_align512_blocks.json 2> $OUT/bench.err; cut -c1-400 $OUT/bench_align512_blocks.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fbk_gauss|fbk_crop|fbk_minmax" -s 12 -c 6 -f \
    -o $OUT/prof_img python bench.py --workload align512_blocks --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log
