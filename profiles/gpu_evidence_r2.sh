#!/bin/bash
# Round-2 evidence of the headline workload: ncu --set full of the fast-path kernels, launch list of the default
# command (serial schedule), ncu of the image kernels in the config-4 block pass.
OUT=gpurun_out/ev_r2
mkdir -p $OUT
bash profiles/gpu_ncu.sh ev_r2 fbk_fast
FB_PIPELINE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fbk_ -c 400 --csv --log-file $OUT/launches_xcorr512.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/launches.log 2>&1
tail -2 $OUT/launches.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fbk_gauss|fbk_crop|fbk_minmax" -s 12 -c 6 -f \
    -o $OUT/prof_img python bench.py --workload align512_blocks --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_img.log 2>&1
tail -2 $OUT/ncu_img.log | cut -c1-200
