#!/bin/bash
# ncu --set full of the warp-fused kernel on the stitch_fine workload.  usage: gpurun --timeout 900 -- 'bash profiles/gpu_ncu_wf.sh <tag>'
TAG=${1:-wf}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests/test_xcorr_gpu.py -x -q -k "warp_fused" 2>&1 | tail -3
timeout 120 python bench.py --workload stitch_fine --steps 30 --no-cpu-baseline --no-e2e > $OUT/bench_stitch_fine.json 2> $OUT/bench.err; cut -c1-300 $OUT/bench_stitch_fine.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fbk_wf -s 4 -c 1 -f \
    -o $OUT/prof_wf python bench.py --workload stitch_fine --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log; ls -la $OUT
