#!/bin/bash
# bench + ncu --set full of the three fast-path kernels for the off-1024 families.  usage: gpurun --timeout 1500 -- 'bash profiles/gpu_ncu_families.sh <tag>'
TAG=${1:-fam}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export FB_PIPELINE=1
for wl in align280 thumb150 xcorr300 xcorr1024 xcorr2048 xcorr256 xcorr512_nopad; do
  timeout 200 python bench.py --workload $wl --steps 30 --no-cpu-baseline --no-e2e > $OUT/bench_$wl.json 2>> $OUT/bench.err
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:fbk_fast -s 9 -c 3 -f -o $OUT/prof_$wl \
      python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_$wl.log 2>&1
  tail -1 $OUT/ncu_$wl.log
  python profiles/ncu_summary.py $OUT/prof_$wl.ncu-rep > $OUT/ncu_full_$wl.txt 2>&1
  python profiles/ncu_source.py $OUT/prof_$wl.ncu-rep 14 > $OUT/ncu_source_$wl.txt 2>&1
  rm -f $OUT/prof_$wl.ncu-rep
done
ls $OUT
