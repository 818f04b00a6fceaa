#!/bin/bash
TAG=${1:-famq}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests/test_xcorr_gpu.py -x -q -k "fast or golden_seeded" 2>&1 | tail -3
for wl in align280 thumb150 xcorr300 xcorr1024 xcorr2048 xcorr256 xcorr512_nopad xcorr128 xcorr512; do
  timeout 200 python bench.py --workload $wl --steps 30 --no-cpu-baseline --no-e2e > $OUT/bench_$wl.json 2>> $OUT/bench.err
done
python profiles/benchsum2.py $OUT/bench_*.json
