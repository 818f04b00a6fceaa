#!/bin/bash
# K2 "solo" experiment: parity with the switch on, then xcorr512 / xcorr256 / xcorr512_nopad with and without it
TAG=${1:-solo}
OUT=gpurun_out/$TAG
mkdir -p $OUT
FEABAS_CUDA_OPTIONS=k2_solo=1 timeout 600 python -m pytest tests/test_xcorr_gpu.py -m gpu -x -q > $OUT/pytest_solo.log 2>&1; tail -3 $OUT/pytest_solo.log
for wl in xcorr512 xcorr256 xcorr512_nopad; do
  for solo in 0 1; do
    FEABAS_CUDA_OPTIONS=k2_solo=$solo timeout 300 python bench.py --workload $wl --steps 100 --no-cpu-baseline --no-e2e > $OUT/bench_${wl}_s$solo.json 2> $OUT/bench_${wl}_s$solo.err
    python - <<P
import json
d=json.loads([l for l in open('$OUT/bench_${wl}_s$solo.json').read().splitlines() if l.startswith('{')][-1])
r=d['roofline']
print('$wl solo=$solo value=%.0f ms/step=%.3f pipe=%.3f' % (d['value'], d['ms_per_step'], r['pipeline']['frac']), {k:round(v['ms_per_launch'],4) for k,v in r['kernels'].items()}, d['clocks']['sm_mhz'])
P
  done
done
