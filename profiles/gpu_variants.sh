#!/bin/bash
# compare builds of the library (feabas_b200/csrc/variants/lib_<name>.so) on a few workloads
TAG=${1:-var}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for v in "$@"; do
  export FEABAS_CUDA_LIB=$PWD/feabas_b200/csrc/variants/lib_$v.so
  timeout 300 python -m pytest tests/test_xcorr_gpu.py -m gpu -x -q -k "golden or seeded" > $OUT/pytest_$v.log 2>&1; tail -1 $OUT/pytest_$v.log
  for wl in ${WORKLOADS:-xcorr512 xcorr256}; do
    timeout 300 python bench.py --workload $wl --steps 100 --no-cpu-baseline --no-e2e > $OUT/bench_${wl}_$v.json 2> $OUT/bench_${wl}_$v.err
    python - <<P
import json
try:
    d=json.loads([l for l in open('$OUT/bench_${wl}_$v.json').read().splitlines() if l.startswith('{')][-1])
    r=d['roofline']
    print('$wl $v value=%.0f ms/step=%.3f pipe=%.3f' % (d['value'], d['ms_per_step'], r['pipeline']['frac']), {k:round(x['ms_per_launch'],4) for k,x in r['kernels'].items()}, d['clocks']['sm_mhz'])
except Exception as e:
    print('$wl $v failed', e)
P
  done
done
