#!/bin/bash
# Sweep of kernel experiment switches on the default workload.  usage: gpurun -- 'bash profiles/gpu_sweep.sh <tag> "<flags...>" [extra bench args]'
TAG=${1:-sw}; FLAGS=${2:-0}; shift; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for f in $FLAGS; do
  timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 100 --fast-flags $f "$@" > $OUT/bench_$f.json 2>> $OUT/bench.err
done
python - <<PY
import json,glob
for f in "$FLAGS".split():
    try:
        d=json.loads(open("$OUT/bench_%s.json"%f).read().strip().splitlines()[-1])
        print("flags",f, round(d['value']), round(d['ms_per_step'],4), {k:round(v['ms_per_launch'],4) for k,v in d['roofline']['kernels'].items()}, d['clocks'])
    except Exception as e:
        print("flags",f,"ERR",e)
PY
tail -3 $OUT/bench.err
