#!/bin/bash
# Quick GPU check: parity tests + default bench (no CPU arm).  usage: gpurun --timeout 600 -- 'bash profiles/gpu_quick.sh <tag> [extra bench args]'
TAG=${1:-q}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline --no-e2e "$@" > $OUT/bench.json 2> $OUT/bench.err; tail -c 2500 $OUT/bench.json; tail -5 $OUT/bench.err
