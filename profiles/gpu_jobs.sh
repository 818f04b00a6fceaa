#!/bin/bash
# Job workloads (BASELINE configs[0]-[3]) at N GPUs.  usage: gpurun [--gpus N] --timeout 1500 -- 'bash profiles/gpu_jobs.sh <tag> <N>'
TAG=${1:-jobs}; N=${2:-1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() {  # workload steps extra...
  wl=$1; st=$2; shift 2
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --workload $wl --steps $st "$@" > $OUT/bench_${wl}_n$N.json 2> $OUT/bench_${wl}_n$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload $wl --steps $st "$@" > $OUT/bench_${wl}_n$N.json 2> $OUT/bench_${wl}_n$N.err
  fi
  echo "== $wl N=$N rc=$?"; tail -c 1500 $OUT/bench_${wl}_n$N.json | cut -c1-1500; tail -3 $OUT/bench_${wl}_n$N.err
}
if [ "$N" != "1" ]; then
  timeout 300 python -m pytest tests/test_many_gpu.py -x -q -k two_gpus > $OUT/pytest_two_gpus.log 2>&1; tail -3 $OUT/pytest_two_gpus.log
fi
run stitch2x3 5 --no-cpu-baseline
run stitch20x20 2 --no-cpu-baseline
run thumb64 2 --no-cpu-baseline
run align512_pairs 2 --no-cpu-baseline
run xcorr512 100 --no-cpu-baseline
