#!/bin/bash
# config 4 job list (align512_pairs, strong scaling) and the headline workload at N GPUs
TAG=${1:-pscale}; N=${2:-8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for wl in align512_pairs xcorr512; do
  st=2; [ $wl = xcorr512 ] && st=100
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload $wl --steps $st --no-cpu-baseline > $OUT/bench_${wl}_n$N.json 2> $OUT/bench_${wl}_n$N.err
  echo "== $wl N=$N rc=$?"; tail -2 $OUT/bench_${wl}_n$N.err | cut -c1-300
done
python profiles/benchsum2.py $OUT/bench_*_n$N.json
