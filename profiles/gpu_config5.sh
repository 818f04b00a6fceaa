#!/bin/bash
# BASELINE config 5: batched xcorr_fft sweep over template sizes 128 .. 2048 (pad, FFT 256 .. 4096) at N GPUs.
# N = 1: with the CPU reference arm and the end-to-end number per size; N > 1: torchrun, weak scaling.
# usage: gpurun [--gpus N] --timeout 1500 -- 'bash profiles/gpu_config5.sh <tag> <N>'
TAG=${1:-c5}; N=${2:-1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for wl in xcorr128 xcorr256 xcorr512 xcorr1024 xcorr2048; do
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --workload $wl --steps 100 > $OUT/bench_${wl}_n1.json 2> $OUT/bench_${wl}_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --workload $wl --steps 100 --no-cpu-baseline > $OUT/bench_${wl}_n$N.json 2> $OUT/bench_${wl}_n$N.err
  fi
done
python - <<P
import json, glob
for f in sorted(glob.glob('$OUT/bench_*_n$N.json')):
    try:
        d = json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
        e, c = d.get('e2e') or {}, d.get('cpu_baseline') or {}
        print(f.split('/')[-1], 'value=%.0f' % d['value'], 'pipe=%.3f' % d['roofline']['pipeline']['frac'], 'e2e=%.0f' % (e.get('value') or 0),
              'cpu=%.1f on %s cores' % (c.get('value') or 0, c.get('cores')), 'clk', (d.get('clocks') or {}).get('sm_mhz'))
    except Exception as ex:
        print(f, 'ERR', ex)
P
