#!/bin/bash
# One GPU session: parity tests, smoke, bench (default + the other workloads), ncu launch list and full
# captures of the fast-path kernels.   usage (here): gpurun --timeout 1800 -- 'bash profiles/gpu_round.sh <tag>'
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; tail -c 3000 $OUT/bench.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/bench_reference.json 2>> $OUT/bench.err; tail -c 600 $OUT/bench_reference.json
for wl in xcorr512_nopad xcorr256 xcorr128 xcorr1024 xcorr2048 stitch_fine thumb150 align280 xcorr300; do
  timeout 120 python bench.py --workload $wl --steps 30 --no-cpu-baseline > $OUT/bench_$wl.json 2>> $OUT/bench.err
done
# matcher-level workloads (configs 1, 3, 4 through stitching_matcher / section_matcher / bboxes_mesh_renderer_matcher), CPU port beside them
timeout 300 python bench.py --workload align512_blocks --steps 50 > $OUT/bench_align512_blocks.json 2>> $OUT/bench.err
for wl in stitch2x3 thumb_sections; do
  timeout 600 python bench.py --workload $wl --steps 5 --workers 8 > $OUT/bench_$wl.json 2>> $OUT/bench.err
done
# ncu serialises kernels: capture the serial schedule (one stream, whole batch per launch), which is also what bench.py's
# per-kernel roofline numbers are taken from
export FB_PIPELINE=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fbk -c 400 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fbk_fast -s 9 -c 3 -f \
    -o $OUT/prof_fast python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full.log 2>&1
ls -la $OUT
