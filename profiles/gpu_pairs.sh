#!/bin/bash
# config 4 as a job list: align512_pairs (64 section pairs, strong scaling list) + the single-pair block pass
TAG=${1:-pairs}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python -m pytest tests/test_many_gpu.py tests/test_image_gpu.py -m gpu -x -q > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
timeout 600 python bench.py --workload align512_pairs --steps 2 --no-cpu-baseline > $OUT/bench_align512_pairs.json 2> $OUT/bench_pairs.err; tail -2 $OUT/bench_pairs.err
timeout 300 python bench.py --workload align512_blocks --steps 30 --no-cpu-baseline > $OUT/bench_align512_blocks.json 2> $OUT/bench_blocks.err
python profiles/benchsum2.py $OUT/bench_align512_pairs.json $OUT/bench_align512_blocks.json
