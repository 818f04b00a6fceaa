#!/bin/bash
# what the driver runs at round end: smoke, default bench, reference arm
TAG=${1:-def}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time python -c "import __graft_entry__ as g; g.smoke()" ) > $OUT/smoke.log 2>&1; tail -4 $OUT/smoke.log
( time python bench.py --impl reference --steps 20 --warmup 3 ) > $OUT/bench_reference.json 2> $OUT/bench_reference.err; tail -c 700 $OUT/bench_reference.json; tail -3 $OUT/bench_reference.err
( time python bench.py ) > $OUT/bench.json 2> $OUT/bench.err; tail -c 2600 $OUT/bench.json; tail -3 $OUT/bench.err
( time python bench.py --steps 20 --warmup 3 ) > $OUT/bench_20.json 2> $OUT/bench_20.err; python profiles/benchsum2.py $OUT/bench.json $OUT/bench_20.json
