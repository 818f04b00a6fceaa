OUT=gpurun_out/r1e
export FB_PIPELINE=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fbk -c 400 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fbk_fast -s 9 -c 3 -f \
    -o $OUT/prof_fast python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log
