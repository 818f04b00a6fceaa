#!/bin/bash
# racecheck + memcheck over the kernel families: warp-fused, fast path (all line lengths), generic, image operators, block pass
OUT=gpurun_out/sanitize2
mkdir -p $OUT
for tool in racecheck memcheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 99 --log-file $OUT/$tool.log \
      python -m pytest tests/test_xcorr_gpu.py tests/test_image_gpu.py tests/test_matcher_gpu.py -m gpu -x -q \
      -k "golden or warp_fused_against or float64_pipeline or sigma_masks or dog_random or crop_blocks_vs or stack_minmax or block_grid_pass or renderer_matcher" > $OUT/pytest_$tool.log 2>&1
  echo "$tool rc=$?"; tail -2 $OUT/pytest_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $OUT/$tool.log | tail -2
  grep -E "Error: Race reported|Invalid|Error:" $OUT/$tool.log | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | sort -rn | head -12
done
