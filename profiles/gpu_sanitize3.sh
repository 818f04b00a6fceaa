#!/bin/bash
# synccheck + initcheck over the golden tests (every kernel family)
OUT=gpurun_out/sanitize3
mkdir -p $OUT
for tool in synccheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 99 --log-file $OUT/$tool.log \
      python -m pytest tests/test_xcorr_gpu.py tests/test_image_gpu.py tests/test_matcher_gpu.py -m gpu -x -q \
      -k "golden or warp_fused_against or dog_random or crop_blocks_vs or block_grid_pass" > $OUT/pytest_$tool.log 2>&1
  echo "$tool rc=$?"; tail -2 $OUT/pytest_$tool.log
  grep -E "ERROR SUMMARY" $OUT/$tool.log | tail -2
  grep -E "Error:|Uninitialized|Barrier" $OUT/$tool.log | sed 's/+0x[0-9a-f]*//; s/0x[0-9a-f]*//g' | sort | uniq -c | sort -rn | head -12
done
