#!/bin/bash
# compute-sanitizer (memcheck, then racecheck) over a subset of the GPU tests.  usage: bash profiles/gpu_sanitize.sh [tools] [pytest -k expr]
OUT=gpurun_out/sanitize
mkdir -p $OUT
TOOLS=${1:-"memcheck racecheck"}
KEXPR=${2:-"golden or resize or float64 or crop"}
for tool in $TOOLS; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 99 --log-file $OUT/$tool.log \
      python -m pytest tests/test_xcorr_gpu.py tests/test_image_gpu.py -m gpu -x -q -k "$KEXPR" > $OUT/pytest_$tool.log 2>&1
  echo "$tool rc=$?"; tail -2 $OUT/pytest_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $OUT/$tool.log | tail -2
  grep -E "Error: Race reported" $OUT/$tool.log | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | sort -rn | head -12
done
