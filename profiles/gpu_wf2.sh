#!/bin/bash
TAG=${1:-wf}
OUT=gpurun_out/$TAG
mkdir -p $OUT
./profiles/microbench/dft_mma > $OUT/dft_mma_b200.txt 2>&1; cat $OUT/dft_mma_b200.txt
bash profiles/gpu_ncu_wf.sh $TAG
for wl in thumb_blocks50; do :; done
