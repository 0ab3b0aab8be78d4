#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2k}
V=single-shot-detector_b200/lib_variants
for lib in default nms256; do
  if [ $lib = default ]; then unset SSDK_LIB; else export SSDK_LIB=$PWD/$V/$lib/libssdk.so; fi
  timeout 300 python scripts/time_infer_variants.py 2>&1 | tail -1
  timeout 600 python scripts/time_overlap.py 2>&1 | tail -1
done | tee $OUT/${TAG}_nms256.txt
