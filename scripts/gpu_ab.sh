#!/bin/bash
# A/B of library builds on ONE box (boxes of the pool differ by a few per cent, so comparisons across calls mean little):
#   bash scripts/gpu_ab.sh <tag> <timing script> <variant> [<variant> ...]
# `default` = single-shot-detector_b200/lib/libssdk.so, any other name = single-shot-detector_b200/lib_variants/<name>/libssdk.so
# (built with `make -C single-shot-detector_b200/csrc OUT=../lib_variants/<name> EXTRA="-D..."`).  The timing scripts
# (scripts/time_*_variants.py, tune_round2.py --quick, sweep_train_split.py) print one JSON line each.
OUT=gpurun_out; mkdir -p $OUT
TAG=$1; SCRIPT=$2; shift 2
V=single-shot-detector_b200/lib_variants
for lib in "$@"; do
  if [ $lib = default ]; then unset SSDK_LIB; else export SSDK_LIB=$PWD/$V/$lib/libssdk.so; fi
  timeout 600 python $SCRIPT 2>&1 | tail -1
done | tee $OUT/${TAG}_ab.txt
