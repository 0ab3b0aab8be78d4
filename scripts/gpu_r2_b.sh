#!/bin/bash
# A/B of the score-scan variants (same box), the training-step sweep, then the ncu evidence.
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2b}
V=single-shot-detector_b200/lib_variants
for lib in default base u8c3 u8c4 u6c4 u4c5 default; do
  if [ $lib = default ]; then unset SSDK_LIB; else export SSDK_LIB=$PWD/$V/$lib/libssdk.so; fi
  timeout 300 python scripts/time_infer_variants.py 2>&1 | tail -1
done | tee $OUT/${TAG}_filter_variants.txt
unset SSDK_LIB
SSDK_PDL=0 timeout 300 python scripts/time_infer_variants.py 2>&1 | tail -1 | tee -a $OUT/${TAG}_filter_variants.txt
timeout 900 python scripts/tune_round2.py > $OUT/${TAG}_tune.json 2> $OUT/${TAG}_tune.err; tail -c 3000 $OUT/${TAG}_tune.json
bash scripts/gpu_profile.sh $TAG
