#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2j}
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 --timeout=300 -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1; tail -4 $OUT/${TAG}_pytest.log
V=single-shot-detector_b200/lib_variants
SSDK_LIB=$PWD/$V/prev/libssdk.so timeout 300 python scripts/time_infer_variants.py 2>&1 | tail -1 | tee $OUT/${TAG}_infer_variants.txt
timeout 300 python scripts/time_infer_variants.py 2>&1 | tail -1 | tee -a $OUT/${TAG}_infer_variants.txt
timeout 600 python scripts/time_overlap.py 2>&1 | tail -1 | tee $OUT/${TAG}_overlap.json
