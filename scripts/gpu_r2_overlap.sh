#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2k}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=8 --timeout=300 -p no:cacheprovider -k "postprocess or candidate or overflow or detect or nms or by_label" > $OUT/${TAG}_pytest.log 2>&1; tail -3 $OUT/${TAG}_pytest.log
timeout 600 python scripts/time_overlap.py 2>&1 | tail -1 | tee $OUT/${TAG}_overlap.json
timeout 300 python scripts/time_infer_variants.py 2>&1 | tail -1
