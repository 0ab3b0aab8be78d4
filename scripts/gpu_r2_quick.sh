#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2d}
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_head.py -m gpu -q --maxfail=8 --timeout=300 -p no:cacheprovider -k "postprocess or candidate or overflow or detect or nms or head_detect or by_label or streams or differentiable" > $OUT/${TAG}_pytest.log 2>&1; tail -12 $OUT/${TAG}_pytest.log
timeout 300 python scripts/time_infer_variants.py 2>&1 | tail -1 | tee $OUT/${TAG}_infer.txt
timeout 600 python scripts/tune_round2.py --quick > $OUT/${TAG}_tune.json 2> $OUT/${TAG}_tune.err; echo "tune exit $?"; cat $OUT/${TAG}_tune.json; tail -3 $OUT/${TAG}_tune.err
