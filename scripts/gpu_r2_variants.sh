#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2f}
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_head.py -m gpu -q --maxfail=8 --timeout=300 -p no:cacheprovider -k "postprocess or candidate or overflow or detect or nms or head_detect or by_label" > $OUT/${TAG}_pytest.log 2>&1; tail -5 $OUT/${TAG}_pytest.log
for v in default nocap cap3; do
  if [ $v = default ]; then unset SSDK_LIB; else export SSDK_LIB=$PWD/single-shot-detector_b200/lib_variants/$v/libssdk.so; fi
  timeout 300 python scripts/time_infer_variants.py 2>&1 | tail -1
done | tee $OUT/${TAG}_variants.txt
unset SSDK_LIB
timeout 600 python scripts/tune_round2.py --quick > $OUT/${TAG}_tune.json 2> $OUT/${TAG}_tune.err; echo "tune exit $?"; python - <<PY
import json
d=json.load(open('$OUT/${TAG}_tune.json'))
print(json.dumps({k:d[k] for k in ('infer','infer_kernels_ms','stress')}))
PY
tail -3 $OUT/${TAG}_tune.err
