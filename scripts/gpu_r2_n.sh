#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2n}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=8 --timeout=300 -p no:cacheprovider -k "postprocess or candidate or overflow or detect or nms or by_label" > $OUT/${TAG}_pytest.log 2>&1; tail -3 $OUT/${TAG}_pytest.log
timeout 300 python scripts/time_infer_variants.py 2>&1 | tail -1
timeout 600 python scripts/tune_round2.py --quick > $OUT/${TAG}_tune.json 2> $OUT/${TAG}_tune.err; python - <<PY
import json
d=json.load(open('$OUT/${TAG}_tune.json'))
print(json.dumps({k:d[k] for k in ('train_fused_default','infer','infer_kernels_ms','step_two_streams','step_one_stream')}))
print(json.dumps(d['stress']))
PY
