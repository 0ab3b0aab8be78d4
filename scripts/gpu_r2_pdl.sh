#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2m}
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 --timeout=300 -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1; tail -4 $OUT/${TAG}_pytest.log
for pdl in 1 0; do SSDK_PDL=$pdl timeout 300 python scripts/time_infer_variants.py 2>&1 | tail -1; done | tee $OUT/${TAG}_pdl.txt
timeout 600 python scripts/time_overlap.py 2>&1 | tail -1 | tee $OUT/${TAG}_overlap.json
timeout 600 python scripts/tune_round2.py --quick > $OUT/${TAG}_tune.json 2> $OUT/${TAG}_tune.err; python - <<PY
import json
d=json.load(open('$OUT/${TAG}_tune.json'))
print(json.dumps({k:d[k] for k in ('train_fused_default','infer','step_two_streams','stress')}))
PY
