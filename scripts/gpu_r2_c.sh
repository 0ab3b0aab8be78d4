#!/bin/bash
# Flat pass without per-chunk bounds tests / segment lookups: parity tests, then A/B against the previous build, then the sweep.
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2c}
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 --timeout=300 -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1; tail -4 $OUT/${TAG}_pytest.log
V=single-shot-detector_b200/lib_variants
SSDK_LIB=$PWD/$V/prev/libssdk.so timeout 600 python scripts/tune_round2.py --quick > $OUT/${TAG}_tune_prev.json 2> $OUT/${TAG}_tune_prev.err
timeout 900 python scripts/tune_round2.py > $OUT/${TAG}_tune.json 2> $OUT/${TAG}_tune.err
python - <<PY
import json
for f in ('$OUT/${TAG}_tune_prev.json', '$OUT/${TAG}_tune.json'):
    try:
        d=json.load(open(f))
        print(f); print(json.dumps({k:d.get(k) for k in ('train_fused_default','train_fused_sweep_ms','train_fused_best','train_unfused','flat_pass_alone','train_fwd_bwd','matcher_alone_ms','infer','step_two_streams','step_one_stream')}))
        print(json.dumps(d['stress']))
    except Exception as e:
        print(f, 'failed', e)
PY
