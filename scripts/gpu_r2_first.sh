#!/bin/bash
# round 2, first GPU call: parity tests of the new kernels, then the tuning sweep
TAG=${1:-r2a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 --timeout=300 -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -40 $OUT/${TAG}_pytest.log
timeout 900 python scripts/tune_round2.py > $OUT/${TAG}_tune.json 2> $OUT/${TAG}_tune.err; echo "tune exit $?"
cat $OUT/${TAG}_tune.json; tail -5 $OUT/${TAG}_tune.err
