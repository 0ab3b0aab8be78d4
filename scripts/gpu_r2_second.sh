#!/bin/bash
TAG=${1:-r2b}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
python -c "import os; print('cpus', os.cpu_count(), len(os.sched_getaffinity(0)))" >> $OUT/${TAG}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 --timeout=300 -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -30 $OUT/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> $OUT/${TAG}_smoke.log
tail -3 $OUT/${TAG}_smoke.log
timeout 900 python scripts/tune_round2.py > $OUT/${TAG}_tune.json 2> $OUT/${TAG}_tune.err; echo "tune exit $?"
cat $OUT/${TAG}_tune.json; tail -5 $OUT/${TAG}_tune.err
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
cat $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "ref exit $?"
cat $OUT/${TAG}_bench_ref.json; tail -3 $OUT/${TAG}_bench_ref.err
