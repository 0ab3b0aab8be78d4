#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (both arms), ncu launch list, ncu full capture of the hot kernels.
# Usage (from the repo root, on the GPU box): bash scripts/gpu_check.sh <tag> [full|quick]
TAG=${1:-r01}
MODE=${2:-full}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
python -c "import os; print('cpus', os.cpu_count(), len(os.sched_getaffinity(0)))" >> $OUT/${TAG}_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -15 $OUT/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> $OUT/${TAG}_smoke.log
tail -3 $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
cat $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
if [ "$MODE" = "full" ]; then
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "ref exit $?"
  cat $OUT/${TAG}_bench_ref.json
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-graph > $OUT/${TAG}_ncu_bench.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on \
      -k regex:'ssd_loss_kernel|ssd_loss_backward_kernel|filter_kernel|nms_kernel|nms_small_kernel|match_kernel|pack_kernel|head_flat_kernel|head_rows_kernel' \
      -s 60 -c 36 -f -o $OUT/${TAG}_prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-graph > $OUT/${TAG}_ncu_full.log 2>&1
  ls -la $OUT
fi
