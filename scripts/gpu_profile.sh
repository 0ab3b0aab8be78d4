#!/bin/bash
# ncu evidence for profiles/ (one GPU): launch list of the bench command, full-set capture of the hot kernels of the step,
# and of the stress configuration's kernels (matcher at G = 300, dense post-processing).
TAG=${1:-round2}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-graph --no-extras"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $BENCH > $OUT/${TAG}_ncu_bench.log 2>&1
echo "launch list exit $?"; tail -2 $OUT/${TAG}_ncu_bench.log
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'train_step_kernel|head_flat_kernel|head_rows_kernel|ssd_loss_kernel|ssd_loss_backward_kernel|filter_kernel|filter_dense_kernel|nms_kernel|nms_small_kernel|nms_rounds_kernel|match_kernel|pack_kernel' \
    -s 40 -c 40 -f -o $OUT/${TAG}_prof $BENCH > $OUT/${TAG}_ncu_full.log 2>&1
echo "full capture exit $?"; tail -2 $OUT/${TAG}_ncu_full.log
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'match_kernel|train_step_kernel|filter_kernel|filter_dense_kernel|nms_rounds_kernel' \
    -s 30 -c 14 -f -o $OUT/${TAG}_stress_prof python scripts/stress_bench.py --steps 4 > $OUT/${TAG}_ncu_stress.log 2>&1
echo "stress capture exit $?"; tail -2 $OUT/${TAG}_ncu_stress.log
ls -la $OUT | tail -8
