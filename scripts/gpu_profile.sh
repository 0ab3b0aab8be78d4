#!/bin/bash
# ncu evidence for profiles/ (one GPU): launch list of the bench command, full-set capture of the hot kernels of the step,
# and of the stress configuration's kernels (matcher at G = 300, dense post-processing).  The .ncu-rep files are summarised
# ON THE BOX (scripts_ncu_summary.py, scripts/ncu_traffic.py) and removed when they would push gpurun_out/ past what
# gpurun copies back (64 MiB): the text summaries are what profiles/ keeps.
TAG=${1:-round2}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-graph --no-extras"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $BENCH > $OUT/${TAG}_ncu_bench.log 2>&1
echo "launch list exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'train_step_kernel|head_flat_kernel|head_rows_kernel|ssd_loss_kernel|ssd_loss_backward_kernel|filter_kernel|filter_dense_kernel|nms_kernel|nms_small_kernel|nms_rounds_kernel|match_kernel|pack_kernel' \
    -s 40 -c 24 -f -o $OUT/${TAG}_prof $BENCH > $OUT/${TAG}_ncu_full.log 2>&1
echo "full capture exit $?"
python scripts_ncu_summary.py $OUT/${TAG}_prof.ncu-rep > $OUT/${TAG}_ncu_summary.txt 2>&1
python scripts/ncu_traffic.py $OUT/${TAG}_prof.ncu-rep $OUT/${TAG}_ncu_traffic.json > /dev/null 2>&1
# stress configuration: the matcher alone at G = 300 (ssd._create_targets: the first launches of the script), then the fused
# training step and the dense post-processing kernels
timeout 600 ncu --set full --clock-control none -k regex:'match_kernel' -s 3 -c 2 -f -o $OUT/${TAG}_stress_match_prof \
    python scripts/stress_bench.py --steps 4 > $OUT/${TAG}_ncu_stress_match.log 2>&1
echo "stress matcher capture exit $?"
python scripts_ncu_summary.py $OUT/${TAG}_stress_match_prof.ncu-rep > $OUT/${TAG}_stress_match_ncu_summary.txt 2>&1
timeout 1200 ncu --set full --clock-control none \
    -k regex:'train_step_kernel|filter_kernel|filter_dense_kernel|nms_rounds_kernel' \
    -s 12 -c 10 -f -o $OUT/${TAG}_stress_prof python scripts/stress_bench.py --steps 4 > $OUT/${TAG}_ncu_stress.log 2>&1
echo "stress capture exit $?"
python scripts_ncu_summary.py $OUT/${TAG}_stress_prof.ncu-rep > $OUT/${TAG}_stress_ncu_summary.txt 2>&1
# keep the reports only while everything stays below ~56 MiB
for f in $OUT/${TAG}_stress_prof.ncu-rep $OUT/${TAG}_stress_match_prof.ncu-rep $OUT/${TAG}_prof.ncu-rep; do
  total=$(du -sm $OUT | cut -f1)
  if [ "$total" -gt 56 ]; then echo "dropping $f ($(du -sm $f | cut -f1) MiB; summary kept)"; rm -f $f; fi
done
ls -la $OUT | tail -12
