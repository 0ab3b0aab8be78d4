#!/bin/bash
# compute-sanitizer over the GPU tests (round 2 kernels: fused training step, bounded candidate regions, rounds, dense filter, split scan / flat loops)
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r3k}
F=$OUT/${TAG}_sanitizer.txt
echo "compute-sanitizer on B200 (round 2)" > $F
echo "memcheck: compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_head.py tests/test_gpu_parity.py tests/test_gpu_backward.py -m gpu -q -x -k 'not full_size and not properties_full and not cfg4 and not plain_c and not many_classes'" >> $F
timeout 1500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_head.py tests/test_gpu_parity.py tests/test_gpu_backward.py -m gpu -q -x -p no:cacheprovider \
   -k "not full_size and not properties_full and not cfg4 and not plain_c and not many_classes" 2>&1 | grep -v "^$" | tail -25 >> $F
echo "" >> $F
echo "racecheck: compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_head.py -m gpu -q -x -k 'fused_train_step or bounded_candidate or overflowing or postprocess_golden or head_loss_matches or by_label or split_phase or segment_sizes or chunk_boundaries'" >> $F
timeout 1500 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_head.py -m gpu -q -x -p no:cacheprovider \
   -k "fused_train_step or bounded_candidate or overflowing or postprocess_golden or head_loss_matches or by_label or split_phase or segment_sizes or chunk_boundaries" 2>&1 | grep -v "^$" | tail -40 >> $F
cat $F | cut -c1-300
