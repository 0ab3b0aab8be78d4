#!/bin/bash
# Multi-GPU check (gpurun --gpus N): the 2-GPU parity tests, bench.py at 1..N ranks launched the way the driver does (every line
# carries check.sharded_equals_single and breakdown.cfg4_strong), and the host-to-device topology table.
N=${1:-2}
TAG=${2:-round2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_multi_gpus.txt
timeout 400 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -p no:cacheprovider > $OUT/${TAG}_multi_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_multi_pytest.log
tail -5 $OUT/${TAG}_multi_pytest.log
timeout 300 python scripts/h2d_topology.py > $OUT/${TAG}_h2d_topology.txt 2>&1; echo "h2d exit $?"; tail -30 $OUT/${TAG}_h2d_topology.txt
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    if [ $n -eq 1 ]; then
      timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_scale_n$n.json 2> $OUT/${TAG}_scale_n$n.err
    else
      timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_scale_n$n.json 2> $OUT/${TAG}_scale_n$n.err
    fi
    echo "n=$n exit $?"; python - <<PY
import json
try:
    d=json.loads([l for l in open('$OUT/${TAG}_scale_n$n.json') if l.startswith('{')][-1])
    b=d['breakdown']
    print('n=%d value %.0f img/s  ms/step %.4f  e2e %.0f (%.1f GB/s/GPU)  train %.0f infer %.0f' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['h2d_gb_per_s_per_gpu'], b['train_images_per_sec'], b['infer_images_per_sec']))
    print('   check', json.dumps(d['check']))
    print('   cfg4', json.dumps(b.get('cfg4_strong')))
    print('   numa', json.dumps(d['e2e'].get('host_placement')))
except Exception as e:
    print('no result', e); print(open('$OUT/${TAG}_scale_n$n.err').read()[-3000:])
PY
  fi
done
