#!/bin/bash
# Multi-GPU check (gpurun --gpus N): the 2-GPU parity test, then bench.py at 1..N ranks launched the way the driver does.
N=${1:-2}
TAG=${2:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_multi_gpus.txt
timeout 240 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > $OUT/${TAG}_multi_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_multi_pytest.log
tail -5 $OUT/${TAG}_multi_pytest.log
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    if [ $n -eq 1 ]; then
      timeout 240 python bench.py --gpus 1 --no-cpu-baseline > $OUT/${TAG}_scale_n$n.json 2> $OUT/${TAG}_scale_n$n.err
    else
      timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $n --no-cpu-baseline > $OUT/${TAG}_scale_n$n.json 2> $OUT/${TAG}_scale_n$n.err
    fi
    echo "n=$n exit $?"; python - <<PY
import json
try:
    d=json.loads([l for l in open('$OUT/${TAG}_scale_n$n.json') if l.startswith('{')][-1])
    print('n=%d value %.0f img/s  ms/step %.4f  e2e %.0f  train %.0f infer %.0f' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['breakdown']['train_images_per_sec'], d['breakdown']['infer_images_per_sec']))
except Exception as e:
    print('no result', e); print(open('$OUT/${TAG}_scale_n$n.err').read()[-2000:])
PY
  fi
done
