#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2e}
timeout 800 python scripts/sweep_train_split.py > $OUT/${TAG}_split_sweep.json 2> $OUT/${TAG}_split_sweep.err; tail -3 $OUT/${TAG}_split_sweep.err
python - <<PY
import json
d=json.load(open("$OUT/${TAG}_split_sweep.json"))
for k,v in d.items():
    print(k, {x:v[x] for x in ("C","G","B","auto","best","best_ms")}, round(v["roofline_ms"],4))
    s=v["sweep"]
    for m in (2,3,4,5): print("   m%d"%m, [s["m%d_s%d"%(m,sh)] for sh in (0,10,20,30)])
PY
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"; tail -3 $OUT/${TAG}_bench.err
python - <<PY
import json
d=json.loads([l for l in open("$OUT/${TAG}_bench.json") if l.startswith('{')][-1])
b=d['breakdown']
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'mode',d['launch_mode'])
print('train',b['train_ms_per_step'],b['train_frac_of_hbm_roofline'],'infer',b['infer_ms_per_step'],b['infer_frac_of_hbm_roofline'],'seq',b['sequential_graph_ms_per_step'])
print('roofline',json.dumps(d['roofline']))
print('check',json.dumps(d['check']))
print('cpu',json.dumps(d['cpu_baseline']))
print('kernels',json.dumps(b['kernel_ms_per_step']))
print('flat',json.dumps(b['roofline_loss_flat_pass']))
print('cfg4',json.dumps(b['cfg4_strong']))
print('stress',json.dumps(b['stress'])[:1500])
print('head',json.dumps({k:v for k,v in b['head_layout'].items() if 'ms' in k or 'frac' in k}))
PY
