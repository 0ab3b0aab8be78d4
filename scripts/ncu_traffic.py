"""python scripts/ncu_traffic.py <file.ncu-rep> <out.json>: per-kernel DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum per
launch, averaged over the captured launches) and duration from one `ncu --set full` capture; bench.py reads the JSON for
`roofline.traffic`."""
import csv, json, re, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
def col(name): return hdr.index(name)
scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
acc = {}
for r in rows[2:]:
    name = re.sub(r'^void ', '', r[col('Kernel Name')]).split('(')[0].split('<')[0]
    b = sum(float(r[col(m)]) * scale[units[col(m)]] for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
    t = float(r[col('gpu__time_duration.sum')]) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}[units[col('gpu__time_duration.sum')]]
    a = acc.setdefault(name, {'launches': 0, 'dram_bytes': 0.0, 'us': 0.0})
    a['launches'] += 1; a['dram_bytes'] += b; a['us'] += t
res = {k: {'launches': v['launches'], 'dram_bytes_per_launch': v['dram_bytes'] / v['launches'], 'us_per_launch_under_ncu': v['us'] / v['launches']} for k, v in acc.items()}
res['_source'] = rep
json.dump(res, open(out, 'w'), indent=1)
print(json.dumps(res, indent=1))
