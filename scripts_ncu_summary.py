"""Summarise an .ncu-rep (raw page) into a compact per-kernel table: python scripts_ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'smsp__inst_executed.sum']
stalls = [h for h in hdr if h.startswith('smsp__average_warp') and 'issue_stalled' in h and h.endswith('_per_issue_active.ratio')] or \
         [h for h in hdr if 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct')]
for r in rows[2:]:
    print('=====', r[hdr.index('Kernel Name')][:90])
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print('  %-72s %s %s' % (w, r[i], units[i]))
    st = []
    for h in stalls:
        try:
            st.append((float(r[hdr.index(h)]), h))
        except ValueError:
            pass
    for v, h in sorted(st, reverse=True)[:7]:
        print('    stall %-80s %.2f' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('smsp__average_warp_latency_issue_stalled_', ''), v))
