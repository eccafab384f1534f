"""Print the metrics we track from an ncu --page raw --csv export (profiling helper).
usage: ncu -i X.ncu-rep --page raw --csv | python tools/ncu_summary.py"""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.per_cycle_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'launch__grid_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__cycles_elapsed.max']
want += [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
want += [h for h in hdr if h.startswith('sm__inst_executed_pipe_') and h.endswith('.sum')]
idx = {h: i for i, h in enumerate(hdr)}
names = [r[idx['Kernel Name']].split('(')[0] for r in rows[2:]]
print('| metric | ' + ' | '.join(names) + ' |')
print('|---|' + '---:|' * len(names))
for w in want:
    if w in idx:
        vals = [r[idx[w]] for r in rows[2:]]
        if any(v not in ('0', '', '0.000000') for v in vals):
            print('| %s | %s |' % (w.replace('smsp__average_warps_issue_stalled_', 'stall_').replace('_per_issue_active.ratio', ''), ' | '.join(vals)))
