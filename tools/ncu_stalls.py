"""Summarise an ncu --set full report: duration, pipe/issue utilisation, DRAM traffic, top stall reasons per kernel."""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__cycles_active.avg', 'lts__t_sector_hit_rate.pct',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed']
stall = [i for i, c in enumerate(h) if 'smsp__average_warps_issue_stalled' in c and 'per_issue_active' in c and 'not_issued' not in c]
for r in rows[2:]:
    print('==', r[h.index('Kernel Name')][:90])
    for w in want:
        if w in h:
            print('   %-75s %s %s' % (w, r[h.index(w)], rows[1][h.index(w)]))
    vals = sorted(((float(r[i]), h[i].split('stalled_')[1].split('_per')[0]) for i in stall), reverse=True)[:7]
    print('   stalls:', ', '.join('%s %.2f' % (n, v) for v, n in vals))
