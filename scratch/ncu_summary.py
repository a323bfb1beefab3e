import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'launch__waves_per_multiprocessor', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__thread_inst_executed.sum']
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    print(name[:60])
    for w in want:
        if w in hdr: print('   %-70s %s %s' % (w, r[hdr.index(w)], rows[1][hdr.index(w)]))
    stall = [i for i, h in enumerate(hdr) if 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio')]
    d = sorted([(float(r[i].replace(',', '')), hdr[i].replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')) for i in stall], reverse=True)[:7]
    print('   stalls/issue:', ', '.join('%s %.2f' % (b, a) for a, b in d))
