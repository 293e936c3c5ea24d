"""Key metrics of every kernel in an `ncu --page raw --csv` dump (ncu -i X.ncu-rep --page raw --csv > raw.csv)."""
import csv, sys
KEYS = ['gpu__time_duration.sum', 'sm__cycles_active.avg', 'smsp__cycles_active.avg', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__clock', 'gpc__cycles_elapsed.avg.per_second', 'sm__cycles_elapsed.avg.per_second']
r = list(csv.reader(open(sys.argv[1])))
hdr, units, rows = r[0], r[1], r[2:]
for row in rows:
    print('==', row[hdr.index('Kernel Name')][:90], ' grid', row[hdr.index('Grid Size')] if 'Grid Size' in hdr else '')
    for i, h in enumerate(hdr):
        if h in KEYS or ('warps_issue_stalled' in h and h.endswith('per_issue_active.ratio')):
            try:
                v = float(row[i].replace(',', ''))
            except ValueError:
                continue
            if 'issue_stalled' in h and v < 0.15:
                continue
            print(f'   {h} [{units[i]}] = {row[i]}')
