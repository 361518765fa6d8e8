"""Print the counters that matter from an `ncu --page raw --csv` dump (development aid): python scripts/ncu_summary.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__warps_eligible.avg.per_cycle_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_ldgsts.sum']
stall = [h for h in hdr if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h]
idx = {h: i for i, h in enumerate(hdr)}
num = lambda s: float(s.replace(',', '')) if s else 0.0
for r in rows[2:]:
    print('-----', r[idx['Kernel Name']][:50], r[idx['Grid Size']], r[idx['Block Size']])
    for w in want:
        if w in idx:
            print('   %-70s %s %s' % (w, r[idx[w]], rows[1][idx[w]]))
    st = sorted([(num(r[idx[h]]), h.replace('smsp__pcsamp_warps_issue_stalled_', '')) for h in stall], reverse=True)
    tot = sum(v for v, _ in st) or 1
    print('   stalls:', ', '.join('%s %.0f%%' % (h, 100 * v / tot) for v, h in st[:8]))
