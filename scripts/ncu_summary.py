"""Summarise an `ncu --page raw --csv` dump: one block per kernel launch with the metrics we track."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__cycles_elapsed.avg.per_second']
for r in rows[2:]:
    print('----', r[idx['Kernel Name']], 'id', r[idx['ID']])
    for w in want:
        if w in idx:
            print(f"  {w:70s} {r[idx[w]]:>16s} {units[idx[w]]}")
    st = []
    for h in hdr:
        if 'pcsamp_warps_issue_stalled' in h and not h.endswith('_not_issued'):
            try: st.append((float(r[idx[h]]), h.replace('smsp__pcsamp_warps_issue_stalled_', '')))
            except ValueError: pass
    tot = sum(v for v, _ in st) or 1
    print('  stalls: ' + ', '.join('%s %.0f%%' % (n, 100 * v / tot) for v, n in sorted(st, reverse=True)[:7]))
