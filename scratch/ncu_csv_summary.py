"""Summarise `ncu --page raw --csv` exports: python scratch/ncu_csv_summary.py gpurun_out/ncu/*.raw.csv"""
import csv, sys
WANT = ['gpu__time_duration.sum','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem',
 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
 'dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','sm__cycles_elapsed.max','sm__cycles_active.avg',
 'smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed_op_shared_atom.sum']
for f in sys.argv[1:]:
    rows = list(csv.reader(open(f)))
    hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr = rows[hi]
    for r in rows[hi + 2:]:
        if len(r) < len(hdr): continue
        print('==', f, r[hdr.index('Kernel Name')][:90])
        for w in WANT:
            if w in hdr: print('   %-80s %s %s' % (w, r[hdr.index(w)], rows[hi+1][hdr.index(w)]))
        def num(x):
            try: return float(x.replace(',', ''))
            except ValueError: return 0.0
        st = [(num(r[i]), h) for i, h in enumerate(hdr) if h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('not_issued')]
        tot = sum(x for x, _ in st) or 1
        for x, h in sorted(st, reverse=True)[:9]:
            print('   stall %-45s %5.1f%%' % (h.replace('smsp__pcsamp_warps_issue_stalled_', ''), 100 * x / tot))
