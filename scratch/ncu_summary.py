import csv, subprocess, sys
WANT = ['gpu__time_duration.sum','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__grid_size',
 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
 'dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sectors_op_red.sum','lts__t_sectors_op_atom.sum','smsp__inst_executed.sum','sm__cycles_elapsed.max','sm__cycles_active.avg',
 'lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','smsp__thread_inst_executed_per_inst_executed.ratio']
for f in sys.argv[1:]:
    out = subprocess.run(['ncu','-i',f,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    for r in rows[2:]:
        print('==', f, r[hdr.index('Kernel Name')][:70])
        for w in WANT:
            if w in hdr: print('   %-80s %s' % (w, r[hdr.index(w)]))
        st = [(float(r[i]) if r[i].replace('.','').isdigit() else 0, h) for i,h in enumerate(hdr) if h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('not_issued')]
        tot = sum(x for x,_ in st)
        for x,h in sorted(st, reverse=True)[:8]:
            print('   stall %-45s %5.1f%%' % (h.replace('smsp__pcsamp_warps_issue_stalled_',''), 100*x/tot))
