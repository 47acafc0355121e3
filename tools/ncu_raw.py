"""Key metrics of an .ncu-rep: python tools/ncu_raw.py file.ncu-rep"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'launch__grid_size', 'launch__block_size',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active']
stalls = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
for r in rows[2:]:
    print('----', r[hdr.index('Kernel Name')][:70])
    for w in want:
        if w in hdr:
            i = hdr.index(w); print(f'  {w[:60]:60s} {r[i][:20]} {units[i]}')
    st = sorted(((float(r[hdr.index(h)] or 0), h[34:-24]) for h in stalls), reverse=True)[:6]
    print('  stalls/issue:', ' '.join(f'{n}={v:.2f}' for v, n in st))
