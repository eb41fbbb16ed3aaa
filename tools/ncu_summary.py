"""Key metrics of an ncu --set full report: python tools/ncu_summary.py file.ncu-rep [...]"""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed',
        'l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed_op_shared_atom.sum', 'smsp__inst_executed_op_global_red.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
for fn in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', fn, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, U = rows[0], rows[1]
    for V in rows[2:]:
        print('Kernel Name'.ljust(70), V[H.index('Kernel Name')])
        for w in WANT:
            if w in H:
                print(w.ljust(70), V[H.index(w)], U[H.index(w)])
        st = []
        for i, h in enumerate(H):
            if 'warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio') and 'not_issued' not in h:
                try:
                    st.append((float(V[i]), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
                except ValueError:
                    pass
        print('stall reasons (warps per issue-active cycle): ' + ', '.join(f'{n}={v:.2f}' for v, n in sorted(st, reverse=True)[:9]))
        print()
