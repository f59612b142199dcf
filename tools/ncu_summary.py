"""Print the key metrics of an .ncu-rep (raw page) -- used to write profiles/*.md."""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_lsu.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_red.sum', 'lts__t_sectors_op_red.sum', 'lts__t_sectors.sum',
        'lts__t_bytes.sum', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'sm__cycles_active.avg', 'launch__occupancy_limit_registers',
        'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum', 'smsp__inst_executed_op_shfl... ']
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    print('== kernel:', vals[hdr.index('Kernel Name')][:90])
    for i, h in enumerate(hdr):
        show = h in WANT
        if 'warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio') and 'not_issued' not in h:
            try:
                show = float(vals[i].replace(',', '')) >= 0.25
            except ValueError:
                show = False
        if show:
            print(f'  {h} [{units[i]}] = {vals[i]}')
