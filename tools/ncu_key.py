"""Key metrics per kernel from an .ncu-rep: python tools/ncu_key.py file.ncu-rep"""
import csv, subprocess, sys, io
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum.per_second', 'lts__t_bytes.sum', 'sm__cycles_active.avg', 'sm__cycles_elapsed.avg',
        'smsp__issue_active.avg.pct', 'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__cycles_active.avg', 'sm__inst_executed_pipe_fma', 'sm__inst_executed_pipe_alu', 'sm__inst_executed_pipe_lsu',
        'sm__inst_executed_pipe_xu', 'sm__pipe_fma_cycles_active.avg.pct', 'sm__pipe_alu_cycles_active.avg.pct', 'smsp__average_warp', 'launch__registers_per_thread',
        'launch__occupancy_limit', 'sm__maximum_warps_per_active_cycle_pct', 'launch__grid_size', 'launch__block_size', 'smsp__warp_issue_stalled', 'smsp__average_warps_issue_stalled']
for r in rows[2:]:
    print('=====', r[hdr.index('Kernel Name')][:90])
    for i, h in enumerate(hdr):
        if any(h.startswith(w) for w in want) and r[i] not in ('', '0'):
            if 'stalled' in h and not h.endswith('_per_warp_active.pct'):
                continue
            print(f'  {h} = {r[i]} {units[i]}')
