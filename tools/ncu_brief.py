"""Print a few decisive metrics per distinct kernel from an `ncu --page raw --csv` dump.  Usage: ncu_brief.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__pcsamp_warps_issue_stalled_long_scoreboard', 'smsp__pcsamp_warps_issue_stalled_lg_throttle',
        'smsp__pcsamp_warps_issue_stalled_math_pipe_throttle', 'smsp__pcsamp_warps_issue_stalled_mio_throttle', 'smsp__pcsamp_warps_issue_stalled_short_scoreboard',
        'smsp__pcsamp_warps_issue_stalled_barrier', 'smsp__pcsamp_warps_issue_stalled_wait', 'smsp__pcsamp_warps_issue_stalled_not_selected',
        'smsp__pcsamp_warps_issue_stalled_selected', 'smsp__pcsamp_warps_issue_stalled_membar', 'smsp__pcsamp_warps_issue_stalled_drain', 'smsp__pcsamp_warps_issue_stalled_sleeping',
        'smsp__pcsamp_warps_issue_stalled_dispatch_stall', 'smsp__pcsamp_warps_issue_stalled_no_instructions', 'smsp__pcsamp_warps_issue_stalled_imc_miss', 'smsp__pcsamp_warps_issue_stalled_tex_throttle', 'smsp__pcsamp_warps_issue_stalled_branch_resolving']
seen = {}
for r in data:
    name = r[idx['Kernel Name']]
    key = name[:90] + r[idx['launch__grid_size']]
    if key in seen:
        continue
    seen[key] = 1
    print('==', name[:110])
    for w in want:
        if w in idx and r[idx[w]] not in ('', '0'):
            print(f'   {w:80s} {r[idx[w]]:>16s} {units[idx[w]]}')
