#!/usr/bin/env python3
"""
Developer tool: condenses an `ncu --set full` report (read here, no GPU needed) into the
text summary committed under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_r1.ncu-rep > profiles/r1_step_kernel_ncu.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_static',
    'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
    'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
    'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__inst_executed_pipe_fp64.sum',
    'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_xu.sum',
    'sm__inst_executed_pipe_lsu.sum', 'sm__cycles_elapsed.max', 'gpc__cycles_elapsed.avg.per_second',
    'dram__cycles_elapsed.avg.per_second',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
]


def main(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print('# source: %s (ncu --set full --clock-control none; per-launch, cold-cache, serialised)' % path)
    for k, r in enumerate(rows[2:]):
        d = dict(zip(hdr, r))
        print('\n== launch %d: %s' % (k, d.get('Kernel Name', '?')))
        for key in KEYS:
            if key in d:
                print('%-82s %s %s' % (key, d[key], units[hdr.index(key)]))
        try:
            rd, wr = float(d['dram__bytes_read.sum']), float(d['dram__bytes_write.sum'])
            unit = units[hdr.index('dram__bytes_read.sum')]
            scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}.get(unit, 1)
            t = float(d['gpu__time_duration.sum']) * {'us': 1e-6, 'ms': 1e-3, 'ns': 1e-9, 's': 1}.get(units[hdr.index('gpu__time_duration.sum')], 1e-6)
            print('derived: DRAM traffic %.4f GB per launch, %.1f GB/s' % ((rd + wr) * scale / 1e9, (rd + wr) * scale / t / 1e9))
        except Exception:
            pass


if __name__ == '__main__':
    main(sys.argv[1])
