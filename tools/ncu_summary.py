#!/usr/bin/env python
"""Summarise an .ncu-rep (read here with `ncu -i`, no GPU needed): per captured launch the key
throughput / occupancy / stall numbers, plus the hottest SASS lines.  Used to write profiles/*.md."""
import csv
import io
import subprocess
import sys

KEYS = ['Grid Size', 'Block Size', 'gpu__time_duration.sum', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__average_warp_latency_per_inst_issued.ratio',
        'sm__cycles_elapsed.max']


def main():
    rep = sys.argv[1]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print('==', d.get('Kernel Name', '')[:60])
        for k in KEYS:
            if k in d:
                print(f'  {k} = {d[k]} {u[k]}')
        st = {h: float(d[h]) for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio') and d[h]}
        for h, v in sorted(st.items(), key=lambda x: -x[1])[:8]:
            print(f'  stall {h.split("stalled_")[1].split("_per_issue")[0]} = {v:.3f}')


if __name__ == '__main__':
    main()
