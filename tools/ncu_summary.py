#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): key raw metrics per kernel + top stall lines of the source page.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--source N]"""
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
        'l1tex__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'launch__shared_mem_per_block_dynamic', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_executed_pipe_fma.sum',
        'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_alu.sum', 'local_load_requests' ]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = set()
    for r in rows[2:]:
        name = r[idx['Kernel Name']]
        if name in seen:
            continue
        seen.add(name)
        print('=====', name)
        for w in WANT:
            if w in idx:
                print(f'  {w:82s} {r[idx[w]]} {units[idx[w]]}')
    if '--source' in sys.argv:
        n = int(sys.argv[sys.argv.index('--source') + 1])
        src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda'], capture_output=True, text=True).stdout
        blocks = src.split('\n\n')
        rows = list(csv.reader(io.StringIO(src)))
        # find header row
        for i, r in enumerate(rows):
            if '# Samples' in ' '.join(r) or 'Warp Stall Sampling (All Samples)' in r:
                hdr = r
                break
        else:
            print('no source header'); return
        col_s = next((j for j, h in enumerate(hdr) if h.startswith('Warp Stall Sampling (All')), None)
        col_src = next((j for j, h in enumerate(hdr) if h == 'Source'), None)
        col_i = next((j for j, h in enumerate(hdr) if h.startswith('Instructions Executed')), None)
        data = []
        for r in rows[i + 1:]:
            if len(r) != len(hdr):
                continue
            try:
                data.append((int(r[col_s] or 0), int(r[col_i] or 0) if col_i is not None else 0, r[0], r[col_src]))
            except ValueError:
                continue
        tot = sum(d[0] for d in data) or 1
        print(f'--- top {n} source lines by stall samples (total {tot}) ---')
        for s, ic, line, text in sorted(data, reverse=True)[:n]:
            print(f'{100.0 * s / tot:6.2f}%  inst={ic:>10}  L{line}: {text.strip()[:140]}')


if __name__ == '__main__':
    main()
