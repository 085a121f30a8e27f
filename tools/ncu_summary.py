#!/usr/bin/env python
"""Write a text summary of an .ncu-rep (per kernel: duration, instructions, IPC, occupancy, DRAM
bytes, top stall reasons, top source lines) for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/r01_ncu_summary.md
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ('gpu__time_duration.sum', 'duration'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('launch__registers_per_thread', 'regs/thread'),
    ('launch__shared_mem_per_block_dynamic', 'dyn smem/block'),
    ('launch__shared_mem_per_block_static', 'static smem/block'),
    ('launch__waves_per_multiprocessor', 'waves/SM'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
    ('smsp__inst_executed.sum', 'warp instructions'),
    ('sm__inst_executed.avg.per_cycle_elapsed', 'IPC (per SM, elapsed)'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput %'),
    ('dram__bytes_read.sum', 'dram bytes read'),
    ('dram__bytes_write.sum', 'dram bytes written'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit rate %'),
    ('l1tex__t_sector_hit_rate.pct', 'L1 hit rate %'),
]
STALLS = ['long_scoreboard', 'short_scoreboard', 'barrier', 'wait', 'branch_resolving', 'no_instruction',
          'not_selected', 'math_pipe_throttle', 'mio_throttle', 'lg_throttle', 'membar', 'dispatch_stall']


def main():
    rep = sys.argv[1]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(hdr)}
    print('# ncu summary of `%s`\n' % rep.split('/')[-1])
    print('Captured with `ncu --set full --clock-control none --import-source on` on one B200 (cold caches, '
          'serialised launches: compare shares, not absolutes).\n')
    for r in data:
        print('## %s\n' % r[col['Kernel Name']][:110])
        for m, label in METRICS:
            if m in col:
                print('- %s: %s %s' % (label, r[col[m]], units[col[m]]))
        st = []
        for s in STALLS:
            k = 'smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % s
            if k in col:
                try:
                    st.append((float(r[col[k]]), s))
                except ValueError:
                    pass
        st.sort(reverse=True)
        print('- warps stalled per issue (top): ' + ', '.join('%s %.2f' % (s, v) for v, s in st[:5]))
        print()
    print('## hottest source lines (share of stall samples / of executed warp instructions)\n')
    out = subprocess.run([sys.executable, __file__.replace('ncu_summary.py', 'ncu_lines.py'), rep, '', '12'],
                         capture_output=True, text=True).stdout
    print('```\n' + out + '```')


if __name__ == '__main__':
    main()
