#!/usr/bin/env python
"""Summarise an .ncu-rep per CUDA source line: instructions executed, stall samples, top stalls.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [kernel-substring] [top-N]
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    filt = sys.argv[2] if len(sys.argv) > 2 else ''
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda'],
                         capture_output=True, text=True).stdout
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(out)):
        if not row:
            continue
        if row[0] == 'File Path':
            cur = {'file': row[1], 'func': '', 'hdr': None, 'rows': []}
            blocks.append(cur)
        elif row[0] == 'Function Name' and cur is not None:
            cur['func'] = row[1]
        elif row[0] == 'Line No' and cur is not None:
            cur['hdr'] = row
        elif cur is not None and cur['hdr'] is not None and len(row) == len(cur['hdr']):
            cur['rows'].append(row)
    by_func = {}
    for b in blocks:
        by_func.setdefault(b['func'], []).append(b)
    for func, bl in by_func.items():
        if filt not in func:
            continue
        lines = []
        for b in bl:
            h = b['hdr']
            ii, si = h.index('Instructions Executed'), h.index('# Samples')
            stall_cols = [(k, n) for k, n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n]
            for r in b['rows']:
                if not r[0]:
                    continue          # SASS rows; the CUDA-line rows carry the aggregate
                try:
                    inst, samp = int(r[ii]), int(r[si])
                except ValueError:
                    continue
                st = sorted(((int(r[k] or 0), n) for k, n in stall_cols), reverse=True)[:3]
                lines.append((samp, inst, b['file'].split('/')[-1], r[0], r[1].strip()[:90],
                              ' '.join('%s=%d' % (n[6:], v) for v, n in st if v)))
        tot_i = sum(l[1] for l in lines) or 1
        tot_s = sum(l[0] for l in lines) or 1
        print('== %s\n   total warp-instructions %d, samples %d' % (func[:100], tot_i, tot_s))
        for samp, inst, f, ln, src, st in sorted(lines, reverse=True)[:topn]:
            print('%5.1f%% smp %5.1f%% ins  %s:%s  %s   [%s]' % (100.0 * samp / tot_s, 100.0 * inst / tot_i, f, ln, src, st))


if __name__ == '__main__':
    main()
