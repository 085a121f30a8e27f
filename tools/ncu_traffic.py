#!/usr/bin/env python
"""DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of every kernel in an
.ncu-rep -> JSON for bench.py's roofline.traffic.

    python tools/ncu_traffic.py gpurun_out/prof.ncu-rep > profiles/r01_traffic.json
"""
import csv
import io
import json
import re
import subprocess
import sys

UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def main():
    rep = sys.argv[1]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(hdr)}
    out = {'source': rep.split('/')[-1], 'how': 'ncu --set full --clock-control none, one launch per kernel',
           'kernels': {}}
    for r in data:
        name = r[col['Kernel Name']].split('(FusedNmsArgs')[0].split('(const')[0]
        name = re.sub(r'\(int\)|\(bool\)|rd::', '', re.sub(r'^void\s+', '', name)).strip()     # keeps the template arguments
        while name in out['kernels']:
            name += "'"
        rd = float(r[col['dram__bytes_read.sum']]) * UNIT[units[col['dram__bytes_read.sum']]]
        wr = float(r[col['dram__bytes_write.sum']]) * UNIT[units[col['dram__bytes_write.sum']]]
        dur = float(r[col['gpu__time_duration.sum']])
        out['kernels'][name] = {'dram_read_bytes': rd, 'dram_write_bytes': wr, 'traffic_bytes': rd + wr,
                                'duration_us_under_ncu': dur if units[col['gpu__time_duration.sum']] == 'us' else dur / 1e3}
    out['stage_traffic_bytes'] = sum(k['traffic_bytes'] for k in out['kernels'].values())
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
