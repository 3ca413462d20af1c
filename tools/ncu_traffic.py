"""Parse DRAM bytes per launch out of `ncu --set full` reports and write profiles/r02_ncu_traffic.json.

usage: python tools/ncu_traffic.py gpurun_out/<capture>.ncu-rep [more.ncu-rep ...]
Reads the report with `ncu -i <rep> --page raw --csv`; for every kernel name keeps the LAST captured launch
(warm-up launches come first) and records dram__bytes_read.sum + dram__bytes_write.sum in bytes.
"""
import csv
import io
import json
import os
import subprocess
import sys

UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
per, dur = {}, {}
for rep in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(hdr)}
    for r in body:
        name = r[col['Kernel Name']].split('(')[0].split('<')[0].replace('void ', '').strip()
        name = name.split('::')[-1]
        tot = 0.0
        for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            tot += float(r[col[m]].replace(',', '')) * UNIT[units[col[m]]]
        per[name] = tot
        dur[name] = r[col['gpu__time_duration.sum']] + ' ' + units[col['gpu__time_duration.sum']]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dst = os.path.join(root, 'profiles', 'r02_ncu_traffic.json')
json.dump({'source': ', '.join(os.path.basename(r) for r in sys.argv[1:]) +
                     ' (ncu --set full --clock-control none, tools/run_once.py --workload cfg2, B=4096 T=48 H=256 bf16; '
                     'last captured launch per kernel)',
           'dram_bytes_per_launch': per, 'ncu_duration': dur}, open(dst, 'w'), indent=1)
print(json.dumps(per, indent=1))
