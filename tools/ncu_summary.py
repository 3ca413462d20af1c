"""Condense `ncu --set full` reports into one small CSV (last captured launch per kernel) for profiles/.

usage: python tools/ncu_summary.py out.csv a.ncu-rep [b.ncu-rep ...]
"""
import csv
import io
import subprocess
import sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size', 'launch__cluster_size',
        'launch__shared_mem_per_block_dynamic', 'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.max']
rows_out = {}
for rep in sys.argv[2:]:
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(hdr)}
    for r in body:
        name = r[col['Kernel Name']].split('(')[0].replace('void ', '').strip()
        rows_out[name] = [(k, r[col[k]] if k in col else '', units[col[k]] if k in col else '') for k in KEEP]
with open(sys.argv[1], 'w') as fh:
    wr = csv.writer(fh)
    names = list(rows_out)
    wr.writerow(['metric', 'unit'] + names)
    for i, k in enumerate(KEEP):
        wr.writerow([k, rows_out[names[0]][i][2]] + [rows_out[n][i][1] for n in names])
print(open(sys.argv[1]).read())
