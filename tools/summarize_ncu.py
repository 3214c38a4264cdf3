"""Turn an .ncu-rep into the small text summary committed under profiles/ (raw metrics + per-line hot spots).

    python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01_name
"""
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__bytes_read.sum.per_second', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_active.avg.per_cycle_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'smsp__thread_inst_executed_per_inst_executed.ratio']
lines, summary = [], []
for r in rows[2:]:
    d = {h: v for h, v in zip(hdr, r)}
    u = {h: x for h, x in zip(hdr, units)}
    rec = {}
    lines.append('## ' + d.get('Kernel Name', '?'))
    for k in KEYS[1:]:
        if k in d:
            lines.append(f'{k:80s} {d[k]:>18s} {u[k]}')
            rec[k] = d[k]
    stalls = sorted(((float(v), h) for h, v in d.items() if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and v),
                    reverse=True)
    lines.append('warp stall reasons (warps per issue-active cycle):')
    for v, h in stalls[:8]:
        lines.append(f'    {v:6.3f}  ' + h.split('issue_stalled_')[1].split('_per_issue')[0])
    if 'dram__bytes_read.sum' in d:
        mul = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}
        rec['dram_bytes'] = float(d['dram__bytes_read.sum']) * mul.get(u['dram__bytes_read.sum'], 1) + \
            float(d['dram__bytes_write.sum']) * mul.get(u['dram__bytes_write.sum'], 1)
        lines.append(f'dram read+write bytes per launch: {rec["dram_bytes"]:.0f}')
    summary.append(rec)
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
agg, cur, ts, ti = [], None, 0, 0
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if len(r) >= 8 and r[0].strip().isdigit():
        try:
            s, n = int(r[4]), int(r[7])
        except ValueError:
            continue
        agg.append((s, n, cur, int(r[0]), r[1].strip()[:100]))
        ts += s
        ti += n
if ts:
    lines.append(f'\n## hottest source lines (of {ts} stall samples, {ti} warp instructions)')
    for s, n, f, l, t in sorted(agg, reverse=True)[:25]:
        lines.append(f'{100 * s / ts:6.2f}% samples {100 * n / max(ti, 1):6.2f}% inst  {f}:{l}  {t}')
open(out + '_summary.txt', 'w').write('\n'.join(lines) + '\n')
json.dump(summary, open(out + '_metrics.json', 'w'), indent=1)
print('\n'.join(lines[:60]))
