"""Per-source-line stall samples and executed instructions from an ncu report captured with --import-source on:
    python tools/ncu_lines.py report.ncu-rep [top_n]
(ncu --page source --print-source cuda,sass: rows whose Address is '-' are the per-line aggregates.)"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
path = None
hdr = None
agg = {}
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        path = r[1].split('/')[-1]; continue
    if len(r) >= 2 and r[0] == 'Function Name':
        continue
    if len(r) > 4 and r[0] == 'Line No':
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[2] != '-':
        continue
    d = dict(zip(hdr, r))
    try:
        line = int(r[0])
    except ValueError:
        continue
    smp = int(d['# Samples'] or 0); ins = int(d['Instructions Executed'] or 0)
    key = (path, line)
    a = agg.setdefault(key, [0, 0, r[1].strip()[:90], collections.Counter()])
    a[0] += smp; a[1] += ins
    for k, v in d.items():
        if k.startswith('stall_') and 'Not Issued' not in k and v not in ('', '0', '-'):
            a[3][k[6:]] += int(v)
tot_s = sum(a[0] for a in agg.values()); tot_i = sum(a[1] for a in agg.values())
print(f'total samples {tot_s}, warp instructions {tot_i}')
byfile = collections.Counter(); byfile_i = collections.Counter()
for (p, l), a in agg.items():
    byfile[p] += a[0]; byfile_i[p] += a[1]
for p, s in byfile.most_common():
    print(f'  {p:28s} samples {100 * s / tot_s:5.1f} %   instructions {100 * byfile_i[p] / tot_i:5.1f} %')
for (p, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ' '.join(f'{k}:{v}' for k, v in a[3].most_common(3))
    print(f'{100 * a[0] / tot_s:5.2f} % smp {100 * a[1] / tot_i:5.2f} % ins  {p}:{l:<5d} {a[2]:90s} {st}')
# regions of pnp_kernel_fast.cuh / totals per stall reason
if len(sys.argv) > 3:
    lo, hi = (int(v) for v in sys.argv[3].split('-'))
    c = collections.Counter(); s = i = 0
    for (p, l), a in agg.items():
        if p == 'pnp_kernel_fast.cuh' and lo <= l <= hi:
            c.update(a[3]); s += a[0]; i += a[1]
    print(f'pnp_kernel_fast.cuh:{lo}-{hi}: samples {100 * s / tot_s:.1f} %, instructions {100 * i / tot_i:.1f} %', dict(c.most_common(8)))
allc = collections.Counter()
for a in agg.values():
    allc.update(a[3])
print('all (inlined frames counted once per frame):', dict(allc.most_common(10)))
