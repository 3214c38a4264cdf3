"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into the summary committed under profiles/.

    python tools/launch_summary.py gpurun_out/launches.csv profiles/r01b_launches "command line that was profiled"
"""
import csv, sys, collections, shutil
src, out, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else '')
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
ours = []
for r in rows:
    name, ns = r[ix['Kernel Name']], float(r[ix['Metric Value']])
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ns
    if 'mrpnp' in name: ours.append((name.split('(')[0].replace('void mrpnp::', ''), round(ns / 1e3), r[ix['Grid Size']]))
total = sum(v[1] for v in agg.values())
lines = [f'# ncu launch list of `{cmd}`',
         '# ncu --metrics gpu__time_duration.sum --clock-control none ; per-launch times are cold-cache and serialised',
         f'# total device time of all launches {total / 1e3:.1f} us', 'launches  total_us  share  kernel']
for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f'{n:8d} {ns / 1e3:9.1f} {100 * ns / total:6.2f}%  {name[:120]}')
lines.append('# launches of this repo\'s kernels in order (name, us, grid):')
lines.append('# ' + ' '.join(f'{n.split("<")[0]}:{t}' for n, t, g in ours))
open(out + '_summary.txt', 'w').write('\n'.join(lines) + '\n')
shutil.copy(src, out + '.csv')
print('\n'.join(lines[:12]))
