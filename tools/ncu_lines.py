"""Summarise an ncu `--page source --print-source cuda,sass --csv` dump per CUDA source line."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
agg = []
tot_s = tot_i = 0
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]; continue
    if len(r) < 8 or r[0] in ('Line No', 'Function Name'):
        continue
    if r[0].strip().isdigit():
        try:
            s = int(r[4]); n = int(r[7])
        except ValueError:
            continue
        agg.append((s, n, cur_file, int(r[0]), r[1].strip()[:110]))
        tot_s += s; tot_i += n
print(f'total samples {tot_s}  total warp-instructions {tot_i}')
print('--- by stall samples')
for s, n, f, l, src in sorted(agg, reverse=True)[:top]:
    print(f'{100*s/tot_s:6.2f}%  inst {100*n/tot_i:6.2f}%  {f}:{l}  {src}')
if len(sys.argv) > 3:
    # group by (file, line-range) buckets given as file:lo-hi=name
    buckets = []
    for spec in sys.argv[3:]:
        rng, name = spec.split('=')
        f, lr = rng.split(':'); lo, hi = map(int, lr.split('-'))
        buckets.append((f, lo, hi, name))
    res = collections.OrderedDict((b[3], [0, 0]) for b in buckets); res['other'] = [0, 0]
    for s, n, f, l, src in agg:
        for bf, lo, hi, name in buckets:
            if f == bf and lo <= l <= hi:
                res[name][0] += s; res[name][1] += n; break
        else:
            res['other'][0] += s; res['other'][1] += n
    print('--- buckets')
    for k, (s, n) in res.items():
        print(f'{k:28s} samples {100*s/tot_s:6.2f}%   inst {100*n/tot_i:6.2f}%')
