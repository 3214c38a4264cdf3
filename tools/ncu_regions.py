"""Scratch: aggregate an ncu source page by source-line regions (instructions + stall samples per phase of the kernel).
usage: ncu_regions.py report.ncu-rep"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
cur = None
per = collections.defaultdict(lambda: [0, 0])
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]; continue
    if len(r) >= 8 and r[0].strip().isdigit():
        try: s, n = int(r[4]), int(r[7])
        except ValueError: continue
        per[(cur, int(r[0]))][0] += s; per[(cur, int(r[0]))][1] += n
ts = sum(v[0] for v in per.values()); ti = sum(v[1] for v in per.values())
import re, os
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'monorun_b200', 'csrc')
# regions = enclosing function (by scanning the source for function heads)
def regions(fn):
    out = []; path = os.path.join(root, fn)
    if not os.path.exists(path): return out
    name = None
    for i, l in enumerate(open(path), 1):
        m = re.match(r'^(?:template.*\n)?\s*(?:__device__|__global__|static|inline).*?(\w+)\s*\(', l)
        if m and not l.startswith(' '): name = m.group(1)
        m2 = re.match(r'\s*// -{8,} (.*?) -{8,}', l)
        if m2: name = (name or '') .split('|')[0] + '|' + m2.group(1)
        out.append(name)
    return out
cache = {}
agg = collections.defaultdict(lambda: [0, 0])
for (f, l), (s, n) in per.items():
    if f not in cache: cache[f] = regions(f)
    rg = cache[f][l - 1] if l - 1 < len(cache[f]) else None
    agg[(f, rg)][0] += s; agg[(f, rg)][1] += n
print(f'total stall samples {ts}, warp instructions {ti}')
for (f, rg), (s, n) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{100*n/ti:6.2f}% inst {100*s/ts:6.2f}% samples  {f}: {rg}')
