"""Scratch: static SASS size of a kernel by source region (nvdisasm --print-line-info output).
usage: code_size.py all.dis kernel_mangled_substring"""
import re, sys, collections, os
dis, key = sys.argv[1], sys.argv[2]
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'monorun_b200', 'csrc')
def regions(fn):
    out = []; path = os.path.join(root, fn)
    if not os.path.exists(path): return out
    name = None
    for l in open(path):
        m = re.match(r'^\s*(?:__device__|__global__|static|inline).*?(\w+)\s*\(', l)
        if m and not l.startswith(' '): name = m.group(1)
        m2 = re.match(r'\s*// -{8,} (.*?) -{8,}', l)
        if m2: name = (name or '').split('|')[0] + '|' + m2.group(1)
        out.append(name)
    return out
cache = {}
inside = False; cur = ('?', 0); cnt = collections.Counter(); total = 0
for l in open(dis):
    if l.startswith('//---') and '.text.' in l:
        inside = key in l
        continue
    if not inside: continue
    m = re.match(r'\s*//## File "(.*?)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s*/\*[0-9a-f]{4,}\*/', l):
        f, ln = cur
        if f not in cache: cache[f] = regions(f)
        rg = cache[f][ln - 1] if 0 < ln <= len(cache[f]) else None
        cnt[(f, rg)] += 1; total += 1
print(f'{total} SASS instructions = {total * 16 / 1024:.1f} KB')
for (f, rg), n in cnt.most_common(40):
    print(f'{n:6d} {n * 16 / 1024:6.1f} KB  {f}: {rg}')
