"""Opcode evidence for the "Blackwell-native" claim: per kernel of the two in-tree libraries, how often the SASS shows the
tcgen05 / TMEM / TMA / bulk-copy / packed-fp32 instructions (cuobjdump -sass; B200_PROFILING.md "What proves a
Blackwell-native kernel").  Writes profiles/sass_opcodes_{pnp,head}.txt."""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = re.compile(r'^(UTC[A-Z]*MMA\S*|LDTM\S*|STTM\S*|UTMALDG\S*|UTMASTG\S*|UBLKCP\S*|UTCBAR\S*|SYNCS\S*|FFMA2|FADD2|FMUL2|CREDUX\S*|HMMA\S*|DFMA|'
                 r'MUFU\.\S+|ATOMG\S*|REDG\S*|LDS\.64|LDS\.128|STS\.64|UTCCP\S*|UTCSHIFT\S*)$')
for lib in ('pnp', 'head'):
    sass = subprocess.run(['cuobjdump', '-sass', os.path.join(ROOT, 'monorun_b200', f'libmonorun_{lib}.so')], capture_output=True, text=True).stdout
    kernels, cur = [], None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = [m.group(1), 0, collections.Counter()]
            kernels.append(cur)
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(.*?);', line)
        if m and cur:
            cur[1] += 1
            for tok in m.group(1).replace(',', ' ').split():
                if PAT.match(tok):
                    cur[2][tok.rstrip(';')] += 1
    demangle = subprocess.run(['c++filt'] + [k[0] for k in kernels], capture_output=True, text=True).stdout.splitlines()
    with open(os.path.join(ROOT, 'profiles', f'sass_opcodes_{lib}.txt'), 'w') as f:
        f.write(f'# cuobjdump -sass monorun_b200/libmonorun_{lib}.so (nvcc -gencode arch=compute_100a,code=sm_100a); per kernel: SASS\n'
                '# instructions in total, then counts of the Blackwell / async-copy / packed-fp32 / fp64 opcodes found\n')
        for (name, total, c), dm in zip(kernels, demangle):
            f.write(f'\n{dm}  ({total} instructions)\n')
            for op, n in sorted(c.items()):
                f.write(f'    {n:6d} {op}\n')
    print(lib, len(kernels), 'kernels')
