"""A/B timing of libmonorun_pnp.so variants (tools/ab/*.so) inside ONE GPU session: each variant runs in its own
process (MRPNP_LIB), interleaved over several rounds so that box-to-box and thermal drift cancel."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
names = sys.argv[1].split(',')
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
bands = sys.argv[3] if len(sys.argv) > 3 else '0:0:0'
res = {n: {} for n in names}
for r in range(rounds):
    for n in names:
        env = dict(os.environ, MRPNP_LIB=os.path.join(ROOT, 'tools', 'ab', n + '.so'))
        out = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'band_timing.py'), '8192', bands, 'fastonly'], env=env,
                             capture_output=True, text=True)
        for l in out.stdout.splitlines():
            try:
                d = json.loads(l)
            except Exception:
                continue
            res[n].setdefault(d['workload'], []).append((d['us_mean'], d['us_min']))
        if out.returncode:
            print(n, 'FAILED', out.stderr[-500:])
for n in names:
    for w, v in res[n].items():
        print(f'{n:24s} {w:5s} mean {sum(a for a, _ in v) / len(v):7.1f} us  min {min(b for _, b in v):7.1f} us   rounds {[round(a, 1) for a, _ in v]}', flush=True)
