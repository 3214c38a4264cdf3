#!/bin/bash
# A/B timing of the 6-DoF mixed kernel over tools/ab/<name>.so variants: tools/ab_6dof.sh name1 name2 ...
cd "$(dirname "$0")/.."
for r in $(seq 1 ${ROUNDS:-2}); do
  for n in "$@"; do
    echo "== $n (round $r)"
    MRPNP_LIB=$PWD/tools/ab/$n.so timeout 120 python tools/sixdof_report.py 0 8192 mixedonly 2>&1 | grep -E '"timing"|Error|error' | sed -E 's/.*"full": (true|false).*"ms": ([0-9.]+).*/\1 \2 ms/'
  done
done
