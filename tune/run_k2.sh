#!/bin/bash
SIZE=$1; shift
for v in "$@"; do
  echo "== $v"
  FSS_LIB=$v python bench.py --workload c3 --size $SIZE --steps 4 --warmup 3 --no-cpu --e2e-steps 1 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['roofline_k2']; print('K2 ms %.3f  GB/s %.0f  frac %.3f | tick ms %.2f' % (k['launch_ms'], k['achieved'], k['frac'], d['ms_per_step']))
"
done
