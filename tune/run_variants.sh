#!/bin/bash
# usage: tune/run_variants.sh SIZE lib1.so lib2.so ...   -- times bench.py per library variant (kernel tuning only)
SIZE=$1; shift
for v in "$@"; do
  echo "== $v"
  FSS_LIB=$v python bench.py --size $SIZE --steps 5 --warmup 3 --no-cpu --e2e-steps 1 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/tick %.2f  G/s %.2f  by_iter %s  e2e ms %.1f' % (d['ms_per_step'], d['value']/1e9, [round(x,2) for x in d['roofline']['launch_ms_by_iter']], d['e2e']['ms_per_step']))
    elif 'Error' in l or 'error' in l: print(l.strip())
"
done
