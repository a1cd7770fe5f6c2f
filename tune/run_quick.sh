#!/bin/bash
# usage: tune/run_quick.sh SIZE lib...   (3 timed ticks: enough to see the iter-0/1 time)
SIZE=$1; shift
for v in "$@"; do
  FSS_LIB=$v python bench.py --size $SIZE --steps 3 --warmup 3 --no-cpu --e2e-steps 0 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$v  ms/tick %.2f  by_iter %s' % (d['ms_per_step'], [round(x,2) for x in d['roofline']['launch_ms_by_iter']]))
"
done
