#!/bin/bash
for v in "$@"; do
  echo "== $v"
  FSS_LIB=$v python bench.py --workload c3 --steps 10 --warmup 5 --no-cpu --e2e-steps 1 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('c3 ms/tick %.2f  G/s %.2f  by_iter %s' % (d['ms_per_step'], d['value']/1e9, [round(x,2) for x in d['roofline']['launch_ms_by_iter']]))"
done
