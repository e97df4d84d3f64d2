#!/bin/bash
# per-kernel timings for the three parameterisations at 16384 rows
for p in eps vel vel_from_eps; do
  python bench.py --param $p --steps 30 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$p', {k:(round(v['ms'],4), round(v.get('frac_of_measured',0),3)) for k,v in d['kernels'].items()}, round(d['value']), round(d['step_hbm']['frac_of_measured'],3))"
done
