#!/bin/bash
# quick GPU iteration: parity tests + kernel timings (+ optional ncu of one kernel)
python -m pytest tests -m gpu -q --tb=line 2>&1 | tail -6
python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); print({k:(round(v['ms'],4), round(v.get('frac_of_measured',0),3)) for k,v in d['kernels'].items()}, round(d['value']), round(d['step_hbm']['frac_of_measured'],3), d['latency_b128']['us_per_step'], d['clocks'])"
tail -3 gpurun_out/bench_quick.err
if [ -n "$1" ]; then
ncu --set full --clock-control none --import-source on -k regex:"$1" -s 2 -c 1 -o gpurun_out/prof_quick -f python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_quick.log 2>&1
fi
