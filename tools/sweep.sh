#!/bin/bash
# Kernel sweep of SURVEY.md 8d: rows in {128, 2048, 16384, 65536} x {eps, vel, vel_from_eps},
# plus the dense-VLB forward (configs[4]).  One JSON line per point -> gpurun_out/sweep.jsonl
out=gpurun_out/sweep.jsonl; : > $out
for p in eps vel vel_from_eps; do
  for r in 128 2048 16384 65536; do
    python bench.py --param $p --rows $r --steps 30 --warmup 3 --no-e2e --no-cpu-baseline >> $out 2>> gpurun_out/sweep.err
  done
done
python bench.py --workload dense_vlb --param vel_from_eps --rows 16384 --launch-rows 2048 --steps 30 --no-cpu-baseline >> $out 2>> gpurun_out/sweep.err
python bench.py --workload dense_vlb --param eps --rows 16384 --launch-rows 2048 --steps 30 --no-cpu-baseline >> $out 2>> gpurun_out/sweep.err
wc -l $out; tail -3 gpurun_out/sweep.err
