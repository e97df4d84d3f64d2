#!/bin/bash
# Kernel sweep of SURVEY.md 8d: rows in {128, 2048, 16384, 65536} x {eps, vel, vel_from_eps},
# the dense-VLB forward (configs[4]), and the A/B variants kept behind environment switches.
# One JSON line per point -> gpurun_out/sweep.jsonl ; render with profiles/sweep_table.py
out=gpurun_out/sweep.jsonl; : > $out
run() {  # note, env assignments..., -- bench args
  note=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" python bench.py "$@" --steps 30 --warmup 3 --no-cpu-baseline --no-configs --no-sustained 2>> gpurun_out/sweep.err \
    | python -c "import json,sys; d=json.loads(sys.stdin.read()); d['note']='$note'; print(json.dumps(d))" >> $out
}
for p in eps vel vel_from_eps; do
  for r in 128 2048 16384 65536; do
    run "" -- --param $p --rows $r --no-e2e
  done
done
run "literal v-from-eps formula" MULAN_VFE_LITERAL=1 -- --param vel_from_eps --rows 16384 --no-e2e
run "fwd_pre 256 threads x 4 CTAs/SM (round-1 shape)" MULAN_FWD_PRE_V=0 -- --param eps --rows 16384 --no-e2e
run "fwd_pre 256 threads x 4 CTAs/SM (round-1 shape)" MULAN_FWD_PRE_V=0 -- --param vel --rows 16384 --no-e2e
run "plain launches (no PDL)" -- --param eps --rows 16384 --no-e2e --no-pdl
run "separate post passes + stand-alone reduction" -- --param eps --rows 16384 --no-e2e --separate-post
run "fwd_pre generic constants" MULAN_NO_BAKED=1 -- --param eps --rows 16384 --no-e2e
run "fwd_pre TMA pipeline" MULAN_FWD_PRE_TMA=1 -- --param eps --rows 16384 --no-e2e
run "fwd_pre TMA pipeline, generic constants" MULAN_FWD_PRE_TMA=1 MULAN_NO_BAKED=1 -- --param eps --rows 16384 --no-e2e
wc -l $out; tail -3 gpurun_out/sweep.err
