#!/bin/bash
# ncu --set full of one fwd_pre launch per kernel shape (MULAN_FWD_PRE_V) and model; the raw
# metric pages are exported as CSV on the box (the .ncu-rep files stay within gpurun's 64 MiB).
#   bash tools/ncu_fwd_pre.sh "vel:0 vel:5 eps:5" -> gpurun_out/ncu_pre_<model>_v<V>.{ncu-rep,csv}
for mv in $1; do
  model=${mv%%:*}; v=${mv##*:}
  MULAN_FWD_PRE_V=$v ncu --set full --clock-control none -k regex:fwd_pre -s 2 -c 1 \
    -o gpurun_out/ncu_pre_${model}_v$v -f python tools/one_fwd_pre.py 16384 $model > gpurun_out/ncu_pre_${model}_v$v.log 2>&1
  ncu -i gpurun_out/ncu_pre_${model}_v$v.ncu-rep --page raw --csv > gpurun_out/ncu_pre_${model}_v$v.csv 2>/dev/null
done
ls -la gpurun_out/ncu_pre_*
