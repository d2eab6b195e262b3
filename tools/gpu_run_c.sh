#!/bin/bash
# launch lists (ncu --metrics gpu__time_duration.sum) of one pooled and one attention train step, after a plain run each
mkdir -p gpurun_out
for v in pooled attention; do
  python tools/profile_step.py --variant $v --mode train --steps 1 --warmup 2 > gpurun_out/plain_$v.log 2>&1 || { echo "plain $v failed"; tail -5 gpurun_out/plain_$v.log; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_${v}_train_v2.csv \
      python tools/profile_step.py --variant $v --mode train --steps 1 --warmup 2 > gpurun_out/ncu_$v.log 2>&1
  echo "ncu $v rc=$?"
done
