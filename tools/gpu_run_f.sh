#!/bin/bash
# round-2 (second session) evidence: launch lists + ncu --set full of the new kernels, each after a plain run exited 0
mkdir -p gpurun_out
for v in pooled attention; do
  python tools/profile_step.py --variant $v --mode train --steps 1 --warmup 2 > gpurun_out/plain_$v.log 2>&1 || { echo "plain $v failed"; tail -5 gpurun_out/plain_$v.log; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_${v}_train_v3.csv \
      python tools/profile_step.py --variant $v --mode train --steps 1 --warmup 2 > gpurun_out/ncu_$v.log 2>&1
  echo "ncu list $v rc=$?"
done
ncu --set full --clock-control none --import-source on -k regex:"ce_fwd_split_kernel|gru_res_fwd_kernel|gru_res_bwd_kernel|gemm_tc_kernel" -s 6 -c 16 -f -o gpurun_out/prof_pooled_v3 \
    python tools/profile_step.py --variant pooled --mode train --steps 1 --warmup 2 > gpurun_out/ncu_full_pooled.log 2>&1; echo "ncu full rc=$?"
timeout 300 python tools/bench_vocab.py > gpurun_out/r02_bench_vocab.txt 2>&1
timeout 300 python tools/bench_gru.py > gpurun_out/r02_bench_gru.txt 2>&1
timeout 300 python tools/bench_decode_parts.py > gpurun_out/r02_bench_decode_parts.txt 2>&1
timeout 300 python tools/bench_decode.py r02-session2 > gpurun_out/r02_bench_decode_v2.txt 2>&1
ls -la gpurun_out/prof_pooled_v3.ncu-rep; tail -2 gpurun_out/ncu_full_pooled.log
