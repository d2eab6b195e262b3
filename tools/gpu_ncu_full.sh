#!/bin/bash
# ncu --set full captures of the top kernels (each after a plain run of the same command exited 0)
mkdir -p gpurun_out
python tools/profile_step.py --variant pooled --mode train --steps 1 --warmup 2 > gpurun_out/plain_p.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"rows_bwd_kernel|rows_fwd_kernel" -s 20 -c 6 -f -o gpurun_out/prof_rows \
    python tools/profile_step.py --variant pooled --mode train --steps 1 --warmup 2 > gpurun_out/ncu_rows.log 2>&1
python tools/profile_step.py --variant attention --mode train --steps 1 --warmup 2 > gpurun_out/plain_a.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel" -s 38 -c 10 -f -o gpurun_out/prof_gemm \
    python tools/profile_step.py --variant attention --mode train --steps 1 --warmup 2 > gpurun_out/ncu_gemm.log 2>&1
python tools/profile_cc.py 100 > gpurun_out/plain_cc.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"attstep_gates_kernel|attbwd_gemm_kernel|gemm_tc_kernel" -s 100 -c 12 -f -o gpurun_out/prof_grouped \
    python tools/profile_cc.py 100 > gpurun_out/ncu_grouped.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/ncu_rows.log gpurun_out/ncu_gemm.log gpurun_out/ncu_grouped.log
