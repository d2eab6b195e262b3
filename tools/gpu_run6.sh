#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_grouped_kernels.py tests/test_gpu_grouped.py tests/test_gpu_fullsize.py -q --tb=short -k "group or many" -s > gpurun_out/r02_pytest_grouped2.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_grouped2.log
python tools/bench_cc.py > gpurun_out/r02_bench_cc.txt 2>&1
python tools/profile_cc.py 100 > gpurun_out/plain_cc.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_cc_g100_v2.csv python tools/profile_cc.py 100 > gpurun_out/ncu_cc.log 2>&1
tail -n 12 gpurun_out/r02_pytest_grouped2.log | cut -c1-250; grep -v Warn gpurun_out/r02_bench_cc.txt | tail -9
