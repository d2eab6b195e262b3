#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/r02_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench_rc=$?" >> gpurun_out/r02_bench_n1.err
python tools/profile_cc.py 100 > gpurun_out/plain_cc.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_cc_g100.csv python tools/profile_cc.py 100 > gpurun_out/ncu_cc.log 2>&1
CAPHN_NO_GRAPH=1 python tools/bench_decode.py > gpurun_out/plain_dec.log 2>&1 &&
CAPHN_NO_GRAPH=1 CAPHN_DECODE_ITERS=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_decode.csv python tools/bench_decode.py > gpurun_out/ncu_dec.log 2>&1
python tools/bench_decode.py > gpurun_out/r02_bench_decode.txt 2>&1
tail -n 6 gpurun_out/r02_pytest_gpu.log | cut -c1-300; tail -n 2 gpurun_out/r02_bench_n1.err; cat gpurun_out/plain_cc.log | tail -2; tail -5 gpurun_out/r02_bench_decode.txt
