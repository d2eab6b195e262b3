#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_pooled.py tests/test_gpu_grouped.py tests/test_gpu_ops.py -q --tb=short > gpurun_out/r02_pytest_dec.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_dec.log
python tools/bench_decode.py > gpurun_out/r02_bench_decode.txt 2>&1
python tools/bench_cc.py > gpurun_out/r02_bench_cc.txt 2>&1
tail -n 4 gpurun_out/r02_pytest_dec.log | cut -c1-300; grep -v Warn gpurun_out/r02_bench_decode.txt | tail -5; grep -v Warn gpurun_out/r02_bench_cc.txt | tail -4
