#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_attention.py tests/test_gpu_bf16.py tests/test_gpu_grouped_kernels.py tests/test_gpu_grouped.py -q --tb=short -k "beam or bf16 or grouped" -s > gpurun_out/r02_pytest_new.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_new.log
python tools/bench_cc.py > gpurun_out/r02_bench_cc.txt 2>&1
python tools/bench_gemm.py > gpurun_out/r02_gemm_bench.txt 2>&1
tail -n 30 gpurun_out/r02_pytest_new.log | cut -c1-250; grep -v Warn gpurun_out/r02_bench_cc.txt | tail -12; tail -n 14 gpurun_out/r02_gemm_bench.txt
