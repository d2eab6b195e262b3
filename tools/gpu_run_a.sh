#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_gemm_tc.py tests/test_gpu_grouped.py -m gpu -q --tb=short -x > gpurun_out/r02_pytest_vocab.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_vocab.log
tail -n 15 gpurun_out/r02_pytest_vocab.log | cut -c1-300
timeout 300 python tools/bench_vocab.py > gpurun_out/r02_bench_vocab.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_bench_vocab.txt
grep -v "stg2=1\|astat=2" gpurun_out/r02_bench_vocab.txt | cut -c1-420
