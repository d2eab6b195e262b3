#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_grouped.py tests/test_gpu_grouped_kernels.py tests/test_gpu_edge_cases.py tests/test_gpu_bf16.py -m gpu -q --tb=short > gpurun_out/r02_pytest_att2.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_att2.log
tail -n 4 gpurun_out/r02_pytest_att2.log | cut -c1-200
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench_rc=$?"
cut -c1-200 gpurun_out/r02_bench_n1.json
