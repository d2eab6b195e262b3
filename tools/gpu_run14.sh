#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_ops.py tests/test_gpu_grouped_kernels.py tests/test_gpu_fullsize.py tests/test_gpu_pooled.py -q --tb=short -x > gpurun_out/r02_pytest_split.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_split.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench_rc=$?" >> gpurun_out/r02_bench_n1.err
tail -n 4 gpurun_out/r02_pytest_split.log | cut -c1-300; tail -n 1 gpurun_out/r02_bench_n1.err; cut -c1-200 gpurun_out/r02_bench_n1.json
