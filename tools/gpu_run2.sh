#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_grouped_kernels.py tests/test_gpu_grouped.py tests/test_gpu_gemm_tc.py -q --tb=short -x > gpurun_out/r02_pytest_grouped.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_grouped.log
python -m pytest tests/test_gpu_fullsize.py -q --tb=short -s > gpurun_out/r02_pytest_fullsize.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_fullsize.log
python -m pytest tests/test_gpu_attention.py tests/test_gpu_edge_cases.py -q --tb=short > gpurun_out/r02_pytest_att.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_att.log
tail -n 25 gpurun_out/r02_pytest_grouped.log; tail -n 40 gpurun_out/r02_pytest_fullsize.log; tail -n 5 gpurun_out/r02_pytest_att.log
