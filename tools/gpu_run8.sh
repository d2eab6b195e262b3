#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_grouped.py tests/test_gpu_fullsize.py tests/test_gpu_pooled.py -q --tb=short -k "grouped or pooled" -s > gpurun_out/r02_pytest_pooled_grouped.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_pooled_grouped.log
tail -n 25 gpurun_out/r02_pytest_pooled_grouped.log | cut -c1-300
