#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -x -k "gru_resident or gru_cluster" > gpurun_out/r02_pytest_gru.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_gru.log
tail -n 12 gpurun_out/r02_pytest_gru.log | cut -c1-300
timeout 300 python tools/bench_gru.py 2>&1 | tail -4
