#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/r02_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench_rc=$?" >> gpurun_out/r02_bench_n1.err
tail -n 6 gpurun_out/r02_pytest_gpu.log | cut -c1-300; tail -n 2 gpurun_out/r02_bench_n1.err; cut -c1-200 gpurun_out/r02_bench_n1.json
