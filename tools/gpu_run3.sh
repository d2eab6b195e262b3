#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/r02_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench_rc=$?" >> gpurun_out/r02_bench_n1.err
python tools/bench_gemm.py > gpurun_out/r02_gemm_bench.txt 2>&1
python tools/bench_sweep.py > gpurun_out/r02_sweep.txt 2> gpurun_out/r02_sweep.err
tail -n 8 gpurun_out/r02_pytest_gpu.log; tail -n 3 gpurun_out/r02_bench_n1.err; tail -n 12 gpurun_out/r02_gemm_bench.txt; cat gpurun_out/r02_sweep.txt | grep -v json
