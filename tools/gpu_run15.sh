#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_pooled.py tests/test_gpu_edge_cases.py -q --tb=short > gpurun_out/r02_pytest_dec.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_dec.log
python tools/bench_decode.py > gpurun_out/r02_bench_decode.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke_rc=$?" >> gpurun_out/r02_smoke.log
tail -n 4 gpurun_out/r02_pytest_dec.log | cut -c1-300; grep -v Warn gpurun_out/r02_bench_decode.txt | tail -5; tail -n 3 gpurun_out/r02_smoke.log
