#!/bin/bash
# round-end style run: smoke, full GPU suite, bench (ours + reference arm); outputs under gpurun_out/
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke_rc=$?" >> gpurun_out/r02_smoke.log
python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/r02_pytest_gpu.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench_rc=$?" >> gpurun_out/r02_bench_n1.err
tail -n 3 gpurun_out/r02_smoke.log; tail -n 5 gpurun_out/r02_pytest_gpu.log | cut -c1-300; tail -n 1 gpurun_out/r02_bench_n1.err; cut -c1-220 gpurun_out/r02_bench_n1.json
