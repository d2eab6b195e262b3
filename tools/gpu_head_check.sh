#!/bin/bash
# short sanity run of the final code: smoke, the TMA-store bit-identity test, the headline bench line (no extras)
mkdir -p gpurun_out
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_head_smoke.log 2>&1; echo "smoke_rc=$?" >> gpurun_out/r02_head_smoke.log
timeout 40 python -m pytest tests/test_gpu_gemm_tc.py -m gpu -q --tb=short -k "tma_store or a_stationary" > gpurun_out/r02_head_pytest_tma.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/r02_head_pytest_tma.log
timeout 60 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r02_head_bench_noextras.json 2> gpurun_out/r02_head_bench.err; echo "bench_rc=$?" >> gpurun_out/r02_head_bench.err
grep -v "Warn\|warn" gpurun_out/r02_head_smoke.log | tail -n 3; tail -n 3 gpurun_out/r02_head_pytest_tma.log | cut -c1-200; tail -n 1 gpurun_out/r02_head_bench.err; cut -c1-260 gpurun_out/r02_head_bench_noextras.json
