#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench_rc=$?" >> gpurun_out/r02_bench_n1.err
SEL="(grouped_gemms and 17) or batched_beam_search_golden or pooled_grouped_matches or attention_grouped_matches or ce_statistics or golden_greedy_token_exact"
python -m pytest tests/test_gpu_grouped_kernels.py tests/test_gpu_attention.py tests/test_gpu_grouped.py tests/test_gpu_ops.py -k "$SEL" -x -q > gpurun_out/r02_sanitizer_plain.log 2>&1 &&
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_grouped_kernels.py tests/test_gpu_attention.py tests/test_gpu_grouped.py tests/test_gpu_ops.py -k "$SEL" -x -q > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "memcheck_rc=$?" >> gpurun_out/r02_sanitizer_memcheck.log
tail -n 2 gpurun_out/r02_bench_n1.err; cut -c1-200 gpurun_out/r02_bench_n1.json; tail -n 3 gpurun_out/r02_sanitizer_plain.log; tail -n 12 gpurun_out/r02_sanitizer_memcheck.log | cut -c1-300
