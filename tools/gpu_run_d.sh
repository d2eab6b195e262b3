#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pooled.py tests/test_gpu_edge_cases.py tests/test_gpu_attention.py tests/test_gpu_gemm_tc.py -m gpu -q --tb=short -x -k "infer or greedy or decode or amax or argmax" > gpurun_out/r02_pytest_dec.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_dec.log
tail -n 6 gpurun_out/r02_pytest_dec.log | cut -c1-300
for m in end step; do CAPHN_DECODE_SOFTMAX=$m timeout 300 python tools/bench_decode.py softmax=$m 2>&1 | head -3 | cut -c1-300; done
