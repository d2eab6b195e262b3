#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -x -k "fused_decode_step" > gpurun_out/r02_pytest_dec.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_dec.log
timeout 900 python -m pytest tests/test_gpu_pooled.py tests/test_gpu_edge_cases.py -m gpu -q --tb=short -x -k "infer or greedy or decode" >> gpurun_out/r02_pytest_dec.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_dec.log
tail -n 12 gpurun_out/r02_pytest_dec.log | cut -c1-300
timeout 300 python tools/bench_decode.py r02b > gpurun_out/r02_bench_decode_b.txt 2>&1; cat gpurun_out/r02_bench_decode_b.txt | cut -c1-300
CAPHN_NO_GRAPH=1 CAPHN_DECODE_ITERS=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_decode_v2.csv \
   python tools/bench_decode.py ncu > gpurun_out/ncu_dec.log 2>&1; echo "ncu rc=$?"
