#!/bin/bash
# first GPU call of round 2: full GPU test suite, bench, launch lists of the two training steps
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests -m gpu -q --tb=short -x -k "fullsize" > gpurun_out/r02_pytest_fullsize.log 2>&1; echo "fullsize_rc=$?" >> gpurun_out/r02_pytest_fullsize.log
python -m pytest tests -m gpu -q --tb=short --deselect tests/test_gpu_fullsize.py > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/r02_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench_rc=$?" >> gpurun_out/r02_bench_n1.err
for v in pooled attention; do
  python tools/profile_step.py --variant $v --mode train --steps 1 --warmup 2 > gpurun_out/plain_$v.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_${v}_train.csv \
      python tools/profile_step.py --variant $v --mode train --steps 1 --warmup 2 > gpurun_out/ncu_$v.log 2>&1
done
tail -3 gpurun_out/r02_pytest_fullsize.log gpurun_out/r02_pytest_gpu.log; tail -c 600 gpurun_out/r02_bench_n1.err
