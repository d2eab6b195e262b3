#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_pooled.py tests/test_gpu_attention.py tests/test_gpu_gemm_tc.py tests/test_gpu_bf16.py tests/test_gpu_grouped_kernels.py -m gpu -q --tb=short --durations=8 > gpurun_out/r02_pytest_b.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_b.log
tail -n 25 gpurun_out/r02_pytest_b.log | cut -c1-250
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_quick.json 2> gpurun_out/r02_bench_quick.err; echo "bench_rc=$?"
cut -c1-400 gpurun_out/r02_bench_quick.json; tail -3 gpurun_out/r02_bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_quick.json'))
print({k:(round(v['ms'],3),round(v['frac'],3)) for k,v in d.get('rooflines',{}).items()})
print({k:v for k,v in d.get('extras',{}).items() if 'captions' in k and not isinstance(v,dict)})
PY
