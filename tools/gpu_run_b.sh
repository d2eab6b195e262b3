#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_pooled.py tests/test_gpu_fullsize.py tests/test_gpu_ops.py tests/test_gpu_edge_cases.py -m gpu -q --tb=short > gpurun_out/r02_pytest_b.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_b.log
tail -n 8 gpurun_out/r02_pytest_b.log | cut -c1-250
timeout 300 python tools/bench_gru.py 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_quick.json 2> gpurun_out/r02_bench_quick.err; echo "bench_rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_quick.json'))
print(d['value'], d['ms_per_step'], d['e2e'])
print({k:(round(v['ms'],3),round(v['frac'],3)) for k,v in d.get('rooflines',{}).items()})
PY
