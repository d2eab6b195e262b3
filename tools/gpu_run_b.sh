#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_pooled.py tests/test_gpu_fullsize.py tests/test_gpu_bf16.py -m gpu -q --tb=short -x -k "fused or graph or bf16 or full_size or async or golden" > gpurun_out/r02_pytest_b.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_b.log
tail -n 6 gpurun_out/r02_pytest_b.log | cut -c1-250
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r02_bench_quick.json 2> gpurun_out/r02_bench_quick.err; echo "bench_rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_quick.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('cuda_graph'), d['loss'])
PY
timeout 400 python tools/timeline_step.py pooled > gpurun_out/r02_timeline_pooled.txt 2>gpurun_out/timeline.err; head -1 gpurun_out/r02_timeline_pooled.txt
