#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/dp_check.py > gpurun_out/r02_dp_check_n$N.log 2>&1; echo "dp_rc=$?" >> gpurun_out/r02_dp_check_n$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 20 --warmup 5 --no-extras > gpurun_out/r02_bench_n${N}_noextras.json 2> gpurun_out/r02_bench_n${N}.err; echo "bench_rc=$?" >> gpurun_out/r02_bench_n${N}.err
python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r02_bench_n1_noextras_samebox.json 2>/dev/null
grep -v "Warn\|warn\|\*\*\*\|OMP" gpurun_out/r02_dp_check_n$N.log | tail -n 3; tail -n 2 gpurun_out/r02_bench_n${N}.err | cut -c1-200; cut -c1-330 gpurun_out/r02_bench_n${N}_noextras.json; cut -c1-330 gpurun_out/r02_bench_n1_noextras_samebox.json
