#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py > gpurun_out/r02_dp_check_n2.log 2>&1; echo "dp_rc=$?" >> gpurun_out/r02_dp_check_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench_rc=$?" >> gpurun_out/r02_bench_n2.err
tail -n 4 gpurun_out/r02_dp_check_n2.log; tail -n 3 gpurun_out/r02_bench_n2.err | cut -c1-300; cut -c1-400 gpurun_out/r02_bench_n2.json
