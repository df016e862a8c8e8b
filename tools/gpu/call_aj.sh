#!/bin/bash
# multi-GPU bench line (N = number of GPUs of the box): gpurun --gpus N -- 'bash tools/gpu/call_aj.sh N'
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2aj_bench_n$N.json 2> gpurun_out/r2aj_bench_n$N.err
echo rc=$?; tail -c 400 gpurun_out/r2aj_bench_n$N.err; head -c 600 gpurun_out/r2aj_bench_n$N.json
