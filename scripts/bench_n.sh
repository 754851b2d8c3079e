#!/bin/bash
# usage: scripts/bench_n.sh N TAG   -> gpurun_out/r02_bench_${N}gpu_${TAG}.json (the driver's launch line for N > 1)
N=$1; TAG=$2; mkdir -p gpurun_out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_${N}gpu_${TAG}.json 2> gpurun_out/r02_bench_${N}gpu_${TAG}.err
tail -c 200 gpurun_out/r02_bench_${N}gpu_${TAG}.err; grep "^{" gpurun_out/r02_bench_${N}gpu_${TAG}.json | cut -c1-180
