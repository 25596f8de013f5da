#!/bin/bash
# N-GPU visit (N = $1, default 8): bench.py under torchrun exactly as the driver launches it.
set -u
cd "$(dirname "$0")/.."
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo$N.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "rc=$?"; tail -1 gpurun_out/bench_${N}gpu.json | cut -c1-4000; tail -3 gpurun_out/bench_${N}gpu.err
