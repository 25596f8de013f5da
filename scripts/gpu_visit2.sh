#!/bin/bash
# Two-GPU visit: the multi-device test, bench.py under torchrun (N = 2: UR10 sharded config, NCCL gather, fused store-to-root), N = 1 e2e for comparison.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
echo "== pytest multi-device"; timeout 600 python -m pytest tests -m gpu -q -k "non_current_device or hardening" > gpurun_out/pytest_2gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_2gpu.log
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "rc=$?"; cat gpurun_out/bench_2gpu.json | cut -c1-3000; tail -5 gpurun_out/bench_2gpu.err
echo "== bench reference arm N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref_2gpu.json 2> gpurun_out/bench_ref_2gpu.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_ref_2gpu.json
