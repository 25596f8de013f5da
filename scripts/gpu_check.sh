#!/bin/bash
# One GPU-box visit: smoke, GPU parity tests, bench (both arms), ncu launch list + one full capture.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_check.sh [quick]
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
lscpu | grep -E 'Model name|^CPU\(s\)|Thread|Socket' > gpurun_out/cpu.txt 2>&1
echo "== smoke" ; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -3 gpurun_out/smoke.log
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -15 gpurun_out/pytest_gpu.log
echo "== bench" ; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err ; echo "bench rc=$?" ; cat gpurun_out/bench.json ; tail -3 gpurun_out/bench.err
if [ "${1:-}" != "quick" ]; then
echo "== bench reference arm" ; timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ; cat gpurun_out/bench_ref.json
echo "== reference CPU arm, other configs" ; timeout 600 python -m oracle.cpu_ref_bench --seconds 0.5 > gpurun_out/cpu_ref.jsonl 2> gpurun_out/cpu_ref.err ; cat gpurun_out/cpu_ref.jsonl
echo "== ncu launch list" ; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1 ; echo "rc=$?"
echo "== ncu full capture (fused fkine+jacob0 kernel)" ; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_fkj_fast -s 3 -c 2 -f -o gpurun_out/prof_fkj python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1 ; echo "rc=$?"
fi
ls -la gpurun_out
