#!/bin/bash
# Short GPU visit for kernel iteration: selected tests + kernel table + headline bench (+ optional ncu).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest (subset: ${PYTEST_K:-all})"
timeout 1500 python -m pytest tests -m gpu -x -q ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -12 gpurun_out/pytest_gpu.log
echo "== kernel table"
timeout 900 python scripts/kernel_bench.py ${KB_ARGS:-} > gpurun_out/kernels.jsonl 2> gpurun_out/kernels.err ; echo "rc=$?"; cat gpurun_out/kernels.jsonl | cut -c1-330 ; tail -3 gpurun_out/kernels.err
echo "== bench"
timeout 600 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err ; echo "bench rc=$?" ; cat gpurun_out/bench.json ; tail -3 gpurun_out/bench.err
if [ -n "${NCU_K:-}" ]; then
  echo "== ncu full capture ($NCU_K)"
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$NCU_K -s ${NCU_S:-3} -c ${NCU_C:-1} -f -o gpurun_out/prof_${NCU_NAME:-k} python scripts/kernel_bench.py --steps 3 --warmup 3 --only "${NCU_ONLY:-fkine_jacob0_panda_f64}" > gpurun_out/ncu_full.log 2>&1 ; echo "rc=$?"
fi
if [ -n "${NCU2_K:-}" ]; then
  echo "== ncu full capture 2 ($NCU2_K)"
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$NCU2_K -s ${NCU2_S:-3} -c 1 -f -o gpurun_out/prof_${NCU2_NAME:-k2} python scripts/kernel_bench.py --steps 3 --warmup 3 --only "${NCU2_ONLY:-rne_puma_f64}" > gpurun_out/ncu_full2.log 2>&1 ; echo "rc=$?"
fi
