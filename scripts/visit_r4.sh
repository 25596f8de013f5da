#!/bin/bash
# Round-2 visit 4: persistent (grid-stride, double-buffered) form of the generated recursion kernel, with and without
# fp32 row pairs; tests under the new mode, then the sweep.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tests with B2K_RNE_SPEC_PERSIST=1"; B2K_RNE_SPEC_PERSIST=1 timeout 1200 python -m pytest tests -m gpu -q -k "rne or tree or dyn or spec or accel or inertia or coriolis" > gpurun_out/pytest_r4.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_r4.log
: > gpurun_out/rne_persist_sweep.jsonl
run() { # case persist pair tpw minb [rows]
  echo "{\"persist\": $2, \"pair\": $3, \"tpw\": $4, \"minb\": $5, \"rows\": ${6:-1000000}}" >> gpurun_out/rne_persist_sweep.jsonl
  B2K_RNE_SPEC_PERSIST=$2 B2K_RNE_SPEC_PAIR=$3 B2K_RNE_SPEC_TPW=$4 B2K_RNE_SPEC_MINB=$5 timeout 300 python scripts/kernel_bench.py --only "$1" --steps 30 --rows ${6:-1000000} 2>/dev/null | grep -v generic | cut -c1-330 >> gpurun_out/rne_persist_sweep.jsonl
}
run rne_puma_f64 0 0 1 5
for m in 3 4 5; do run rne_puma_f64 1 0 1 $m; done
run rne_puma_f32 0 0 2 8
for m in 4 6 8; do run rne_puma_f32 1 0 1 $m; done
for m in 3 4 5; do run rne_puma_f32 1 1 1 $m; done
run rne_puma_f64 0 0 1 5 4000000; run rne_puma_f64 1 0 1 4 4000000
run rne_puma_f32 0 0 2 8 4000000; run rne_puma_f32 1 1 1 4 4000000; run rne_puma_f32 1 0 1 8 4000000
cat gpurun_out/rne_persist_sweep.jsonl
