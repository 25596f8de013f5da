#!/bin/bash
cd "$(dirname "$0")/.."
: > gpurun_out/rne_sweep32.jsonl
for tpw in 1 2; do for minb in 6 7 8; do
  echo "{\"tpw\": $tpw, \"minb\": $minb}" >> gpurun_out/rne_sweep32.jsonl
  B2K_RNE_SPEC_TPW=$tpw B2K_RNE_SPEC_MINB=$minb timeout 300 python scripts/kernel_bench.py --only rne_puma_f32 --steps 30 2>/dev/null | grep -v generic | cut -c1-420 >> gpurun_out/rne_sweep32.jsonl
done; done
cat gpurun_out/rne_sweep32.jsonl
