#!/bin/bash
# Sweep of the specialised RNE kernel's scheduling knobs (tiles per warp, resident blocks): one JSON line per case.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/rne_sweep.jsonl
for tpw in 1 2 4; do for minb in 3 4 5 6; do
  echo "{\"tpw\": $tpw, \"minb\": $minb}" >> gpurun_out/rne_sweep.jsonl
  B2K_RNE_SPEC_TPW=$tpw B2K_RNE_SPEC_MINB=$minb timeout 300 python scripts/kernel_bench.py --only rne_puma_f --steps 30 2>/dev/null | grep -v generic | cut -c1-420 >> gpurun_out/rne_sweep.jsonl
done; done
cat gpurun_out/rne_sweep.jsonl
