#!/bin/bash
# Round-2 visit 3: packed-fp32 (f32x2) recursion kernels -- tests, then a sweep of pairs x tiles per warp x resident
# blocks; IK first-segment length sweep and the fp64 3-block build.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests -m gpu -q -k "rne or tree or dyn or spec or fdyn or accel or inertia or coriolis" > gpurun_out/pytest_r3.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_r3.log
: > gpurun_out/rne_pair_sweep.jsonl
run() { # pair tpw minb
  echo "{\"pair\": $1, \"tpw\": $2, \"minb\": $3}" >> gpurun_out/rne_pair_sweep.jsonl
  B2K_RNE_SPEC_PAIR=$1 B2K_RNE_SPEC_TPW=$2 B2K_RNE_SPEC_MINB=$3 timeout 300 python scripts/kernel_bench.py --only "${4:-rne_puma_f32}" --steps 30 2>/dev/null | grep -v generic | cut -c1-330 >> gpurun_out/rne_pair_sweep.jsonl
}
run 0 2 8
for tpw in 1 2; do for minb in 3 4 5 6; do run 1 $tpw $minb; done; done
run 0 2 8 tree_rne_puma_f32; run 1 1 4 tree_rne_puma_f32; run 1 2 4 tree_rne_puma_f32
cat gpurun_out/rne_pair_sweep.jsonl
echo "== dyn fan-outs fp32/fp64 with defaults"; timeout 600 python scripts/kernel_bench.py --only dyn_ 2>/dev/null | grep -v generic | cut -c1-300 | tee gpurun_out/dyn_r3.jsonl
echo "== IK first-segment sweep"
: > gpurun_out/ik_seg_sweep.jsonl
for seg in 6 8 10 12 14; do
  B2K_IK_SEG1=$seg timeout 300 python scripts/kernel_bench.py --only ik_lm_panda 2>/dev/null | sed "s/^{/{\"seg1\": $seg, /" | cut -c1-300 >> gpurun_out/ik_seg_sweep.jsonl
done
B2K_LIB=$PWD/robotics-toolbox-python_b200/lib/exp/libb2kin_d3.so timeout 300 python scripts/kernel_bench.py --only ik_lm_panda_f64 2>/dev/null | sed 's/^{/{"f64_minb": 3, /' | cut -c1-300 >> gpurun_out/ik_seg_sweep.jsonl
cat gpurun_out/ik_seg_sweep.jsonl
