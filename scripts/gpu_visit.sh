#!/bin/bash
# One GPU-box visit (round 2): GPU tests, bench (both arms), kernel table, ncu captures of the kernels that had
# none in round 1 (IK phases, RNE fp32, the coriolis fan-out).  Usage under gpurun: bash scripts/gpu_visit.sh [what...]
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
WHAT="${*:-tests bench ref kernels ncu}"
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
(lscpu | grep -E 'Model name|^CPU\(s\)|Thread|Socket|NUMA'; cat /sys/fs/cgroup/cpu.max 2>/dev/null; nproc) > gpurun_out/cpu.txt 2>&1
cap() { # name, kernel regex, --only filter, skip
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s ${4:-2} -c 1 -f -o gpurun_out/prof_$1 python scripts/kernel_bench.py --steps 2 --warmup 2 --only "$3" > gpurun_out/ncu_$1.log 2>&1; echo "ncu $1 rc=$?"
}
for w in $WHAT; do case $w in
tests) echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log;;
smoke) timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log;;
bench) echo "== bench"; timeout 900 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err;;
ref) echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err;;
kernels) echo "== kernel table"; timeout 900 python scripts/kernel_bench.py ${KB_ARGS:-} > gpurun_out/kernels.jsonl 2> gpurun_out/kernels.err; echo "rc=$?"; cut -c1-300 gpurun_out/kernels.jsonl; tail -3 gpurun_out/kernels.err;;
ncu) echo "== ncu captures"
  cap ikA k_ik_lm ik_lm_panda_f32_chan0.1_jl0 1
  cap ikB k_ik_restarts ik_lm_panda_f32_chan0.1_jl0 1
  cap rne32 k_rne rne_puma_f32 2
  cap coriolis k_rne_fan dyn_coriolis 1;;
ncu_rne) cap rne64s k_rne_spec rne_puma_f64 2; cap rne32s k_rne_spec rne_puma_f32 2;; ncu_rne_old) cap rne64 "${NCU_RNE_K:-k_rne}" rne_puma_f64 2; cap rne32 "${NCU_RNE_K:-k_rne}" rne_puma_f32 2;;
ncu_late) cap hesstile k_hessian_tile extra_hessian 1; cap frames k_fk_frames extra_fkine_all_panda_f64 1;;
launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "launch list rc=$?";;
esac; done
ls -la gpurun_out | head -40
