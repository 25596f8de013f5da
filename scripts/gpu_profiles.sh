#!/bin/bash
# ncu --set full captures of the kernels that are not the headline one (one GPU, short commands).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cap() { # name, kernel regex, --only filter, skip
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s ${4:-2} -c 1 -f -o gpurun_out/prof_$1 python scripts/kernel_bench.py --steps 2 --warmup 2 --only "$3" > gpurun_out/ncu_$1.log 2>&1; echo "$1 rc=$?"
}
cap fk32 k_fkj_fast fkine_panda_f32 2
cap rne64 k_rne rne_puma_f64 2
cap ikA k_ik_lm ik_lm_panda_f32_chan0.1 1
cap ikB k_ik_restarts ik_lm_panda_f32_chan0.1 1
cap inertia k_rne_fan dyn_inertia 2
