#!/bin/bash
# IK occupancy experiment (round 2): the same IK kernels compiled for 4 / 5 / 6 / 8 resident blocks per SM
# (lib/exp/libb2kin_m*.so, built by `make -C csrc ik_exp`), timed on the config-4 protocols, plus the
# per-kernel durations of the first-search segments at several batch sizes.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for m in ${IK_VARIANTS:-1 5 6 8}; do
  L=robotics-toolbox-python_b200/lib/exp/libb2kin_m$m.so
  [ -f "$L" ] || continue
  for rows in ${IK_ROWS:-100000}; do
    B2K_LIB=$PWD/$L timeout 300 python scripts/kernel_bench.py --only "${IK_ONLY:-ik_lm_panda_f32}" --ik-rows $rows 2> gpurun_out/ik_occ_m$m.err | sed "s/^{/{\"minb\": $m, \"rows\": $rows, /" | tee -a gpurun_out/ik_occ.jsonl | cut -c1-260
  done
done
for m in ${IK_NCU_VARIANTS:-1 6}; do
  L=robotics-toolbox-python_b200/lib/exp/libb2kin_m$m.so
  [ -f "$L" ] || continue
  for rows in ${IK_NCU_ROWS:-18944 75776 100000}; do
    B2K_LIB=$PWD/$L timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ik -c 24 --csv --log-file gpurun_out/ik_occ_launches_m${m}_$rows.csv python scripts/kernel_bench.py --only ik_lm_panda_f32_chan0.1_jl0 --ik-rows $rows > /dev/null 2>&1
    python - <<P
import csv
rows=list(csv.reader(open("gpurun_out/ik_occ_launches_m${m}_$rows.csv")))
h=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]
ki=rows[h].index("Kernel Name"); vi=rows[h].index("Metric Value")
d=[(r[ki].split("(")[0].replace("void ",""), float(r[vi].replace(",",""))/1e3) for r in rows[h+2:] if len(r)>vi]
print("m${m} rows=$rows last 4 launches (us):", [(n[:22], round(t,1)) for n,t in d[-4:]])
P
  done
done
