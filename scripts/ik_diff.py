#!/usr/bin/env python3
"""Which rows of the IK fixtures (tests/golden/panda_ik.npz, produced by the compiled reference fknm.IK_LM_c with
explicit q0, slimit = 1, fp64) does the kernel NOT reproduce counter for counter, and why?  Prints one JSON line per
protocol: the differing rows with both sides' (success, iterations, residual), and the oracle's per-iteration residual
trace for those rows (the CPU restatement reproduces the reference on ALL rows, tests/test_oracle_cpu.py)."""
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import b2kin as rtb  # noqa: E402
from oracle import oracle as orc  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "panda_ik.npz"))
e = rtb.models.Panda().ets()
C = orc.Chain(e.describe())
Tep, q0 = torch.from_numpy(z["Tep"]).cuda(), torch.from_numpy(z["q0"]).cuda()
for tag, method, k, jl in (("chan1", "chan", 1.0, False), ("chan01", "chan", 0.1, False), ("sugi", "sugihara", 1e-4, False), ("jl", "chan", 1.0, True)):
    q, s, it, sr, E = (x.cpu().numpy() for x in e.ik_LM(Tep, q0=q0, slimit=1, joint_limits=jl, k=k, method=method))
    rs, rit = z[tag + "_success"], z[tag + "_it"]
    diff = np.nonzero((s != rs) | (it != rit))[0]
    rows = []
    for i in diff:
        # oracle trace: residual after each iteration budget 1..30 (ilimit = j stops after j iterations)
        tr = []
        for lim in range(1, 31):
            _, so, ito, _, Eo = C.ik_lm(z["Tep"][i:i + 1], q0=z["q0"][i:i + 1], ilimit=lim, slimit=1, joint_limits=jl, k=k, method=method)
            tr.append(float(Eo[0]))
            if so[0]:
                break
        rows.append({"row": int(i), "ref": [int(rs[i]), int(rit[i]), float(z[tag + "_E"][i]) if tag + "_E" in z else None],
                     "gpu": [int(s[i]), int(it[i]), float(E[i])], "oracle_E_trace_tail": tr[-6:]})
    print(json.dumps({"protocol": tag, "targets": int(len(rs)), "identical": float(1 - len(diff) / len(rs)), "differing_rows": rows}), flush=True)
