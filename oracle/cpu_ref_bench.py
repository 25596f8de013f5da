#!/usr/bin/env python3
"""The reference's own CPU implementation of the OTHER BASELINE.json configs, timed on the host cores
(bench.py --impl reference covers configs[1]).  TEST / MEASUREMENT INFRASTRUCTURE (lives under oracle/ for that
reason; nothing in the product imports it): drives oracle/_ref
(the reference's fknm/frne, built from /root/reference) or, when that is absent, the oracle port.

One JSON line per case, all host threads (one process per core, like bench.py's RefArm) and one core:
  puma560_rne_f64      configs[2]  frne.frne per row (the loop DHRobot.rne runs: reference DHRobot.py:1442-1451)
  panda_ik_lm_f64      configs[3]  fknm.IK_LM_c per target (reference ETS.py ik_LM -> fknm.cpp:81-139)
  ur10_fkine_jacob0    configs[4]  fknm.ETS_fkine batch + per-row ETS_jacob0 (fp64: the reference has no fp32)

Usage: python -m oracle.cpu_ref_bench [--seconds 5]   (from the repository root)
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

_W = {}


def _init(case):
    from oracle import chains as ch
    from oracle import oracle as orc
    from oracle import ref_driver as ref

    use_ref = ref.available()
    _W["use_ref"] = use_ref
    orc.set_threads(1)
    if case == "puma560_rne_f64":
        links, grav = ch.puma560_links(), np.array([0.0, 0.0, 9.81])
        L = ch.pack_rne(links)
        _W["rne"] = ref.RefRNE(6, 0, L, grav) if use_ref else (lambda q, qd, qdd: orc.rne(6, 0, L, grav, q, qd, qdd))
    elif case == "panda_ik_lm_f64":
        d = ch.panda_ets()
        _W["ets"] = ref.RefETS(d) if use_ref else orc.Chain(d)
        _W["fk"] = orc.Chain(d)
    else:
        d = ch.dh_to_ets(ch.ur10_links())
        _W["ets"] = ref.RefETS(d) if use_ref else orc.Chain(d)


def _run(a):
    case, seed, rows = a
    rng = np.random.default_rng(seed)
    if case == "puma560_rne_f64":
        q, qd, qdd = (rng.uniform(-np.pi, np.pi, (rows, 6)) for _ in range(3))
        r = _W["rne"]
        tau = r.rne(q, qd, qdd) if _W["use_ref"] else r(q, qd, qdd)
        return float(tau[-1, 0])
    if case == "panda_ik_lm_f64":
        from oracle import chains as ch

        lim = np.asarray(ch.panda_ets()["qlim"]).reshape(-1, 2)
        lim = lim[np.asarray(ch.panda_ets()["isjoint"], bool)]
        qt = rng.uniform(lim[:, 0], lim[:, 1], (rows, 7))
        Tep = _W["fk"].fkine(qt)
        if _W["use_ref"]:
            out = _W["ets"].ik_lm(Tep, ilimit=30, slimit=100, tol=1e-6, joint_limits=False, k=0.1, method="chan")
        else:
            out = _W["ets"].ik_lm(Tep, None, 30, 100, 1e-6, False, np.ones(6), 0.1, "chan", seed=seed)
        return float(np.sum(out[1]))
    Q = rng.uniform(-np.pi, np.pi, (rows, 6))
    e = _W["ets"]
    if _W["use_ref"]:
        from oracle import ref_driver as ref

        f = ref.fknm()
        T = f.ETS_fkine(e.ets, Q, None, None, 1)
        for i in range(rows):
            J = f.ETS_jacob0(e.ets, Q[i], None)
        return float(T[-1, 0, 3]) + float(J[0, 0])
    return float(e.fkine(Q)[-1, 0, 3]) + float(e.jacob0(Q)[-1, 0, 0])


def time_case(case, cores, rows_per_core, reps=2):
    pool = mp.get_context("fork").Pool(cores, initializer=_init, initargs=(case,))
    pool.map(_run, [(case, i, 8) for i in range(cores)])
    best, chk = None, 0.0
    for r in range(reps):
        t = time.perf_counter()
        out = pool.map(_run, [(case, 100 * r + i, rows_per_core) for i in range(cores)])
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
        chk = float(np.sum(out))
    pool.close()
    pool.join()
    return rows_per_core * cores / best, chk


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=4.0, help="target CPU seconds per timed repetition")
    args = ap.parse_args()
    from oracle import ref_driver as ref

    kind = "reference" if ref.available() else "port"
    cores = len(os.sched_getaffinity(0))
    guess = {"puma560_rne_f64": 4e5, "panda_ik_lm_f64": 2.5e3, "ur10_fkine_jacob0": 8e5}  # rows/s/core, first guess
    units = {"puma560_rne_f64": "rows/s", "panda_ik_lm_f64": "solves/s", "ur10_fkine_jacob0": "evals/s"}
    for case in guess:
        one, _ = time_case(case, 1, max(64, int(guess[case] * 0.5)), reps=1)  # calibrate
        rows = max(64, int(one * args.seconds))
        v1, _ = time_case(case, 1, rows)
        vall, chk = time_case(case, cores, rows)
        line = {"case": case, "kind": kind, "unit": units[case], "cores": cores, "value_all_cores": vall,
                "value_one_core": v1, "sample": f"{rows} rows per core per repetition, best of 2", "checksum": chk}
        if case == "panda_ik_lm_f64":
            line["success_rate"] = chk / (rows * cores)
            line["settings"] = "ilimit 30, slimit 100, tol 1e-6, chan k=0.1, no joint-limit check, random q0"
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    sys.exit(main())
