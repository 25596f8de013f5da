"""The reference's CPU implementation of the hot path, timed on the host cores -- TEST / MEASUREMENT
INFRASTRUCTURE ONLY (used by bench.py's cpu_baseline leg, `bench.py --impl reference` and
oracle/cpu_ref_bench.py; nothing under robotics-toolbox-python_b200/ imports it).

What runs is the reference's OWN native code (oracle/_ref/{fknm,frne}*.so, compiled from the sources under
/root/reference by `make -C oracle ref`), driven the way the reference's Python layer drives it; where
oracle/_ref is absent the plain-C restatement (oracle_kin.c) stands in and `kind` says "port".

Cases (BASELINE.json configs):
  panda_fkj   configs[1]  fknm.ETS_fkine on the batch + the per-row fknm.ETS_jacob0 loop (the reference has no
                          batched Jacobian: fknm.cpp:785-850 takes ONE q)
  puma_rne    configs[2]  frne.frne per row, the loop DHRobot.rne runs (DHRobot.py:1442-1451)
  panda_ik    configs[3]  fknm.IK_LM_c per target (ETS.ik_LM -> fknm.cpp:394-525), chan, lambda and the
                          joint-limit check as given
  ur10_fkj    configs[4]  as panda_fkj on the UR10 DH chain (fp64: the reference has no fp32 path)

The reference is single-threaded and holds the GIL for the whole call; "all cores" here means one worker
PROCESS per schedulable core, each with its own rows (the path is embarrassingly parallel), which is the most
the reference can be made to deliver on a host.  The single-core leg runs IN THE CALLING PROCESS, so the
compiled reference modules are mapped into the process the driver inspects.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

UNITS = {"panda_fkj": "evals/s", "puma_rne": "rows/s", "panda_ik": "solves/s", "ur10_fkj": "evals/s"}
# rows per second per core, first guess (used to size bounded samples)
RATE_GUESS = {"panda_fkj": 8e5, "puma_rne": 3e5, "panda_ik": 3e4, "ur10_fkj": 8e5}
DESCRIBE = {
    "panda_fkj": "fknm.ETS_fkine batch call + per-row fknm.ETS_jacob0 loop (the reference has no batched Jacobian)",
    "puma_rne": "frne.frne per row (the loop DHRobot.rne runs)",
    "panda_ik": "fknm.IK_LM_c per target (ilimit 30, slimit 100, tol 1e-6, chan, random restarts from libc rand())",
    "ur10_fkj": "fknm.ETS_fkine batch call + per-row fknm.ETS_jacob0 loop on the UR10 DH chain (fp64: no fp32 in the reference)",
}

_W = {}


def effective_cores():
    """(schedulable cores, cgroup CPU quota in cores or None)."""
    n = len(os.sched_getaffinity(0))
    quota = None
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            a, b = f.read().split()
        if a != "max":
            quota = float(a) / float(b)
    except Exception:
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            p = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                quota = q / p
        except Exception:
            pass
    return n, quota


def _setup(case, opts):
    """Build the per-process state of a case (reference capsules or oracle objects)."""
    from oracle import chains as ch
    from oracle import oracle as orc
    from oracle import ref_driver as ref

    use_ref = ref.available()
    orc.set_threads(1)
    st = {"use_ref": use_ref, "case": case, "opts": dict(opts or {})}
    if case == "puma_rne":
        L, grav = ch.pack_rne(ch.puma560_links()), np.array([0.0, 0.0, -9.81])
        st["L"], st["grav"] = L, grav
        st["rne"] = ref.RefRNE(6, 0, L, grav) if use_ref else None
    else:
        d = ch.panda_ets() if case in ("panda_fkj", "panda_ik") else ch.dh_to_ets(ch.ur10_links())
        st["desc"] = d
        st["ets"] = ref.RefETS(d) if use_ref else None
        st["chain"] = orc.Chain(d)
        st["n"] = int(np.sum(d["isjoint"]))
    return st


def _state(case, opts=None):
    key = (case, tuple(sorted((opts or {}).items())))
    if key not in _W:
        _W[key] = _setup(case, opts)
    return _W[key]


def evaluate(case, inputs, opts=None):
    """The reference's outputs for explicit inputs (numpy fp64), in the calling process.
    panda_fkj / ur10_fkj: inputs (Q,) -> (T, J0);  puma_rne: (q, qd, qdd) -> tau;
    panda_ik: (Tep, q0) -> (q, success, iterations, searches, E) with explicit q0 (deterministic)."""
    from oracle import oracle as orc

    st = _state(case, opts)
    o = st["opts"]
    if case in ("panda_fkj", "ur10_fkj"):
        (Q,) = inputs
        if st["use_ref"]:
            return st["ets"].fkine(Q), st["ets"].jacob0(Q)
        return st["chain"].fkine(Q), st["chain"].jacob0(Q)
    if case == "puma_rne":
        q, qd, qdd = inputs
        if st["use_ref"]:
            return st["rne"].rne(q, qd, qdd)
        return orc.rne(6, 0, st["L"], -st["grav"], q, qd, qdd)
    if case == "panda_ik":
        Tep, q0 = inputs
        kw = dict(ilimit=o.get("ilimit", 30), slimit=o.get("slimit", 100), tol=o.get("tol", 1e-6),
                  joint_limits=bool(o.get("jl", False)), k=float(o.get("k", 0.1)), method="chan")
        if st["use_ref"]:
            return st["ets"].ik_lm(Tep, q0=q0, **kw)
        return st["chain"].ik_lm(Tep, q0, kw["ilimit"], kw["slimit"], kw["tol"], kw["joint_limits"], np.ones(6),
                                 kw["k"], "chan", seed=1)
    raise ValueError(case)


def _gen_inputs(case, rng, rows, st):
    if case in ("panda_fkj", "ur10_fkj"):
        return (rng.uniform(-np.pi, np.pi, (rows, st["n"])),)
    if case == "puma_rne":
        return (rng.uniform(-np.pi, np.pi, (rows, 6)), rng.normal(size=(rows, 6)), rng.normal(size=(rows, 6)))
    qt = rng.uniform(-np.pi, np.pi, (rows, st["n"]))  # reachable targets: Tep = FK(q*), SURVEY 8d config 4
    return (st["chain"].fkine(qt), None)


def _work(args):
    """One worker's slice of a step; returns a checksum so the work cannot be optimised away."""
    case, opts, seed, rows = args
    st = _state(case, opts)
    rng = np.random.default_rng(seed)
    inp = _gen_inputs(case, rng, rows, st)
    if case in ("panda_fkj", "ur10_fkj"):
        (Q,) = inp
        if st["use_ref"]:
            from oracle import ref_driver as ref

            f, ets = ref.fknm(), st["ets"].ets
            T = f.ETS_fkine(ets, Q, None, None, 1)
            jac = f.ETS_jacob0
            J = None
            for i in range(rows):
                J = jac(ets, Q[i], None)
            return float(T[-1, 0, 3]) + float(J[0, 0])
        C = st["chain"]
        return float(C.fkine(Q)[-1, 0, 3]) + float(C.jacob0(Q)[-1, 0, 0])
    if case == "puma_rne":
        return float(evaluate(case, inp, opts)[-1, 0])
    out = evaluate(case, inp, opts)
    return float(np.sum(out[1]))


def _warm(args):
    case, opts = args
    _state(case, opts)
    return _work((case, opts, 0, 8))


class CpuArm:
    """`cores` worker processes, each holding its own reference capsules for `case`."""

    def __init__(self, case, cores=None, opts=None):
        from oracle import ref_driver as ref

        self.case, self.opts = case, dict(opts or {})
        self.kind = "reference" if ref.available() else "port"
        self.cores = cores or len(os.sched_getaffinity(0))
        self.pool = mp.get_context("fork").Pool(self.cores)
        self.pool.map(_warm, [(case, self.opts)] * self.cores)

    def step(self, rows_per_core, seed0=0):
        t = time.perf_counter()
        out = self.pool.map(_work, [(self.case, self.opts, seed0 + i, rows_per_core) for i in range(self.cores)])
        self.checksum = float(np.sum(out))
        return time.perf_counter() - t

    def close(self):
        self.pool.close()
        self.pool.join()


def time_inline(case, rows, opts=None, reps=2, seed0=7):
    """Single-core leg IN THIS PROCESS (loads the reference modules here): rows / best time."""
    _warm((case, opts or {}))
    best = None
    for r in range(reps):
        t = time.perf_counter()
        _work((case, opts or {}, seed0 + r, rows))
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    return rows / best


def loaded_reference_modules():
    """Paths of the compiled reference modules mapped into this process (evidence for `kind: reference`)."""
    out = []
    try:
        with open("/proc/self/maps") as f:
            for line in f:
                p = line.rsplit(" ", 1)[-1].strip()
                if "/oracle/_ref/" in p and p not in out:
                    out.append(p)
    except Exception:
        pass
    return [os.path.relpath(p, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))) for p in out]


def baseline(case, opts=None, seconds=1.0, reps=3):
    """The `cpu_baseline` object of bench.py for one case: all-core and single-core throughput of the
    reference on a bounded sample (about `seconds` of work per core per repetition)."""
    one_rows = max(64, int(RATE_GUESS[case] * 0.2))
    v1 = time_inline(case, one_rows, opts, reps=1)  # calibrate
    rows = max(64, int(v1 * seconds))
    v1 = time_inline(case, max(64, rows // 2), opts, reps=2)
    arm = CpuArm(case, opts=opts)
    best = min(arm.step(rows, seed0=100 * r) for r in range(reps))
    arm.close()
    total = rows * arm.cores
    n_sched, quota = effective_cores()
    return {
        "value": total / best, "unit": UNITS[case], "cores": arm.cores, "kind": arm.kind,
        "sample": f"{total} rows ({rows}/process x {arm.cores} processes), best of {reps}: {DESCRIBE[case]}",
        "single_core_value": v1,
        "effective_cores": round((total / best) / v1, 2),
        "cgroup_cpu_quota_cores": quota,
        "modules_loaded": loaded_reference_modules(),
    }
