#!/usr/bin/env python3
"""bench.py -- the BASELINE.json headline metric: Panda 7-DOF fkine+jacob0 evaluations/s at
batch 1M (configs[1]: Panda ETS, fp64, seed 0, q ~ U(-pi, pi), 1M rows per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]   # the reference's CPU path

A "step" is one pass of the hot path over one 1M-row batch: ONE launch of the fused
fkine+jacob0 kernel (b2k_fkine_jacob0 through the C ABI).  Rows are sharded over ranks with no
data-path collective (weak scaling: every rank evaluates its own 1M rows per step).

Timing: W warm-up steps, then exactly K steps bracketed by barrier + synchronize, CUDA events on
the launching stream, MAX over ranks.  Inputs are resident in HBM; four distinct q batches
(4 x 56 MB) are rotated and every step writes 464 MB, so nothing is served from the 126 MB L2.

Extra objects in the JSON line (see DESIGN.md "Measurement"):
  roofline      algorithmic bytes (520 B/eval) / average kernel time vs MEASURED_PEAKS hbm_gbs
  cpu_baseline  the reference's own fknm (oracle/_ref, built from /root/reference) on the host cores,
                bounded sample, rank 0 at N=1 only
  e2e           same metric through the public API with pinned HOST buffers (H2D + kernel + D2H)
  clocks        SM clocks / throttle reasons sampled through NVML during the timed region
  gpu_launches  kernels this library launched inside the timed region
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ROWS_PER_GPU = 1_000_000
N_JOINTS = 7
BYTES_PER_EVAL = (7 + 16 + 42) * 8  # q + T + J0, fp64
METRIC = "Panda 7-DOF fkine+jacob0 evals/sec @ batch 1M"
WORKLOAD = "panda_ets_fkine_jacob0_f64_batch1M"


def make_q(seed, rows=ROWS_PER_GPU):
    return np.random.default_rng(seed).uniform(-np.pi, np.pi, (rows, N_JOINTS))


# ------------------------------------------------------------------ reference CPU arm
_W = {}


def _ref_worker_init(use_ref):
    from oracle import chains as ch

    d = ch.panda_ets()
    if use_ref:
        from oracle import ref_driver as ref

        _W["ets"] = ref.RefETS(d)
        _W["f"] = ref.fknm()
    else:
        from oracle import oracle as orc

        orc.set_threads(1)
        _W["chain"] = orc.Chain(d)
    _W["use_ref"] = use_ref


def _ref_worker_run(args):
    """One worker's slice of a step: the reference's own way to get N poses + N Jacobians --
    one batched ETS_fkine call plus a per-row ETS_jacob0 loop (the reference has no batched
    Jacobian: SURVEY 3.2).  Returns a checksum so the work cannot be optimised away."""
    seed, rows = args
    Q = make_q(seed, rows)
    if _W["use_ref"]:
        f, ets = _W["f"], _W["ets"].ets
        T = f.ETS_fkine(ets, Q, None, None, 1)
        s = float(T[-1, 0, 3])
        jac = f.ETS_jacob0
        for i in range(rows):
            J = jac(ets, Q[i], None)
        return s + float(J[0, 0])
    C = _W["chain"]
    return float(C.fkine(Q)[-1, 0, 3]) + float(C.jacob0(Q)[-1, 0, 0])


class RefArm:
    """The reference implementation of the path on the host cores (all of them)."""

    def __init__(self, cores=None):
        import multiprocessing as mp
        from oracle import ref_driver as ref

        self.use_ref = ref.available()
        self.kind = "reference" if self.use_ref else "port"
        self.cores = cores or len(os.sched_getaffinity(0))
        self.pool = mp.get_context("fork").Pool(self.cores, initializer=_ref_worker_init, initargs=(self.use_ref,))
        self.pool.map(_ref_worker_run, [(i, 64) for i in range(self.cores)])  # spin up + build chains

    def step(self, rows_per_core, seed0=0):
        t = time.perf_counter()
        self.pool.map(_ref_worker_run, [(seed0 + i, rows_per_core) for i in range(self.cores)])
        return time.perf_counter() - t

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_baseline(rows_per_core=100_000, reps=3):
    arm = RefArm()
    best = min(arm.step(rows_per_core, seed0=100 * r) for r in range(reps))
    total = rows_per_core * arm.cores
    # single-core figure: what the (single-threaded, GIL-holding) reference delivers out of the box
    one = RefArm(cores=1)
    t1 = min(one.step(rows_per_core // 2, seed0=7 + r) for r in range(2))
    one.close()
    arm.close()
    return {
        "value": total / best, "unit": "evals/s", "cores": arm.cores, "kind": arm.kind,
        "sample": f"{total} rows ({rows_per_core}/core x {arm.cores} processes), best of {reps}: "
                  "fknm.ETS_fkine batch call + per-row fknm.ETS_jacob0 loop (the reference has no batched Jacobian)",
        "single_core_value": (rows_per_core // 2) / t1,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    arm = RefArm()
    rows_per_core = 20_000
    for _ in range(args.warmup):
        arm.step(rows_per_core)
    t = 0.0
    for k in range(args.steps):
        t += arm.step(rows_per_core, seed0=1000 + k)
    arm.close()
    total = rows_per_core * arm.cores
    ms = 1e3 * t / args.steps
    value = total / (t / args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_rows_per_step": total,
                   "note": "reference CPU implementation (fknm built from /root/reference) on the host cores; "
                           "a step is a bounded sample of the 1M-row workload"},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": arm.cores, "kind": arm.kind,
                         "sample": f"{total} rows per step ({rows_per_core}/core x {arm.cores} processes)"},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------ clocks sampler (NVML)
class ClockSampler(threading.Thread):
    def __init__(self, torch_dev, period=0.002):
        super().__init__(daemon=True)
        self.period = period
        self.samples = []  # (t, sm_mhz, reasons_bitmask)
        self.windows = []
        self._stop = threading.Event()
        self.h = None
        self.max_mhz = None
        try:
            import pynvml
            import torch

            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(torch_dev).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(torch_dev)
            self.nv = pynvml
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def run(self):
        if self.h is None:
            return
        nv = self.nv
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), int(mhz), int(rs)))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop.set()

    def summary(self, t0, t1):
        names = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
                 0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}
        inwin = [s for s in self.samples if t0 <= s[0] <= t1]
        note = None
        if not inwin:  # timed region shorter than one sampling period: use the closest samples
            inwin = sorted(self.samples, key=lambda s: abs(s[0] - 0.5 * (t0 + t1)))[:3]
            note = "timed region shorter than the sampling period; nearest samples used"
        if not inwin:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        mask = 0
        for s in inwin:
            mask |= s[2]
        reasons = [n for b, n in names.items() if mask & b and n != "gpu_idle"]
        out = {"sm_mhz": float(np.median([s[1] for s in inwin])), "sm_max_mhz": self.max_mhz, "reasons": reasons,
               "samples": len(inwin)}
        if note:
            out["note"] = note
        return out


# ------------------------------------------------------------------ the B200 arm
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def load_traffic():
    """dram bytes per launch of the fused kernel from the committed ncu --set full capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(WORKLOAD)
        except Exception:
            return None
    return None


def run_b200(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = args.gpus
    if world != n_gpus and world > 1:
        n_gpus = world

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline()  # before CUDA is initialised in this process (the pool forks)

    import torch

    import b2kin as rtb

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    ets = rtb.models.Panda().ets()
    NBUF = 4
    lo, hi = rtb.dist.shard_bounds(ROWS_PER_GPU * world, world, rank)  # this rank's rows of the global batch
    assert hi - lo == ROWS_PER_GPU
    qs = [torch.from_numpy(make_q(1000 * b + rank)).to(dev) for b in range(NBUF)]
    T = torch.empty((ROWS_PER_GPU, 4, 4), dtype=torch.float64, device=dev)
    J = torch.empty((ROWS_PER_GPU, 6, N_JOINTS), dtype=torch.float64, device=dev)
    L = rtb._lib.lib()
    chain = ets._chain
    stream = torch.cuda.current_stream(dev)

    def step(i):
        q = qs[i % NBUF]
        rtb._lib.check(L.b2k_fkine_jacob0(chain, rtb._lib.F64, q.data_ptr(), ROWS_PER_GPU, N_JOINTS, None, None,
                                          T.data_ptr(), J.data_ptr(), stream.cuda_stream))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local)
    sampler.start()
    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = rtb.launch_count()
    t0 = time.perf_counter()
    e0.record(stream)
    for i in range(args.steps):
        step(i)
    e1.record(stream)
    barrier()
    t1 = time.perf_counter()
    launches = rtb.launch_count() - n0
    ms_total = e0.elapsed_time(e1)
    tt = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total = float(tt.item())
    ms_step = ms_total / args.steps
    value = ROWS_PER_GPU * world / (ms_step * 1e-3)
    clocks = sampler.summary(t0, t1)

    # ---- optional reassembly of the shards (reported separately; not part of the metric)
    gather_ms = None
    if dist is not None and args.gather:
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        rtb.dist.gather_rows(T, ROWS_PER_GPU * world)
        rtb.dist.gather_rows(J, ROWS_PER_GPU * world)
        g1.record(stream)
        barrier()
        gt = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
        dist.all_reduce(gt, op=dist.ReduceOp.MAX)
        gather_ms = float(gt.item())

    # ---- end to end through the public API with pinned host buffers (H2D + kernel + D2H per step)
    e2e_steps = max(2, min(args.steps, 5))
    qp = [rtb.pinned_empty((ROWS_PER_GPU, N_JOINTS)) for _ in range(2)]
    for b in range(2):
        qp[b][:] = make_q(77 + b + 10 * rank)
    Tp = rtb.pinned_empty((ROWS_PER_GPU, 4, 4))
    Jp = rtb.pinned_empty((ROWS_PER_GPU, 6, N_JOINTS))
    ets.fkine_jacob0_into(qp[0], Tp, Jp)  # warm-up (allocates the pipeline's device staging)
    ets.fkine_jacob0_into(qp[1], Tp, Jp)
    barrier()
    w0 = time.perf_counter()
    for i in range(e2e_steps):
        ets.fkine_jacob0_into(qp[i % 2], Tp, Jp)  # synchronous: results are in Tp / Jp on return
    torch.cuda.synchronize(dev)
    w1 = time.perf_counter()
    et = torch.tensor([(w1 - w0) / e2e_steps], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(et, op=dist.ReduceOp.MAX)
    e2e_value = ROWS_PER_GPU * world / float(et.item())
    checksum = float(Tp[-1, 0, 3]) + float(Jp[-1, 0, 0])
    sampler.stop()

    if rank == 0:
        peak, peak_src = load_peaks()
        achieved = BYTES_PER_EVAL * ROWS_PER_GPU / (ms_step * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rows_per_gpu": ROWS_PER_GPU, "global_batch": ROWS_PER_GPU * world,
                       "parallelism": f"rows sharded over {world} rank(s), no data-path collective",
                       "l2": "4 distinct 56 MB q batches rotated + 464 MB written per step (> 126 MB L2)",
                       "kernel": "k_fkj_fast<double,7,T,J0> (1 launch per step)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": load_traffic(), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": BYTES_PER_EVAL * ROWS_PER_GPU},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": ROWS_PER_GPU * N_JOINTS * 8 * world,
                    "d2h_bytes_per_step": ROWS_PER_GPU * (16 + 42) * 8 * world, "steps": e2e_steps,
                    "api": "ETS.fkine_jacob0_into(pinned q, T, J) -> b2k_fkine_jacob0_host", "checksum": checksum},
            "clocks": clocks, "gpu_launches": int(launches),
        }
        if gather_ms is not None:
            line["gather"] = {"ms": gather_ms, "bytes_received_per_rank": ROWS_PER_GPU * (world - 1) * 58 * 8,
                              "note": "NCCL all-gather of T and J shards, outside the metric"}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--gather", action="store_true", help="also time an NCCL all-gather of the result shards (N>1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
