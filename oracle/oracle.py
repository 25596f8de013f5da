"""ctypes front-end of the CPU oracle (oracle_kin.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's CPU legs import this
module.  The product package never does (tests/test_boundary_cpu.py greps for it).

A kinematic chain is passed as a neutral *description* dict so the oracle does
not depend on the product's classes::

    {"isjoint": int32[m], "axis": int32[m], "flip": int32[m], "jindex": int32[m],
     "T": float64[m,4,4] (row-major constant of each ET; identity for joints),
     "qlim": float64[m,2], "n": number_of_joints}

which is exactly the per-ET payload the reference hands ``fknm.ET_init``
(reference ET.py:100-125).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    """Compile oracle_kin.c -> liboracle_kin.so (gcc, a second or two)."""
    so = os.path.join(_HERE, "liboracle_kin.so")
    src = os.path.join(_HERE, "oracle_kin.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(
            ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-shared", "-fPIC", src, "-o", so, "-lm"]
        )
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_rand_u01.restype = C.c_double
        _LIB.orc_rand_u01.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]
        _LIB.orc_num_threads.restype = C.c_int
    return _LIB


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


class Chain:
    """Holds a chain description in the C layout the oracle functions take."""

    def __init__(self, desc):
        self.isjoint = _i32(desc["isjoint"])
        self.axis = _i32(desc["axis"])
        self.flip = _i32(desc["flip"])
        self.jindex = _i32(desc["jindex"])
        self.T = _f64(desc["T"]).reshape(-1, 16)
        self.m = int(self.isjoint.shape[0])
        self.n = int(self.isjoint.sum())
        ql = _f64(desc["qlim"]).reshape(-1, 2)
        sel = self.isjoint.astype(bool)
        # joint limits in chain order, as ETS_init caches them (fknm.cpp:1096-1108)
        self.qlim_l = np.ascontiguousarray(ql[sel, 0])
        self.qlim_h = np.ascontiguousarray(ql[sel, 1])

    def _head(self):
        return (
            C.c_int(self.m),
            _p(self.isjoint, C.c_int),
            _p(self.axis, C.c_int),
            _p(self.flip, C.c_int),
            _p(self.jindex, C.c_int),
            _p(self.T, C.c_double),
        )

    def _head_n(self):
        h = self._head()
        return (h[0], C.c_int(self.n)) + h[1:]

    # ------------------------------------------------------------------ FK
    def fkine(self, q, base=None, tool=None):
        q = np.atleast_2d(_f64(q))
        N, ld = q.shape
        out = np.empty((N, 4, 4))
        base, tool = _f64(base), _f64(tool)
        lib().orc_fkine(*self._head(), _p(q, C.c_double), C.c_long(N), C.c_long(ld),
                        _p(base, C.c_double), _p(tool, C.c_double), _p(out, C.c_double))
        return out

    def _jac(self, fn, q, tool):
        q = np.atleast_2d(_f64(q))
        N, ld = q.shape
        out = np.empty((N, 6, self.n))
        tool = _f64(tool)
        fn(*self._head_n(), _p(q, C.c_double), C.c_long(N), C.c_long(ld), _p(tool, C.c_double),
           _p(out, C.c_double))
        return out

    def jacob0(self, q, tool=None):
        return self._jac(lib().orc_jacob0, q, tool)

    def jacobe(self, q, tool=None):
        return self._jac(lib().orc_jacobe, q, tool)

    def fkine_jacob0(self, q, base=None, tool=None):
        return self.fkine(q, base, tool), self.jacob0(q, tool)

    # ------------------------------------------------------------------ IK
    def ik_lm(self, Tep, q0=None, ilimit=30, slimit=100, tol=1e-6, joint_limits=True, mask=None,
              k=1.0, method="chan", seed=0, semantics=0, rng_per_row=True):
        Tep = _f64(Tep).reshape(-1, 4, 4)
        N = Tep.shape[0]
        n = self.n
        if q0 is not None:
            q0 = np.ascontiguousarray(np.broadcast_to(_f64(q0).reshape(-1, n), (N, n)))
        we = None if mask is None else _f64(mask)
        # chan / wampler / sugihara (LM, k = lambda); nr (k = pinv_damping); gn
        meth = {"c": 0, "w": 1, "s": 2, "n": 3, "g": 4}[method[0].lower()]
        q = np.empty((N, n))
        succ = np.empty(N, dtype=np.int32)
        its = np.empty(N, dtype=np.int32)
        srch = np.empty(N, dtype=np.int32)
        E = np.empty(N)
        lib().orc_ik_lm(*self._head_n(), _p(self.qlim_l, C.c_double), _p(self.qlim_h, C.c_double),
                        _p(Tep, C.c_double), C.c_long(N), _p(q0, C.c_double), C.c_int(ilimit),
                        C.c_int(slimit), C.c_double(tol), C.c_int(int(bool(joint_limits))),
                        _p(we, C.c_double), C.c_double(k), C.c_int(meth), C.c_uint64(seed),
                        C.c_int(semantics), C.c_int(int(bool(rng_per_row))), _p(q, C.c_double),
                        _p(succ, C.c_int), _p(its, C.c_int), _p(srch, C.c_int), _p(E, C.c_double))
        return q, succ, its, srch, E


def angle_axis(Te, Tep):
    Te, Tep = _f64(Te), _f64(Tep)
    e = np.empty(6)
    lib().orc_angle_axis(_p(Te, C.c_double), _p(Tep, C.c_double), _p(e, C.c_double))
    return e


def rand_u01(seed, row, search, joint):
    return lib().orc_rand_u01(seed, row, search, joint)


def rne(n, mdh, L, grav, q, qd, qdd, fext=None):
    """L: 24 doubles per link (reference DHRobot.py:1340-1358); grav: the vector handed to
    frne, i.e. MINUS the robot's gravity (reference DHRobot.py:1449)."""
    L, grav = _f64(L), _f64(grav)
    q = np.atleast_2d(_f64(q)); qd = np.atleast_2d(_f64(qd)); qdd = np.atleast_2d(_f64(qdd))
    N = q.shape[0]
    fext = _f64(fext)
    tau = np.empty((N, n))
    lib().orc_rne(C.c_int(n), C.c_int(mdh), _p(L, C.c_double), _p(grav, C.c_double),
                  _p(q, C.c_double), _p(qd, C.c_double), _p(qdd, C.c_double), C.c_long(N),
                  _p(fext, C.c_double), _p(tau, C.c_double))
    return tau


def set_threads(t: int):
    lib().orc_set_threads(C.c_int(t))


def num_threads() -> int:
    return lib().orc_num_threads()


# ------------------------------------------------------------------ dynamics fan-outs (reference Dynamics.py), on top of rne()
# Restated loop for loop from the reference's DynamicsMixin; `rne_fn(q, qd, qdd, grav)` is either this
# module's rne (C restatement) or the compiled reference's frne via ref_driver.RefRNE.
def dyn_inertia(rne_fn, n, q):
    """Dynamics.py:752-758"""
    q = np.atleast_2d(q)
    out = np.zeros((q.shape[0], n, n))
    for k, qk in enumerate(q):
        out[k] = rne_fn(np.tile(qk, (n, 1)), np.zeros((n, n)), np.eye(n), np.zeros(3))
    return out


def dyn_gravload(rne_fn, n, q, gravity):
    """Dynamics.py:912-915"""
    q = np.atleast_2d(q)
    return rne_fn(q, np.zeros_like(q), np.zeros_like(q), gravity)


def dyn_itorque(rne_fn, n, q, qdd):
    """Dynamics.py:1456-1459"""
    q = np.atleast_2d(q)
    return rne_fn(q, np.zeros_like(q), np.atleast_2d(qdd), np.zeros(3))


def dyn_coriolis(rne_nofriction_fn, n, q, qd):
    """Dynamics.py:818-857 (the caller passes an rne of the friction-free robot, line 818)"""
    q, qd = np.atleast_2d(q), np.atleast_2d(qd)
    N = q.shape[0]
    C = np.zeros((N, n, n))
    Csq = np.zeros((N, n, n))
    z = np.zeros(n)
    g0 = np.zeros(3)
    for k, qk in enumerate(q):
        for i in range(n):
            QD = np.zeros(n); QD[i] = 1
            Csq[k, :, i] = Csq[k, :, i] + rne_nofriction_fn(qk, QD, z, g0)[0]
    for k, (qk, qdk) in enumerate(zip(q, qd)):
        for i in range(n):
            for j in range(i + 1, n):
                QD = np.zeros(n); QD[i] = 1; QD[j] = 1
                tau = rne_nofriction_fn(qk, QD, z, g0)[0]
                C[k, :, j] = C[k, :, j] + (tau - Csq[k, :, j] - Csq[k, :, i]) * qdk[i] / 2
                C[k, :, i] = C[k, :, i] + (tau - Csq[k, :, j] - Csq[k, :, i]) * qdk[j] / 2
        C[k] = C[k] + Csq[k] @ np.diag(qdk)
    return C


def dyn_accel(rne_fn, n, q, qd, torque, gravity):
    """Dynamics.py:490-503"""
    q, qd, torque = np.atleast_2d(q), np.atleast_2d(qd), np.atleast_2d(torque)
    out = np.zeros_like(q)
    for k, (qk, qdk, tk) in enumerate(zip(q, qd, torque)):
        M = rne_fn(np.tile(qk, (n, 1)), np.zeros((n, n)), np.eye(n), np.zeros(3))
        tau = rne_fn(qk, qdk, np.zeros(n), gravity)[0]
        out[k] = np.linalg.solve(M, tk - tau)
    return out


def nofriction_L(L):
    """robot.nofriction(coulomb=True, viscous=True): B = 0, Tc = 0 in the packed link table (Link.py:1548-1590)."""
    L = np.array(L, dtype=np.float64).reshape(-1, 24).copy()
    L[:, 21] = 0.0
    L[:, 22:24] = 0.0
    return L.ravel()


def hessian(J):
    """_ETS_hessian, methods.cpp:16-32: H (n,6,n) from J (6,n)."""
    J = np.asarray(J, dtype=np.float64)
    n = J.shape[1]
    H = np.zeros((n, 6, n))
    for j in range(n):
        for i in range(j, n):
            H[j, :3, i] = np.cross(J[3:, j], J[:3, i])
            H[j, 3:, i] = np.cross(J[3:, j], J[3:, i])
            if i != j:
                H[i, :3, j] = H[j, :3, i]
                H[i, 3:, j] = 0.0
    return H


def yoshikawa(J, axes=(True,) * 6):
    """ETS.py:1780-1787"""
    Ja = np.asarray(J, dtype=np.float64)[np.asarray(axes, dtype=bool), :]
    if Ja.shape[0] == Ja.shape[1]:
        return abs(np.linalg.det(Ja))
    return np.sqrt(abs(np.linalg.det(Ja @ Ja.T)))


def jacobm(J, H, axes=None):
    """ETS.jacobm ETS.py:1672-1685 / Robot.jacobm Robot.py:1215-1232: J (6,n), H (n,6,n)."""
    J = _f64(J); H = _f64(H)
    n = J.shape[1]
    ax = np.ones(6, bool) if axes is None else np.asarray(axes, bool)
    m = yoshikawa(J, ax)
    Ja, Ha = J[ax, :], H[:, ax, :]
    b = np.linalg.inv(Ja @ Ja.T)
    Jm = np.zeros((n, 1))
    for i in range(n):
        c = Ja @ Ha[i].T
        Jm[i, 0] = m * c.flatten("F") @ b.flatten("F")
    return Jm


def jacob_dot(H, qd):
    """Robot.jacob0_dot Robot.py:1099: np.tensordot(H, qd, (0, 0))."""
    return np.tensordot(_f64(H), _f64(qd), (0, 0))


def jtraj(q0, qf, t, qd0=None, qd1=None):
    """tools/trajectory.py:730-775 restated line by line: returns (tv, q, qd, qdd)."""
    if isinstance(t, (int, np.integer)):
        tscal = 1.0
        ts = np.linspace(0, 1, t)
        tv = ts * t
    else:
        t = _f64(t)
        tv = t.flatten()
        tscal = max(t)
        ts = t.flatten() / tscal
    q0, qf = _f64(q0).ravel(), _f64(qf).ravel()
    qd0 = np.zeros(q0.shape) if qd0 is None else _f64(qd0).ravel()
    qd1 = np.zeros(q0.shape) if qd1 is None else _f64(qd1).ravel()
    A = 6 * (qf - q0) - 3 * (qd1 + qd0) * tscal
    Bc = -15 * (qf - q0) + (8 * qd0 + 7 * qd1) * tscal
    Cc = 10 * (qf - q0) - (6 * qd0 + 4 * qd1) * tscal
    E = qd0 * tscal
    F = q0
    tt = np.array([ts**5, ts**4, ts**3, ts**2, ts, np.ones(ts.shape)]).T
    z = np.zeros(A.shape)
    qt = tt @ np.array([A, Bc, Cc, z, E, F])
    qdt = tt @ np.array([z, 5 * A, 4 * Bc, 3 * Cc, z, E]) / tscal
    qddt = tt @ np.array([z, z, 20 * A, 12 * Bc, 6 * Cc, z]) / tscal**2
    return tv, qt, qdt, qddt


def quintic(q0, qf, t, qd0=0, qdf=0):
    """tools/trajectory.py:329-345 + quintic_func 385-416, restated: returns (t, s, sd, sdd)."""
    t = np.arange(0, t) if isinstance(t, (int, np.integer)) else _f64(t).ravel()
    T = max(t)
    X = [[0.0, 0.0, 0.0, 0.0, 0.0, 1.0], [T**5, T**4, T**3, T**2, T, 1.0], [0.0, 0.0, 0.0, 0.0, 1.0, 0.0],
         [5.0 * T**4, 4.0 * T**3, 3.0 * T**2, 2.0 * T, 1.0, 0.0], [0.0, 0.0, 0.0, 2.0, 0.0, 0.0],
         [20.0 * T**3, 12.0 * T**2, 6.0 * T, 2.0, 0.0, 0.0]]
    coeffs = np.linalg.lstsq(X, np.r_[q0, qf, qd0, qdf, 0, 0], rcond=None)[0]
    coeffs_d = coeffs[0:5] * np.arange(5, 0, -1)
    coeffs_dd = coeffs_d[0:4] * np.arange(4, 0, -1)
    return t, np.polyval(coeffs, t), np.polyval(coeffs_d, t), np.polyval(coeffs_dd, t)


def trapezoidal(q0, qf, t, V=None):
    """tools/trajectory.py:488-505 + trapezoidal_func 552-607, restated: returns (t, s, sd, sdd, tb)."""
    t = np.arange(0, t) if isinstance(t, (int, np.integer)) else _f64(t).ravel()
    T = max(t)
    if V is None:
        V = (qf - q0) / T * 1.5
    else:
        V = abs(V) * np.sign(qf - q0)
        if abs(V) < (abs(qf - q0) / T):
            raise ValueError("V too small")
        elif abs(V) > (2 * abs(qf - q0) / T):
            raise ValueError("V too big")
    if V == 0:
        tb, a = np.inf, 0
    else:
        tb = (q0 - qf + V * T) / V
        a = V / tb
    p, pd, pdd = [], [], []
    for tk in t:
        if tk < 0:
            pk, pdk, pddk = q0, 0, 0
        elif tk <= tb:
            pk, pdk, pddk = q0 + a / 2 * tk**2, a * tk, a
        elif tk <= (T - tb):
            pk, pdk, pddk = (qf + q0 - V * T) / 2 + V * tk, V, 0
        elif tk <= T:
            pk, pdk, pddk = qf - a / 2 * T**2 + a * T * tk - a / 2 * tk**2, a * T - a * tk, -a
        else:
            pk, pdk, pddk = qf, 0, 0
        p.append(pk); pd.append(pdk); pdd.append(pddk)
    return t, np.array(p), np.array(pd), np.array(pdd), tb


def manip_svd(J, axes=(True,) * 6, kind="minsingular"):
    """ETS.py:1789-1796: minsingular = svd(Ja)[-1]; invcondition = 1 / cond(Ja)  (numpy, as the reference)."""
    J = np.asarray(J, dtype=np.float64)
    if J.ndim == 2:
        J = J[None]
    ax = np.asarray(axes, dtype=bool)
    out = np.empty(J.shape[0])
    for k, Jk in enumerate(J):
        s = np.linalg.svd(Jk[ax, :], compute_uv=False)
        out[k] = s[-1] if kind == "minsingular" else (s[-1] / s[0] if s[0] > 0 else 0.0)
    return out


# ------------------------------------------------------------------ pose representations (spatialmath conventions, restated)
# tr2rpy / tr2eul / trlog / rotvelxform / qslerp live in spatialmath-python, which is NOT under /root/reference: these
# are restatements of its documented definitions (parity unpinned against the package itself; the tests pin them by
# derivative / end-point identities).
_EPS = np.finfo(np.float64).eps


def tr2rpy(R, order="zyx"):
    """spatialmath.base.tr2rpy: 'zyx' R = Rz(yaw) Ry(pitch) Rx(roll); 'xyz' R = Rx(yaw) Ry(pitch) Rz(roll); -> (roll, pitch, yaw)"""
    R = np.asarray(R, dtype=np.float64)[:3, :3]
    rpy = np.zeros(3)
    if order == "xyz":
        if abs(abs(R[0, 2]) - 1) < 10 * _EPS:
            rpy[0] = 0
            rpy[2] = np.arctan2(R[2, 1], R[1, 1]) if R[0, 2] > 0 else -np.arctan2(R[1, 0], R[2, 0])
            rpy[1] = np.arcsin(np.clip(R[0, 2], -1.0, 1.0))
        else:
            rpy[0] = -np.arctan2(R[0, 1], R[0, 0])
            rpy[2] = -np.arctan2(R[1, 2], R[2, 2])
            k = int(np.argmax(np.abs([R[0, 0], R[0, 1], R[1, 2], R[2, 2]])))
            if k == 0:
                rpy[1] = np.arctan(R[0, 2] * np.cos(rpy[0]) / R[0, 0])
            elif k == 1:
                rpy[1] = -np.arctan(R[0, 2] * np.sin(rpy[0]) / R[0, 1])
            elif k == 2:
                rpy[1] = -np.arctan(R[0, 2] * np.sin(rpy[2]) / R[1, 2])
            else:
                rpy[1] = np.arctan(R[0, 2] * np.cos(rpy[2]) / R[2, 2])
    else:
        if abs(abs(R[2, 0]) - 1) < 10 * _EPS:
            rpy[0] = 0
            rpy[2] = -np.arctan2(R[0, 1], R[0, 2]) if R[2, 0] < 0 else np.arctan2(-R[0, 1], -R[0, 2])
            rpy[1] = -np.arcsin(np.clip(R[2, 0], -1.0, 1.0))
        else:
            rpy[0] = np.arctan2(R[2, 1], R[2, 2])
            rpy[2] = np.arctan2(R[1, 0], R[0, 0])
            k = int(np.argmax(np.abs([R[0, 0], R[1, 0], R[2, 1], R[2, 2]])))
            if k == 0:
                rpy[1] = -np.arctan(R[2, 0] * np.cos(rpy[2]) / R[0, 0])
            elif k == 1:
                rpy[1] = -np.arctan(R[2, 0] * np.sin(rpy[2]) / R[1, 0])
            elif k == 2:
                rpy[1] = -np.arctan(R[2, 0] * np.sin(rpy[0]) / R[2, 1])
            else:
                rpy[1] = -np.arctan(R[2, 0] * np.cos(rpy[0]) / R[2, 2])
    return rpy


def tr2eul(R):
    """spatialmath.base.tr2eul: R = Rz(phi) Ry(theta) Rz(psi)"""
    R = np.asarray(R, dtype=np.float64)[:3, :3]
    eul = np.zeros(3)
    if abs(R[0, 2]) < 10 * _EPS and abs(R[1, 2]) < 10 * _EPS:
        sp, cp = 0.0, 1.0
    else:
        eul[0] = np.arctan2(R[1, 2], R[0, 2])
        sp, cp = np.sin(eul[0]), np.cos(eul[0])
    eul[1] = np.arctan2(cp * R[0, 2] + sp * R[1, 2], R[2, 2])
    eul[2] = np.arctan2(-sp * R[0, 0] + cp * R[1, 0], -sp * R[0, 1] + cp * R[1, 1])
    return eul


def trlog(R):
    """exponential coordinates theta * axis of a rotation matrix"""
    R = np.asarray(R, dtype=np.float64)[:3, :3]
    li = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    ln = np.linalg.norm(li)
    tr = np.trace(R)
    if ln < 1e-9:
        if tr > 0:
            return np.zeros(3)
        w, V = np.linalg.eigh((R + R.T) / 2)
        return np.pi * V[:, -1]
    return np.arctan2(ln, tr - 1) * li / ln


def r2x(R, representation):
    return {"rpy/xyz": lambda: tr2rpy(R, "xyz"), "rpy/zyx": lambda: tr2rpy(R, "zyx"), "eul": lambda: tr2eul(R),
            "exp": lambda: trlog(R)}[representation]()


def rotvelxform_inv(g, representation):
    """rotvelxform(gamma, inverse=True): angular velocity -> rates of the representation"""
    if representation == "rpy/zyx":
        sb, cb, sg, cg = np.sin(g[1]), np.cos(g[1]), np.sin(g[2]), np.cos(g[2])
        A = np.array([[cb * cg, -sg, 0], [cb * sg, cg, 0], [-sb, 0, 1]])
        return np.linalg.inv(A)
    if representation == "rpy/xyz":
        sb, cb, sg, cg = np.sin(g[1]), np.cos(g[1]), np.sin(g[2]), np.cos(g[2])
        A = np.array([[sb, 0, 1], [-cb * sg, cg, 0], [cb * cg, sg, 0]])
        return np.linalg.inv(A)
    if representation == "eul":
        sp, cp, st, ct = np.sin(g[0]), np.cos(g[0]), np.sin(g[1]), np.cos(g[1])
        A = np.array([[0, -sp, cp * st], [0, cp, sp * st], [1, 0, ct]])
        return np.linalg.inv(A)
    th = np.linalg.norm(g)
    sk = np.array([[0, -g[2], g[1]], [g[2], 0, -g[0]], [-g[1], g[0], 0]])
    if th < 1e-8:
        return np.eye(3) - sk / 2
    A = np.eye(3) + sk * (1 - np.cos(th)) / th**2 + sk @ sk * (th - np.sin(th)) / th**3
    return np.linalg.inv(A)


def jacob0_analytical(T, J, representation):
    """ETS.py:1617-1624: A @ J with A = blkdiag(I, rotvelxform(R, inverse=True))"""
    T = np.asarray(T).reshape(-1, 4, 4)
    J = np.asarray(J).reshape(T.shape[0], 6, -1)
    out = J.copy()
    for k in range(T.shape[0]):
        out[k, 3:] = rotvelxform_inv(r2x(T[k, :3, :3], representation), representation) @ J[k, 3:]
    return out


def p_servo_rpy(Te, Tep, gain, threshold):
    """tools/p_servo.py:80-106, method='rpy'"""
    Te = np.asarray(Te).reshape(-1, 4, 4)
    Tep = np.broadcast_to(np.asarray(Tep).reshape(-1, 4, 4), Te.shape)
    K = np.eye(6) * gain if np.isscalar(gain) else np.diag(gain)
    v = np.empty((Te.shape[0], 6))
    arrived = np.empty(Te.shape[0], dtype=bool)
    for i in range(Te.shape[0]):
        eTep = np.linalg.inv(Te[i]) @ Tep[i]
        e = np.r_[eTep[:3, -1], tr2rpy(eTep, "zyx")]
        v[i] = K @ e
        arrived[i] = np.sum(np.abs(e)) < threshold
    return v, arrived


def _r2q(R):
    """unit quaternion (s, v) of a rotation matrix, s >= 0"""
    R = np.asarray(R, dtype=np.float64)[:3, :3]
    w, V = np.linalg.eigh(np.array([  # Bar-Itzhack: the dominant eigenvector of K is the quaternion
        [R[0, 0] + R[1, 1] + R[2, 2], R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]],
        [R[2, 1] - R[1, 2], R[0, 0] - R[1, 1] - R[2, 2], R[0, 1] + R[1, 0], R[0, 2] + R[2, 0]],
        [R[0, 2] - R[2, 0], R[0, 1] + R[1, 0], R[1, 1] - R[0, 0] - R[2, 2], R[1, 2] + R[2, 1]],
        [R[1, 0] - R[0, 1], R[0, 2] + R[2, 0], R[1, 2] + R[2, 1], R[2, 2] - R[0, 0] - R[1, 1]]]) / 3.0)
    q = V[:, -1]
    return q if q[0] >= 0 else -q


def _q2r(q):
    s, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - s * z), 2 * (x * z + s * y)],
                     [2 * (x * y + s * z), 1 - 2 * (x * x + z * z), 2 * (y * z - s * x)],
                     [2 * (x * z - s * y), 2 * (y * z + s * x), 1 - 2 * (x * x + y * y)]])


def ctraj_poses(T0, T1, s):
    """SE3.interp(T1, s) -> trinterp: translation lerp + qslerp(shortest=True); s clipped to [0, 1]"""
    T0, T1 = np.asarray(T0, dtype=np.float64), np.asarray(T1, dtype=np.float64)
    q0, q1 = _r2q(T0), _r2q(T1)
    out = np.zeros((len(s), 4, 4))
    for i, u in enumerate(np.clip(np.asarray(s, dtype=np.float64), 0, 1)):
        if u == 0:
            q = q0
        elif u == 1:
            q = q1
        else:
            a, dot = q0, float(np.dot(q0, q1))
            if dot < 0:
                a, dot = -q0, -dot
            th = np.arccos(np.clip(dot, -1, 1))
            q = (a * np.sin((1 - u) * th) + q1 * np.sin(u * th)) / np.sin(th) if abs(th) > 10 * _EPS else a
        out[i, :3, :3] = _q2r(q)
        out[i, :3, 3] = T0[:3, 3] * (1 - u) + u * T1[:3, 3]
        out[i, 3, 3] = 1
    return out


def mstraj(viapoints, dt, tacc, qdmax=None, tsegment=None, q0=None, qd0=None, qdf=None):
    """tools/trajectory.py:852-1152 restated in numpy (blends through this module's jtraj): returns (t, q, arrive)"""
    import math

    viapoints = np.asarray(viapoints, dtype=np.float64)
    if q0 is None:
        q0, viapoints = viapoints[0, :], viapoints[1:, :]
    q0 = np.asarray(q0, dtype=np.float64)
    ns, nj = viapoints.shape
    if tsegment is None:
        qdmax = np.tile(float(qdmax), (nj,)) if np.isscalar(qdmax) else np.asarray(qdmax, dtype=np.float64)
    Tacc = np.tile(float(tacc), (ns,)) if np.isscalar(tacc) else np.asarray(tacc, dtype=np.float64)
    qd0 = np.zeros(nj) if qd0 is None else np.asarray(qd0, dtype=np.float64)
    qdf = np.zeros(nj) if qdf is None else np.asarray(qdf, dtype=np.float64)

    def mrange(start, stop, step):
        return np.arange(round(start / step), round(stop / step) + 1) * step

    q_prev, qd_prev = q0, qd0
    clock = 0.0
    arrive = np.zeros(ns)
    tg = np.zeros((0, nj))
    tacc2 = 0.0
    q_next = q_prev
    for seg in range(ns):
        q_next = viapoints[seg, :]
        tacc_ = math.ceil(Tacc[seg] / dt) * dt
        tacc2 = math.ceil(tacc_ / 2 / dt) * dt
        taccx = tacc2 if seg == 0 else tacc_
        dq = q_next - q_prev
        if qdmax is not None:
            tl = np.ceil(np.abs(dq) / qdmax / dt) * dt
            tt = taccx + tl
            tseg = tt[int(np.argmax(tt))]
            if tseg <= 2 * tacc_:
                tseg = 2 * tacc_
        else:
            tseg = tsegment[seg]
        arrive[seg] = clock + tseg + (tacc2 if seg > 0 else 0.0)
        qd = dq / tseg
        if taccx > 0:
            qb = jtraj(q0, q_prev + tacc2 * qd, mrange(0, taccx, dt), qd0=qd_prev, qd1=qd)[1]
            tg = np.vstack([tg, qb[1:, :]])
        clock += taccx
        for t in mrange(tacc2 + dt, tseg - tacc2, dt):
            s = t / tseg
            q0 = (1 - s) * q_prev + s * q_next
            tg = np.vstack([tg, q0])
            clock += dt
        q_prev, qd_prev = q_next, qd
    if tacc2 > 0:
        qb = jtraj(q0, q_next, mrange(0, tacc2, dt), qd0=qd_prev, qd1=qdf)[1]
        tg = np.vstack([tg, qb[1:, :]])
    return dt * np.arange(0, tg.shape[0]), tg, arrive


# ------------------------------------------------------------------ Robot.rne (rigid-body trees), reference Robot.py:1704-1903
# Spatial-vector helpers as spatialmath defines them ([linear; angular] order): SE3.Ad, SpatialVelocity.cross (crm),
# its force dual (crf = -crm^T), SpatialInertia(m, r) WITHOUT rotational inertia (the reference passes only m and r).
def _skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], dtype=np.float64)


def _Ad(T):
    R, p = T[:3, :3], T[:3, 3]
    A = np.zeros((6, 6))
    A[:3, :3] = R
    A[:3, 3:] = _skew(p) @ R
    A[3:, 3:] = R
    return A


def _crm(v):
    X = np.zeros((6, 6))
    X[:3, :3] = _skew(v[3:]); X[:3, 3:] = _skew(v[:3]); X[3:, 3:] = _skew(v[3:])
    return X


def spatial_inertia(m, r):
    C = _skew(np.asarray(r, dtype=np.float64))
    return np.block([[m * np.eye(3), m * C.T], [m * C, m * C @ C.T]])


_AXIS_S = {0: [0, 0, 0, 1, 0, 0], 1: [0, 0, 0, 0, 1, 0], 2: [0, 0, 0, 0, 0, 1], 3: [1, 0, 0, 0, 0, 0], 4: [0, 1, 0, 0, 0, 0],
           5: [0, 0, 1, 0, 0, 0]}


def _joint_T(axis, eta):
    c, s = np.cos(eta), np.sin(eta)
    T = np.eye(4)
    if axis == 0:
        T[:3, :3] = [[1, 0, 0], [0, c, -s], [0, s, c]]
    elif axis == 1:
        T[:3, :3] = [[c, 0, s], [0, 1, 0], [-s, 0, c]]
    elif axis == 2:
        T[:3, :3] = [[c, -s, 0], [s, c, 0], [0, 0, 1]]
    else:
        T[axis - 3, 3] = eta
    return T


def tree_rne(tree, q, qd, qdd, gravity):
    """tree: dict(parent (n), axis, flip, jindex, C (n,4,4) constant part of each group, I6 (n,6,6)); gravity = the
    robot's gravity (NOT negated).  Follows Robot.rne step for step: Xup = (C J(q))^-1, v/a forward, f = I a + v x* I v,
    Q[j] = s . f, f[parent] += Xup^T f; s ignores flip (ET.s), torques in group order."""
    q, qd, qdd = (np.atleast_2d(np.asarray(x, dtype=np.float64)) for x in (q, qd, qdd))
    n = len(tree["parent"])
    a_grav = -np.r_[np.asarray(gravity, dtype=np.float64), 0, 0, 0]
    Q = np.zeros((q.shape[0], n))
    for k in range(q.shape[0]):
        Xup, v, a, f = [None] * n, [None] * n, [None] * n, [None] * n
        for j in range(n):
            ji = tree["jindex"][j]
            s = np.array(_AXIS_S[tree["axis"][j]], dtype=np.float64)
            eta = -q[k, ji] if tree["flip"][j] else q[k, ji]
            T = np.asarray(tree["C"][j]) @ _joint_T(tree["axis"][j], eta)
            Xup[j] = _Ad(np.linalg.inv(T))
            vJ = s * qd[k, ji]
            pa = tree["parent"][j]
            if pa < 0:
                v[j] = vJ
                a[j] = Xup[j] @ a_grav + s * qdd[k, ji]
            else:
                v[j] = Xup[j] @ v[pa] + vJ
                a[j] = Xup[j] @ a[pa] + s * qdd[k, ji] + _crm(v[j]) @ vJ
            I = np.asarray(tree["I6"][j])
            f[j] = I @ a[j] + (-_crm(v[j]).T) @ (I @ v[j])
        for j in reversed(range(n)):
            Q[k, j] = np.sum(f[j] * np.array(_AXIS_S[tree["axis"][j]]))
            pa = tree["parent"][j]
            if pa >= 0:
                f[pa] = f[pa] + Xup[j].T @ f[j]
    return Q


def fdyn(accel_fn, n, T, q0, qd0=None, torque_fn=None, solver="RK45", solver_args=None, dt=None):
    """DynamicsMixin.fdyn restated (Dynamics.py:300-377): scipy's integrator stepping on x = [q, qd] with
    xd = [qd, accel(q, qd, tau)]; accel_fn(q, qd, tau) -> qdd is the oracle's accel.  Returns (t, q, qd)."""
    from scipy import integrate, interpolate

    q0 = np.asarray(q0, dtype=np.float64).reshape(n)
    qd0 = np.zeros(n) if qd0 is None else np.asarray(qd0, dtype=np.float64).reshape(n)

    def f(t, x):
        q, qd = x[:n], x[n:]
        tau = np.zeros(n) if torque_fn is None else np.asarray(torque_fn(t, q, qd), dtype=np.float64)
        return np.r_[qd, accel_fn(q, qd, tau)]

    integ = integrate.__dict__[solver](f, t0=0.0, y0=np.r_[q0, qd0], t_bound=T, **(solver_args or {}))
    tl, xl = [0], [np.r_[q0, qd0]]
    while integ.status == "running":
        integ.step()
        if integ.status == "failed":
            raise RuntimeError("integration completed with failed status ")
        tl.append(integ.t)
        xl.append(integ.y)
    ta, xa = np.array(tl), np.array(xl)
    if dt is not None:
        tnew = np.arange(0, T, dt)
        xnew = interpolate.interp1d(ta, xa, axis=0)(tnew)
        return tnew, xnew[:, :n], xnew[:, n:]
    return ta, xa[:, :n], xa[:, n:]
