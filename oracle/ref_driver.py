"""Drive the reference's OWN compiled extension modules (oracle/_ref/{fknm,frne}*.so) --
TEST INFRASTRUCTURE ONLY.

The reference's Python package cannot be imported in this image (its dependency
spatialmath-python is absent), but its native modules compile from the sources
under /root/reference (oracle/Makefile `ref` target) and can be driven bare.
This module restates just the marshalling the reference's Python layer does:

* ``fknm.ET_init(isstaticsym, isjoint, isflip, jindex, axis, T(F-order 4x4), qlim)``
  -- reference ET.py:100-125 (defaults: jindex 0 for constants, qlim [-pi,pi] / [0,1]);
  the C struct keeps a *borrowed* pointer into T (fknm.cpp:1207) so T is kept alive here.
* ``fknm.ETS_init(list_of_ET_capsules, n, m)`` -- reference ETS.py:62-69.
* ``frne.init(n, mdh, L, -gravity)`` / ``frne.frne(ob, q, qd, qdd, -gravity, fext)``
  -- reference DHRobot.py:1340-1361, 1442-1451.

Takes the same neutral chain description dict as oracle/oracle.py.
"""
from __future__ import annotations

import glob
import importlib.util
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")
_mods = {}


def available() -> bool:
    return bool(glob.glob(os.path.join(_REF, "fknm*.so"))) and bool(
        glob.glob(os.path.join(_REF, "frne*.so"))
    )


def _load(name):
    if name not in _mods:
        cands = glob.glob(os.path.join(_REF, name + "*.so"))
        if not cands:
            raise ImportError(
                f"oracle/_ref/{name}*.so not built; run `make -C oracle ref` where /root/reference exists"
            )
        spec = importlib.util.spec_from_file_location(name, cands[0])
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _mods[name] = mod
    return _mods[name]


def fknm():
    return _load("fknm")


def frne():
    return _load("frne")


class RefETS:
    """A reference `ETS` capsule built from a chain description."""

    def __init__(self, desc):
        f = fknm()
        self._keep = []
        caps = []
        isjoint = np.asarray(desc["isjoint"], dtype=int)
        m = len(isjoint)
        for i in range(m):
            T = np.asfortranarray(np.asarray(desc["T"], dtype=np.float64).reshape(m, 4, 4)[i])
            qlim = np.ascontiguousarray(np.asarray(desc["qlim"], dtype=np.float64).reshape(m, 2)[i])
            self._keep += [T, qlim]
            caps.append(
                f.ET_init(0, int(isjoint[i]), int(desc["flip"][i]), int(desc["jindex"][i]),
                          int(desc["axis"][i]), T, qlim)
            )
        self._caps = caps
        self.m = m
        self.n = int(isjoint.sum())
        self.ets = f.ETS_init(caps, self.n, self.m)

    def fkine(self, q, base=None, tool=None):
        """(N,4,4) like fknm.ETS_fkine's batch form; always returns 3-D."""
        q = np.ascontiguousarray(q, dtype=np.float64)
        T = fknm().ETS_fkine(self.ets, q, base, tool, 1)
        T = np.asarray(T)
        return T.reshape(1, 4, 4) if T.ndim == 2 else T

    def fkine_rows(self, q, base=None, tool=None):
        """Row-by-row calls: immune to the (N,1)/(1,n) shape sniffing (fknm.cpp:968-988)."""
        q = np.atleast_2d(np.ascontiguousarray(q, dtype=np.float64))
        f = fknm()
        return np.stack([np.asarray(f.ETS_fkine(self.ets, q[i], base, tool, 1)) for i in range(q.shape[0])])

    def jacob0(self, q, tool=None):
        q = np.atleast_2d(np.ascontiguousarray(q, dtype=np.float64))
        f = fknm()
        return np.stack([np.asarray(f.ETS_jacob0(self.ets, q[i], tool)) for i in range(q.shape[0])])

    def jacobe(self, q, tool=None):
        q = np.atleast_2d(np.ascontiguousarray(q, dtype=np.float64))
        f = fknm()
        return np.stack([np.asarray(f.ETS_jacobe(self.ets, q[i], tool)) for i in range(q.shape[0])])

    def hessian0(self, q, tool=None):
        q = np.atleast_2d(np.ascontiguousarray(q, dtype=np.float64))
        f = fknm()
        return np.stack([np.asarray(f.ETS_hessian0(self.ets, q[i], None, tool)) for i in range(q.shape[0])])

    def hessiane(self, q, tool=None):
        q = np.atleast_2d(np.ascontiguousarray(q, dtype=np.float64))
        f = fknm()
        return np.stack([np.asarray(f.ETS_hessiane(self.ets, q[i], None, tool)) for i in range(q.shape[0])])

    def ik_lm(self, Tep, q0=None, ilimit=30, slimit=100, tol=1e-6, joint_limits=True, mask=None,
              k=1.0, method="chan"):
        """Per-target loop over fknm.IK_LM_c (the reference has no batched IK)."""
        Tep = np.ascontiguousarray(Tep, dtype=np.float64).reshape(-1, 4, 4)
        N = Tep.shape[0]
        f = fknm()
        q = np.empty((N, self.n)); succ = np.empty(N, np.int32); its = np.empty(N, np.int32)
        srch = np.empty(N, np.int32); E = np.empty(N)
        if q0 is not None:
            q0 = np.broadcast_to(np.asarray(q0, dtype=np.float64).reshape(-1, self.n), (N, self.n))
        for i in range(N):
            qi0 = None if q0 is None else np.ascontiguousarray(q0[i])
            r = f.IK_LM_c(self.ets, Tep[i], qi0, ilimit, slimit, tol, int(bool(joint_limits)), mask,
                          float(k), method)
            q[i], succ[i], its[i], srch[i], E[i] = r
        return q, succ, its, srch, E


def _ik_pinv(ets, fn, Tep, q0, ilimit, slimit, tol, joint_limits, mask, pinv, pinv_damping):
    """Per-target loop over fknm.IK_NR_c / IK_GN_c (fknm.cpp:164-392): args
    (ets, Tep, q0, ilimit, slimit, tol, reject_jl, we, use_pinv, pinv_damping)."""
    Tep = np.ascontiguousarray(Tep, dtype=np.float64).reshape(-1, 4, 4)
    N = Tep.shape[0]
    q = np.empty((N, ets.n)); succ = np.empty(N, np.int32); its = np.empty(N, np.int32)
    srch = np.empty(N, np.int32); E = np.empty(N)
    if q0 is not None:
        q0 = np.broadcast_to(np.asarray(q0, dtype=np.float64).reshape(-1, ets.n), (N, ets.n))
    for i in range(N):
        qi0 = None if q0 is None else np.ascontiguousarray(q0[i])
        q[i], succ[i], its[i], srch[i], E[i] = fn(ets.ets, Tep[i], qi0, ilimit, slimit, tol,
                                                  int(bool(joint_limits)), mask, int(bool(pinv)),
                                                  float(pinv_damping))
    return q, succ, its, srch, E


def ik_nr(ets, Tep, q0=None, ilimit=30, slimit=100, tol=1e-6, joint_limits=True, mask=None, pinv=True,
          pinv_damping=0.0):
    return _ik_pinv(ets, fknm().IK_NR_c, Tep, q0, ilimit, slimit, tol, joint_limits, mask, pinv, pinv_damping)


def ik_gn(ets, Tep, q0=None, ilimit=30, slimit=100, tol=1e-6, joint_limits=True, mask=None, pinv=True,
          pinv_damping=0.0):
    return _ik_pinv(ets, fknm().IK_GN_c, Tep, q0, ilimit, slimit, tol, joint_limits, mask, pinv, pinv_damping)


def angle_axis(Te, Tep):
    return np.asarray(fknm().Angle_Axis(np.ascontiguousarray(Te, dtype=np.float64),
                                        np.ascontiguousarray(Tep, dtype=np.float64)))


class RefRNE:
    """A reference frne `Robot` capsule; grav is the robot's gravity (NOT negated)."""

    def __init__(self, n, mdh, L, gravity):
        self.n = n
        self.gravity = np.asarray(gravity, dtype=np.float64)
        self.ob = frne().init(n, int(mdh), [float(x) for x in np.asarray(L).ravel()],
                              [float(x) for x in -self.gravity])

    def rne(self, q, qd, qdd, gravity=None, fext=None):
        q = np.atleast_2d(np.asarray(q, dtype=np.float64))
        qd = np.atleast_2d(np.asarray(qd, dtype=np.float64))
        qdd = np.atleast_2d(np.asarray(qdd, dtype=np.float64))
        g = self.gravity if gravity is None else np.asarray(gravity, dtype=np.float64)
        fx = np.zeros(6) if fext is None else np.asarray(fext, dtype=np.float64)
        fr = frne().frne
        ng = -g
        out = np.empty((q.shape[0], self.n))
        for i in range(q.shape[0]):
            out[i] = fr(self.ob, q[i], qd[i], qdd[i], ng, fx)
        return out

    def __del__(self):
        try:
            frne().delete(self.ob)
        except Exception:
            pass
