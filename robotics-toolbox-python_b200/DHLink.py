"""Denavit-Hartenberg links -- host-side description
(reference src/roboticstoolbox/robot/DHLink.py:60-225, Link.py:120-190 for the dynamic parameters)."""
from __future__ import annotations

import numpy as np

from .ET import ET
from .ETS import ETS


def _inertia(I):
    """(3,3) / 9 / 6 / 3 -> 3x3 (reference Link.py:720-747; 6-vector = [Ixx Iyy Izz Ixy Iyz Ixz])."""
    I = np.asarray(I, dtype=np.float64)
    if I.shape == (3, 3):
        if np.any(np.abs(I - I.T) > 1e-8):
            raise ValueError("3x3 matrix is not symmetric")
        return I.copy()
    I = I.ravel()
    if I.size == 9:
        I = I.reshape(3, 3)
        if np.any(np.abs(I - I.T) > 1e-8):
            raise ValueError("3x3 matrix is not symmetric")
        return I.copy()
    if I.size == 6:
        return np.array([[I[0], I[3], I[5]], [I[3], I[1], I[4]], [I[5], I[4], I[2]]])
    if I.size == 3:
        return np.diag(I)
    raise ValueError("invalid shape passed: must be (3,3), (6,), (3,)")


class DHLink:
    """One DH / MDH link with its dynamic parameters (reference DHLink.py:60-171)."""

    def __init__(self, d=0.0, alpha=0.0, theta=0.0, a=0.0, sigma=0, mdh=False, offset=0.0, flip=False,
                 qlim=None, m=None, r=None, I=None, Jm=None, B=None, Tc=None, G=None, name=None):
        self._robot = None
        self.ets = None  # built at the end of __init__; the kinematic setters below rebuild it
        self._d, self._alpha, self._theta, self._a = float(d), float(alpha), float(theta), float(a)
        self._sigma = int(sigma)
        self._mdh = bool(mdh)
        self._offset = float(offset)
        self._flip = bool(flip)
        self._qlim = None if qlim is None else np.asarray(qlim, dtype=np.float64).reshape(2)
        self.name = name
        # dynamic parameters with the reference defaults (Link.py:172-184)
        self._m = 0.0 if m is None else float(m)
        self._r = np.zeros(3) if r is None else np.asarray(r, dtype=np.float64).reshape(3)
        self._I = np.zeros((3, 3)) if I is None else _inertia(I)
        self._Jm = 0.0 if Jm is None else float(Jm)
        self._B = 0.0 if B is None else float(B)
        self._G = 0.0 if G is None else float(G)
        self._Tc = self._tc(np.zeros(2) if Tc is None else Tc)
        self.ets = self._to_ets()

    @staticmethod
    def _tc(Tc):
        Tc = np.asarray(Tc, dtype=np.float64).ravel()
        if Tc.size == 1:  # symmetric Coulomb friction (reference Link.py:850-856)
            return np.array([Tc[0], -Tc[0]])
        return Tc.reshape(2).copy()

    # every dynamic-parameter setter marks the owning robot dirty so the packed RNE table is rebuilt,
    # the reference's @_listen_dyn -> robot.dynchanged() protocol (Link.py:26-59, DHLink.py:24-51)
    def _dirty(self):
        if self._robot is not None:
            self._robot.dynchanged()

    def _dynprop(name, conv):  # noqa: N805
        def get(self):
            return getattr(self, "_" + name)

        def set(self, v):
            setattr(self, "_" + name, conv(v))
            self._dirty()

        return property(get, set)

    # Kinematic parameters: the reference decorates theta, d, a, alpha, sigma, mdh (and offset) with
    # @_listen_dyn (DHLink.py:448-563) so a change re-serialises the frne table.  Here a change must ALSO
    # rebuild this link's elementary transforms and drop the robot's compiled chain (its device handle is
    # released with the old ETS object), otherwise fkine / jacob / ik would keep walking the old geometry.
    def _kinprop(name, conv):  # noqa: N805
        def get(self):
            return getattr(self, "_" + name)

        def set(self, v):
            setattr(self, "_" + name, conv(v))
            if self.ets is not None:
                self.ets = self._to_ets()
            if self._robot is not None:
                self._robot._kinchanged()

        return property(get, set)

    d = _kinprop("d", float)
    a = _kinprop("a", float)
    alpha = _kinprop("alpha", float)
    theta = _kinprop("theta", float)
    offset = _kinprop("offset", float)
    sigma = _kinprop("sigma", int)
    mdh = _kinprop("mdh", bool)
    flip = _kinprop("flip", bool)
    qlim = _kinprop("qlim", lambda v: None if v is None else np.asarray(v, dtype=np.float64).reshape(2))
    del _kinprop

    m = _dynprop("m", float)
    r = _dynprop("r", lambda v: np.asarray(v, dtype=np.float64).reshape(3))
    I = _dynprop("I", _inertia)  # noqa: E741
    Jm = _dynprop("Jm", float)
    B = _dynprop("B", float)
    G = _dynprop("G", float)
    Tc = _dynprop("Tc", lambda v: DHLink._tc(v))
    del _dynprop

    @property
    def isrevolute(self):
        return self.sigma == 0

    @property
    def isprismatic(self):
        return self.sigma == 1

    def _to_ets(self) -> ETS:
        """DH -> elementary transforms, zero terms omitted (reference DHLink._to_ets, DHLink.py:173-225)."""
        ets = []
        rev = self.sigma == 0
        kw = dict(flip=self.flip, qlim=self.qlim)
        if self.mdh:
            if self.a != 0:
                ets.append(ET.tx(self.a))
            if self.alpha != 0:
                ets.append(ET.Rx(self.alpha))
            if rev:
                if self.offset != 0:
                    ets.append(ET.Rz(self.offset))
                if self.d != 0:
                    ets.append(ET.tz(self.d))
                ets.append(ET.Rz(**kw))
            else:
                if self.theta != 0:
                    ets.append(ET.Rz(self.theta))
                if self.offset != 0:
                    ets.append(ET.tz(self.offset))
                ets.append(ET.tz(**kw))
        else:
            if rev:
                if self.offset != 0:
                    ets.append(ET.Rz(self.offset))
                ets.append(ET.Rz(**kw))
                if self.d != 0:
                    ets.append(ET.tz(self.d))
            else:
                if self.theta != 0:
                    ets.append(ET.Rz(self.theta))
                if self.offset != 0:
                    ets.append(ET.tz(self.offset))
                ets.append(ET.tz(**kw))
            if self.a != 0:
                ets.append(ET.tx(self.a))
            if self.alpha != 0:
                ets.append(ET.Rx(self.alpha))
        return ETS(ets)


class RevoluteDH(DHLink):
    def __init__(self, d=0.0, a=0.0, alpha=0.0, offset=0.0, qlim=None, flip=False, **kw):
        super().__init__(d=d, a=a, alpha=alpha, theta=0.0, sigma=0, mdh=False, offset=offset, qlim=qlim, flip=flip, **kw)


class PrismaticDH(DHLink):
    def __init__(self, theta=0.0, a=0.0, alpha=0.0, offset=0.0, qlim=None, flip=False, **kw):
        super().__init__(theta=theta, a=a, alpha=alpha, d=0.0, sigma=1, mdh=False, offset=offset, qlim=qlim, flip=flip, **kw)


class RevoluteMDH(DHLink):
    def __init__(self, d=0.0, a=0.0, alpha=0.0, offset=0.0, qlim=None, flip=False, **kw):
        super().__init__(d=d, a=a, alpha=alpha, theta=0.0, sigma=0, mdh=True, offset=offset, qlim=qlim, flip=flip, **kw)


class PrismaticMDH(DHLink):
    def __init__(self, theta=0.0, a=0.0, alpha=0.0, offset=0.0, qlim=None, flip=False, **kw):
        super().__init__(theta=theta, a=a, alpha=alpha, d=0.0, sigma=1, mdh=True, offset=offset, qlim=qlim, flip=flip, **kw)
