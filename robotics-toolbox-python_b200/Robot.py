"""ETS-based robots: the thin delegations of the reference's RobotKinematicsMixin
(reference src/roboticstoolbox/robot/RobotKinematics.py:28-97 fkine, 158 jacob0, 219 jacobe,
736-746 ik_LM, 1209-1226 ikine_LM) over a serial chain of links
(reference Link.py / BaseRobot.ets(), BaseRobot.py:1554-1652, for the unbranched case)."""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from ._se3 import SE3
from .ET import ET
from .ETS import ETS, _mat44


class Link:
    """A rigid link = an ETS ending in at most one joint (reference Link.py:120-215), with a parent."""

    def __init__(self, ets=None, name: Optional[str] = None, parent: Optional["Link"] = None, **kwargs):
        if ets is None:
            ets = ETS()
        elif isinstance(ets, ET):
            ets = ETS(ets)
        elif not isinstance(ets, ETS):
            raise TypeError("The ets argument must be of type ETS or ET")
        if ets.n > 1 or (ets.n == 1 and not ets[-1].isjoint):
            raise ValueError("a Link's ETS may contain one joint, as its last transform")
        self.ets = ets
        self.name = name
        self.parent = parent

    @property
    def isjoint(self) -> bool:
        return self.ets.n == 1


class Robot:
    """A serial ETS robot (reference Robot / ERobot for an unbranched tree)."""

    def __init__(self, arg, name: str = "", manufacturer: str = "", base=None, tool=None, **kwargs):
        if isinstance(arg, ETS):
            self.links = self._split(arg)
        elif isinstance(arg, (list, tuple)) and all(isinstance(l, Link) for l in arg):
            self.links = list(arg)
        else:
            raise TypeError("arg must be an ETS or a list of Link")
        self.name, self.manufacturer = name, manufacturer
        self._sub_ets = {}
        self._T = np.eye(4) if base is None else _mat44(base, "base")
        self._tool = None if tool is None else _mat44(tool, "tool")
        self._ets = None
        self._configs = {}

    @staticmethod
    def _split(ets: ETS) -> List[Link]:
        links, cur, parent = [], [], None
        for e in ets:
            cur.append(e)
            if e.isjoint:
                parent = Link(ETS(cur), name=f"link{len(links)}", parent=parent)
                links.append(parent)
                cur = []
        if cur:
            links.append(Link(ETS(cur), name="ee", parent=parent))
        return links

    @property
    def base(self) -> SE3:
        return SE3(self._T)

    @base.setter
    def base(self, T):
        self._T = np.eye(4) if T is None else _mat44(T, "base")

    @property
    def n(self) -> int:
        return sum(1 for l in self.links if l.isjoint)

    @property
    def qlim(self):
        return self.ets().qlim

    def addconfiguration(self, name, q):
        self._configs[name] = np.asarray(q, dtype=np.float64)
        setattr(self, name, self._configs[name])

    def _getlink(self, link, default):
        """Link reference or name -> index into self.links (reference BaseRobot._getlink 1377-1424)."""
        if link is None:
            return default
        if isinstance(link, str):
            for i, l in enumerate(self.links):
                if l.name == link:
                    return i
            raise ValueError(f"no link named {link}")
        if isinstance(link, Link):
            for i, l in enumerate(self.links):
                if l is link:
                    return i
            raise ValueError("link not in robot links")
        raise TypeError("unknown argument")

    def ets(self, start=None, end=None) -> ETS:
        """``robot.ets()``: the chain from the base link to the end-effector; ``robot.ets(start=l1, end=l2)``: the
        kinematics from link ``l1`` to link ``l2`` (Link reference or name), start link included -- reference
        BaseRobot.ets 1554-1652 / _find_ets 1426-1467 for an unbranched tree.  The joints of a sub-chain keep the
        jindex they have in the whole robot, so q stays the robot's full joint vector.  A path that runs towards the
        base is the inverse of the corresponding forward range.  Branched trees are outside this repository's scope
        (SURVEY 8f row 4)."""
        if self._ets is None:
            self._ets = ETS.from_links([l.ets for l in self.links])
        if start is None and end is None:
            return self._ets
        i = self._getlink(start, 0)
        j = self._getlink(end, len(self.links) - 1)
        key = (i, j)
        if key not in self._sub_ets:
            offs = np.cumsum([0] + [len(l.ets) for l in self.links])
            if i <= j:  # towards the tip: start link's own transform included (_find_ets 1445-1453)
                self._sub_ets[key] = ETS([et.copy() for et in self._ets._ets[offs[i]:offs[j + 1]]])
            else:       # towards the base: inverted link transforms of links i .. j+1 (_find_ets 1457-1467)
                self._sub_ets[key] = ETS([et.copy() for et in self._ets._ets[offs[j + 1]:offs[i + 1]]]).inv()
        return self._sub_ets[key]

    def _base_arg(self):
        return None if np.array_equal(self._T, np.eye(4)) else self._T

    def _tool_arg(self, tool):
        if tool is not None:
            return tool
        return self._tool

    # RobotKinematics.py:92-97: fkine applies the robot's base
    def fkine(self, q, end=None, start=None, tool=None, include_base: bool = True, **kw) -> SE3:
        return self.ets(start, end).fkine(q, base=self._base_arg(), tool=self._tool_arg(tool), include_base=include_base, **kw)

    def eval(self, q, tool=None, include_base: bool = True, **kw):
        return self.ets().eval(q, base=self._base_arg(), tool=self._tool_arg(tool), include_base=include_base, **kw)

    # RobotKinematics.py:158 / 219: the Jacobians do NOT see the base
    def jacob0(self, q, end=None, start=None, tool=None, **kw):
        return self.ets(start, end).jacob0(q, tool=self._tool_arg(tool), **kw)

    def jacobe(self, q, end=None, start=None, tool=None, **kw):
        return self.ets(start, end).jacobe(q, tool=self._tool_arg(tool), **kw)

    def fkine_jacob0(self, q, tool=None, **kw):
        return self.ets().fkine_jacob0(q, base=self._base_arg(), tool=self._tool_arg(tool), **kw)

    def hessian0(self, q=None, J0=None, end=None, start=None, tool=None, **kw):
        return self.ets(start, end).hessian0(q, J0=J0, tool=self._tool_arg(tool), **kw)

    def hessiane(self, q=None, Je=None, end=None, start=None, tool=None, **kw):
        return self.ets(start, end).hessiane(q, Je=Je, tool=self._tool_arg(tool), **kw)

    def manipulability(self, q=None, J=None, method="yoshikawa", axes="all", **kw):
        return self.ets().manipulability(q, J=J, method=method, axes=axes, **kw)

    def jacobm(self, q=None, J=None, H=None, end=None, start=None, axes="all", **kw):
        """reference Robot.jacobm, Robot.py:1101-1232"""
        return self.ets(start, end).jacobm(q, J=J, H=H, axes=axes, **kw)

    def jacob0_dot(self, q, qd, J0=None, representation=None, **kw):
        """reference Robot.jacob0_dot, Robot.py:964-1099"""
        return self.ets().jacob0_dot(q, qd, J0=J0, representation=representation, **kw)

    def ik_LM(self, Tep, end=None, start=None, **kw):
        return self.ets(start, end).ik_LM(Tep, **kw)

    def ikine_LM(self, Tep, end=None, start=None, **kw):
        return self.ets(start, end).ikine_LM(Tep, **kw)

    # reference RobotKinematics.py:748-1027 (ik_NR, ik_GN), 1228-1525 (ikine_NR, ikine_GN)
    def ik_NR(self, Tep, end=None, start=None, **kw):
        return self.ets(start, end).ik_NR(Tep, **kw)

    def ik_GN(self, Tep, end=None, start=None, **kw):
        return self.ets(start, end).ik_GN(Tep, **kw)

    def ikine_NR(self, Tep, end=None, start=None, **kw):
        return self.ets(start, end).ikine_NR(Tep, **kw)

    def ikine_GN(self, Tep, end=None, start=None, **kw):
        return self.ets(start, end).ikine_GN(Tep, **kw)


ERobot = Robot
