"""Franka-Emika Panda, 7 DoF, as an elementary transform sequence
(reference models/ETS/Panda.py:28-64; the same chain is spelled out in tests/test_ETS.py:267-293).
No joint limits are set, so the reference's C defaults [-pi, pi] apply (ET.py:109-115)."""
import numpy as np

from ..ET import ET
from ..Robot import Link, Robot


class Panda(Robot):
    def __init__(self):
        deg = np.pi / 180
        mm = 1e-3
        tool_offset = 103 * mm
        l0 = Link(ET.tz(0.333) * ET.Rz(), name="link0", parent=None)
        l1 = Link(ET.Rx(-90 * deg) * ET.Rz(), name="link1", parent=l0)
        l2 = Link(ET.Rx(90 * deg) * ET.tz(0.316) * ET.Rz(), name="link2", parent=l1)
        l3 = Link(ET.tx(0.0825) * ET.Rx(90, "deg") * ET.Rz(), name="link3", parent=l2)
        l4 = Link(ET.tx(-0.0825) * ET.Rx(-90, "deg") * ET.tz(0.384) * ET.Rz(), name="link4", parent=l3)
        l5 = Link(ET.Rx(90, "deg") * ET.Rz(), name="link5", parent=l4)
        l6 = Link(ET.tx(0.088) * ET.Rx(90, "deg") * ET.tz(0.107) * ET.Rz(), name="link6", parent=l5)
        ee = Link(ET.tz(tool_offset) * ET.Rz(-np.pi / 4), name="ee", parent=l6)
        super().__init__([l0, l1, l2, l3, l4, l5, l6, ee], name="Panda", manufacturer="Franka Emika")
        self.addconfiguration("qr", np.array([0, -0.3, 0, -2.2, 0, 2.0, np.pi / 4]))
        self.addconfiguration("qz", np.zeros(7))
