"""Franka-Emika Panda in modified DH form with link inertias (reference models/DH/Panda.py:36-173)."""
import numpy as np

from ..DHLink import RevoluteMDH
from ..DHRobot import DHRobot


class PandaMDH(DHRobot):
    def __init__(self):
        pi = np.pi
        rows = [
            (0.0, 0.333, 0.0, [-2.8973, 2.8973], 4.970684, [7.03370e-01, 7.06610e-01, 9.11700e-03, -1.39000e-04, 1.91690e-02, 6.77200e-03]),
            (0.0, 0.0, -pi / 2, [-1.7628, 1.7628], 0.646926, [7.96200e-03, 2.81100e-02, 2.59950e-02, -3.92500e-03, 7.04000e-04, 1.02540e-02]),
            (0.0, 0.316, pi / 2, [-2.8973, 2.8973], 3.228604, [3.72420e-02, 3.61550e-02, 1.08300e-02, -4.76100e-03, -1.28050e-02, -1.13960e-02]),
            (0.0825, 0.0, pi / 2, [-3.0718, -0.0698], 3.587895, [2.58530e-02, 1.95520e-02, 2.83230e-02, 7.79600e-03, 8.64100e-03, -1.33200e-03]),
            (-0.0825, 0.384, -pi / 2, [-2.8973, 2.8973], 1.225946, [3.55490e-02, 2.94740e-02, 8.62700e-03, -2.11700e-03, 2.29000e-04, -4.03700e-03]),
            (0.0, 0.0, pi / 2, [-0.0175, 3.7525], 1.666555, [1.96400e-03, 4.35400e-03, 5.43300e-03, 1.09000e-04, 3.41000e-04, -1.15800e-03]),
            (0.088, 107 * 1e-3, pi / 2, [-2.8973, 2.8973], 7.35522e-01, [1.25160e-02, 1.00270e-02, 4.81500e-03, -4.28000e-04, -7.41000e-04, -1.19600e-03]),
        ]
        L = [RevoluteMDH(a=a, d=d, alpha=al, qlim=np.array(ql), m=m, I=I, G=1) for a, d, al, ql, m, I in rows]
        tool = np.eye(4)
        tool[2, 3] = 103 * 1e-3
        c, s = np.cos(-pi / 4), np.sin(-pi / 4)
        tool = tool @ np.array([[c, -s, 0, 0], [s, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1.0]])
        super().__init__(L, name="Panda", manufacturer="Franka Emika", tool=tool)
        self.addconfiguration("qr", np.array([0, -0.3, 0, -2.2, 0, 2.0, pi / 4]))
        self.addconfiguration("qz", np.zeros(7))
