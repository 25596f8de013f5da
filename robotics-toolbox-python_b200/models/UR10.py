"""Universal Robots UR10, standard DH (reference models/DH/UR10.py:52-118).
No joint limits (so [-pi, pi]), gear ratio 1, no friction."""
from math import pi

import numpy as np

from ..DHLink import RevoluteDH
from ..DHRobot import DHRobot


class UR10(DHRobot):
    def __init__(self):
        deg = pi / 180
        a = [0, -0.612, -0.5723, 0, 0, 0]
        d = [0.1273, 0, 0, 0.163941, 0.1157, 0.0922]
        alpha = [pi / 2, 0.0, 0.0, pi / 2, -pi / 2, 0.0]
        mass = [7.1, 12.7, 4.27, 2.000, 2.000, 0.365]
        com = [[0.021, 0, 0.027], [0.38, 0, 0.158], [0.24, 0, 0.068], [0.0, 0.007, 0.018],
               [0.0, 0.007, 0.018], [0, 0, -0.026]]
        inertia = [
            np.array([[0.0341, 0, -0.0043], [0, 0.0353, 0.0001], [-0.0043, 0.0001, 0.0216]]),
            np.array([[0.0281, 0.0001, -0.0156], [0.0001, 0.7707, 0], [-0.0156, 0, 0.7694]]),
            np.array([[0.0101, 0.0001, 0.0092], [0.0001, 0.3093, 0], [0.0092, 0, 0.3065]]),
            np.array([[0.0030, -0.0000, 0], [-0.0000, 0.0022, -0.0002], [0, -0.0002, 0.0026]]),
            np.array([[0.0030, -0.0000, 0], [-0.0000, 0.0022, -0.0002], [0, -0.0002, 0.0026]]),
            np.array([[0, 0, 0], [0, 0.0004, 0], [0, 0, 0.0003]]),
        ]
        links = [RevoluteDH(d=d[j], a=a[j], alpha=alpha[j], m=mass[j], r=com[j], G=1, I=inertia[j]) for j in range(6)]
        super().__init__(links, name="UR10", manufacturer="Universal Robotics")
        self.addconfiguration("qr", np.array([180, 0, 0, 0, 90, 0]) * deg)
        self.addconfiguration("qz", np.zeros(6))
