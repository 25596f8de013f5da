"""Batched pose error and position-based servo (reference tools/p_servo.py).

``angle_axis(T, Td)`` is the batched ``fknm.Angle_Axis`` (fknm.cpp:112-162); ``p_servo`` is
tools/p_servo.py:46-106 for both of its methods: "angle-axis" (error in the base frame) and the reference's
default "rpy" (error in the end-effector frame, e = [t(Te^-1 Tep); tr2rpy(Te^-1 Tep, order="zyx")]).  ``tr2rpy``
belongs to spatialmath, which is not part of the reference tree: its zyx convention (R = Rz(yaw) Ry(pitch) Rx(roll),
result (roll, pitch, yaw), roll := 0 at the pitch = +-pi/2 singularity) is restated in csrc/b2k_pose.cu.
"""
from __future__ import annotations

import numpy as np

from . import _buffers as B
from . import _lib


def _poses(T, Td, dtype):
    T = getattr(T, "A", T)
    Td = getattr(Td, "A", Td)
    B.check_numeric(T, "T")
    B.check_numeric(Td, "Td")
    host = not (B.is_tensor(T) or B.is_tensor(Td))
    dt = B.pick_dtype(T if B.is_tensor(T) or isinstance(T, np.ndarray) else None, dtype)
    a = B.to_device(T, dt)
    b = B.to_device(Td, dt, device=a.device)
    single = a.dim() == 2 and b.dim() == 2
    if a.dim() == 2:
        a = a.reshape(1, 4, 4)
    if tuple(a.shape[1:]) != (4, 4) or tuple(b.shape[-2:]) != (4, 4):
        raise ValueError("poses must be (4,4) or (N,4,4)")
    N = a.shape[0]
    if b.dim() == 2 or b.shape[0] == 1:
        stride = 0
    elif b.shape[0] == N:
        stride = 16
    elif N == 1:  # one current pose against N targets
        a = a.expand(b.shape[0], 4, 4).contiguous()
        N, stride = b.shape[0], 16
    else:
        raise ValueError("T and Td must have the same number of poses (or one of them a single pose)")
    return a.contiguous(), b.contiguous(), N, stride, dt, host, single


def angle_axis(T, Td, dtype=None):
    """6-vector pose error(s) [translation; angle * axis] between T and Td (reference tools/p_servo.py:15-45)."""
    a, b, N, stride, dt, host, single = _poses(T, Td, dtype)
    e = B.empty((N, 6), dt, like=a)
    _lib.check(_lib.lib().b2k_angle_axis(B.code(dt), B.ptr(a), B.ptr(b), N, stride, B.ptr(e), B.stream_ptr(a)))
    if host:
        e = B.to_host(e)
    return e[0] if single else e


def p_servo(wTe, wTep, gain=1.0, threshold=0.1, method="rpy", dtype=None):
    """End-effector velocity that drives wTe towards wTep, and the `arrived` flag(s)
    (reference tools/p_servo.py:46-106).  Returns (v, arrived): (6,), bool for one pose pair;
    (N,6), (N,) bool for a batch."""
    if method not in ("rpy", "angle-axis"):
        raise ValueError("method must be 'rpy' or 'angle-axis'")
    a, b, N, stride, dt, host, single = _poses(wTe, wTep, dtype)
    g = np.full(6, float(gain)) if np.isscalar(gain) else np.ascontiguousarray(np.asarray(gain, dtype=np.float64).reshape(6))
    v = B.empty((N, 6), dt, like=a)
    arrived = B.empty_i32((N,), like=a)
    fn = _lib.lib().b2k_p_servo_rpy if method == "rpy" else _lib.lib().b2k_p_servo
    _lib.check(fn(B.code(dt), B.ptr(a), B.ptr(b), N, stride, _lib.dptr(g), float(threshold), B.ptr(v), B.ptr(arrived),
                  B.stream_ptr(a)))
    if host:
        v, arrived = B.to_host(v), B.to_host(arrived).astype(bool)
    else:
        arrived = arrived.bool()
    if single:
        return v[0], bool(arrived[0])
    return v, arrived
