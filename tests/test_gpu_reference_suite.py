"""The reference's own unit tests for the hot path, re-run against the CUDA library through the mirrored API
(`-m gpu`).  Each test names the reference test it follows; literal expected values are the reference's.
Helper formulas (numjac, tr2jac) restate spatialmath.base's definitions for the geometric Jacobian."""
import os
import sys

import numpy as np
import numpy.testing as nt
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import b2kin as rtb  # noqa: E402
from oracle import chains as ch  # noqa: E402  (trotx / transl helpers: plain numpy)


def tr2jac(T):
    """blkdiag(R^T, R^T) of a pose (spatialmath.base.tr2jac)."""
    R = np.asarray(T)[:3, :3]
    J = np.zeros((6, 6))
    J[:3, :3] = R.T
    J[3:, 3:] = R.T
    return J


def numjac(fk, q, h=1e-7):
    """Numerical geometric Jacobian of a pose function (spatialmath.base.numjac, SE=3)."""
    q = np.asarray(q, dtype=float)
    T0 = fk(q)
    J = np.zeros((6, len(q)))
    for i in range(len(q)):
        dq = np.zeros(len(q)); dq[i] = h
        Ti = fk(q + dq)
        J[:3, i] = (Ti[:3, 3] - T0[:3, 3]) / h
        S = (Ti[:3, :3] - T0[:3, :3]) / h @ T0[:3, :3].T  # skew(omega)
        J[3:, i] = [S[2, 1] - S[1, 2], S[0, 2] - S[2, 0], S[1, 0] - S[0, 1]]
        J[3:, i] *= 0.5
    return J


# ------------------------------------------------------------------ tests/test_ETS.py
def test_fkine_all_et_kinds():
    """tests/test_ETS.py:155-187 test_fkine"""
    q = np.array([1.0, 2.0, 3.0, 4.0, 5.0, 6.0])
    r = (rtb.ET.Rx(1.543) * rtb.ET.Ry(1.543) * rtb.ET.Rz(1.543) * rtb.ET.tx(1.543) * rtb.ET.ty(1.543) * rtb.ET.tz(1.543)
         * rtb.ET.Rx(jindex=0) * rtb.ET.Ry(jindex=1) * rtb.ET.Rz(jindex=2) * rtb.ET.tx(jindex=3) * rtb.ET.ty(jindex=4)
         * rtb.ET.tz(jindex=5))
    ans = (ch.trotx(1.543) @ ch.troty(1.543) @ ch.trotz(1.543) @ ch.transl(1.543, 0, 0) @ ch.transl(0, 1.543, 0)
           @ ch.transl(0, 0, 1.543) @ ch.trotx(q[0]) @ ch.troty(q[1]) @ ch.trotz(q[2]) @ ch.transl(q[3], 0, 0)
           @ ch.transl(0, q[4], 0) @ ch.transl(0, 0, q[5]))
    nt.assert_almost_equal(r.fkine(q).A, ans)
    # base / tool arguments (tests/test_ETS.py:224-234, numeric part)
    base, tool = ch.trotx(1.0), ch.transl(0, 0, 0.5)
    nt.assert_almost_equal(r.fkine(q, base=base).A, base @ ans)
    r2 = rtb.ETS([rtb.ET.Rx(jindex=0)])
    nt.assert_almost_equal(r2.fkine([0.7], tool=tool).A, ch.trotx(0.7) @ tool)


def test_fkine_traj():
    """tests/test_ETS.py:236-260 test_fkine_traj: a trajectory evaluates like its rows one by one"""
    robot = rtb.ERobot([rtb.Link(rtb.ET.Rx()), rtb.Link(rtb.ET.Ry()), rtb.Link(rtb.ET.Rz()), rtb.Link(rtb.ET.tx()),
                        rtb.Link(rtb.ET.ty()), rtb.Link(rtb.ET.tz())])
    ets = robot.ets()
    qt = np.arange(10 * ets.n).reshape(10, ets.n)
    T_traj = ets.eval(qt)
    for i in range(10):
        nt.assert_allclose(T_traj[i], ets.eval(qt[i]), rtol=1e-12, atol=1e-12)


def _unit_chains():
    return (rtb.ETS(rtb.ET.Rx()), rtb.ETS(rtb.ET.Ry()), rtb.ETS(rtb.ET.Rz()), rtb.ETS(rtb.ET.tx()), rtb.ETS(rtb.ET.ty()),
            rtb.ETS(rtb.ET.tz()))


@pytest.mark.parametrize("which", ["jacob0", "jacobe"])
def test_jacobians_of_unit_joints(which):
    """tests/test_ETS.py:590-626 test_jacob0 / test_jacobe"""
    rx, ry, rz, tx, ty, tz = _unit_chains()
    q = [0.0]
    eye = np.eye(6)
    for k, e in enumerate((tx, ty, tz, rx, ry, rz)):
        nt.assert_almost_equal(getattr(e, which)(q), eye[:, k:k + 1])
    r = tx + ty + tz + rx + ry + rz
    nt.assert_almost_equal(getattr(r, which)(np.zeros(6)), eye)


def _panda_by_hand():
    deg, mm = np.pi / 180, 1e-3
    l0 = rtb.ET.tz(0.333) * rtb.ET.Rz(jindex=0)
    l1 = rtb.ET.Rx(-90 * deg) * rtb.ET.Rz(jindex=1)
    l2 = rtb.ET.Rx(90 * deg) * rtb.ET.tz(0.316) * rtb.ET.Rz(jindex=2)
    l3 = rtb.ET.tx(0.0825) * rtb.ET.Rx(90 * deg) * rtb.ET.Rz(jindex=3)
    l4 = rtb.ET.tx(-0.0825) * rtb.ET.Rx(-90 * deg) * rtb.ET.tz(0.384) * rtb.ET.Rz(jindex=4)
    l5 = rtb.ET.Rx(90 * deg) * rtb.ET.Rz(jindex=5)
    l6 = rtb.ET.tx(0.088) * rtb.ET.Rx(90 * deg) * rtb.ET.tz(0.107) * rtb.ET.Rz(jindex=6)
    ee = rtb.ET.tz(103 * mm) * rtb.ET.Rz(-np.pi / 4)
    return l0 + l1 + l2 + l3 + l4 + l5 + l6 + ee


def test_jacobe_panda_is_rotated_jacob0():
    """tests/test_ETS.py:365-398 test_jacobe_panda: jacobe == tr2jac(T) @ jacob0 on the hand-built Panda"""
    r = _panda_by_hand()
    q1 = np.array([1.4, 0.2, 1.8, 0.7, 0.1, 3.1, 2.9])
    ans = tr2jac(r.eval(q1)) @ r.jacob0(q1)
    nt.assert_array_almost_equal(r.jacobe(q1), ans)
    for qq in (list(q1), q1[None, :], q1[:, None]):  # the argument forms of tests/test_ETS.py:296-299
        nt.assert_array_almost_equal(r.jacobe(qq), ans)
    nt.assert_array_almost_equal(r.jacob0(q1), rtb.models.ETS.Panda().ets().jacob0(q1))


# ------------------------------------------------------------------ tests/test_jacob.py
def test_jacob0_jacobe_against_numerical_derivative():
    """tests/test_jacob.py:27-39 test_jacob0 / test_jacobe (Puma560, q = [0.1 0.2 0.3 0.1 0.2 0.3])"""
    robot = rtb.models.DH.Puma560()
    q = np.array([0.1, 0.2, 0.3, 0.1, 0.2, 0.3])
    fk = lambda x: robot.ets().eval(x)  # noqa: E731
    J0 = numjac(fk, q)
    nt.assert_array_almost_equal(robot.jacob0(q), J0, decimal=5)
    TE = fk(q)
    Je = np.kron(np.eye(2), TE[:3, :3].T) @ J0
    nt.assert_array_almost_equal(robot.jacobe(q), Je, decimal=5)


def test_jacob0_flipped_joint_chain():
    """tests/test_jacob.py:101-145 test_jacob0_flipped0: a 7-joint Rz/Ry chain with one flipped joint"""
    def se3(x, z):
        T = np.eye(4); T[0, 3] = x; T[2, 3] = z
        return rtb.ET.SE3(T=T)
    robot = rtb.ETS([
        rtb.ET.Rz(jindex=0, qlim=np.array([-2.9668, 2.9668])), se3(-4.3624e-04, 3.6000e-01),
        rtb.ET.Ry(jindex=1, qlim=np.array([-2.0942, 2.0942])), rtb.ET.Rz(jindex=2, qlim=np.array([-2.9668, 2.9668])),
        se3(4.3624e-04, 4.2000e-01), rtb.ET.Ry(jindex=3, flip=True, qlim=np.array([-2.0942, 2.0942])),
        rtb.ET.Rz(jindex=4, qlim=np.array([-2.9668, 2.9668])), rtb.ET.tz(0.4),
        rtb.ET.Ry(jindex=5, qlim=np.array([-2.0942, 2.0942])), rtb.ET.Rz(jindex=6, qlim=np.array([-3.0541, 3.0541])),
        rtb.ET.tz(0.126)])
    q = np.array([0, -0.3, 0, -2.2, 0, 2, 0.79])
    nt.assert_array_almost_equal(robot.jacob0(q), numjac(lambda x: robot.eval(x), q), decimal=5)
    Je = np.kron(np.eye(2), robot.eval(q)[:3, :3].T) @ robot.jacob0(q)
    nt.assert_array_almost_equal(robot.jacobe(q), Je)


# ------------------------------------------------------------------ tests/test_DHRobot.py
def test_dhrobot_fkine_prismatic_mix():
    """tests/test_DHRobot.py:170-208 test_fkine / test_fkine_traj"""
    r0 = rtb.DHRobot([rtb.PrismaticDH(), rtb.RevoluteDH(), rtb.PrismaticDH(theta=2.0), rtb.RevoluteDH()])
    q = np.array([1, 2, 3, 4])
    T1 = np.array([[-0.14550003, -0.98935825, 0, 0], [0.98935825, -0.14550003, 0, 0], [0, 0, 1, 4], [0, 0, 0, 1]])
    nt.assert_array_almost_equal(r0.fkine(q).A, T1)
    TT = r0.fkine(np.tile(q, (4, 1)))
    for i in range(4):
        nt.assert_array_almost_equal(TT[i].A, T1)


def test_dhrobot_fkine_panda():
    """tests/test_DHRobot.py:438-451 test_fkine_panda (modified-DH Panda)"""
    panda = rtb.models.DH.Panda()
    T = np.array([[-0.8583, 0.1178, 0.4994, 0.1372], [0.1980, 0.9739, 0.1106, 0.3246], [-0.4734, 0.1938, -0.8593, 0.4436],
                  [0, 0, 0, 1]])
    nt.assert_array_almost_equal(panda.fkine([1, 2, 3, 4, 5, 6, 7]).A, T, decimal=4)


def test_dhrobot_jacobians_prismatic_mix():
    """tests/test_DHRobot.py:453-493 test_jacobe / test_jacob0"""
    r0 = rtb.DHRobot([rtb.PrismaticDH(theta=4), rtb.RevoluteDH(a=2), rtb.PrismaticDH(theta=2), rtb.RevoluteDH()])
    q = [1, 2, 3, 4]
    Je = np.array([[0, -0.5588, 0, 0], [0, 1.9203, 0, 0], [1.0, 0, 1.0, 0], [0, 0, 0, 0], [0, 0, 0, 0], [0, 1.0, 0, 1.0]])
    J0 = np.array([[0, 0.5588, 0, 0], [0, 1.9203, 0, 0], [1.0, 0, 1.0, 0], [0, 0, 0, 0], [0, 0, 0, 0], [0, 1.0, 0, 1.0]])
    nt.assert_array_almost_equal(r0.jacobe(q), Je, decimal=4)
    nt.assert_array_almost_equal(r0.jacob0(q), J0, decimal=4)


def test_dhrobot_jacobe_panda():
    """tests/test_DHRobot.py:495-514 test_jacobe_panda"""
    panda = rtb.models.DH.Panda()
    Je = np.array([[0.3058, 0.1315, -0.2364, -0.0323, 0.0018, 0.2095, 0], [0.0954, 0.0303, -0.0721, 0.1494, -0.0258, 0.0144, 0],
                   [-0.1469, 0.3385, 0.0506, 0.0847, -0.0000, -0.0880, 0], [-0.4734, 0.8292, -0.0732, 0.8991, -0.2788, -0.0685, 0],
                   [0.1938, 0.4271, 0.7224, 0.3461, -0.0191, 0.9976, 0], [-0.8593, -0.3605, 0.6876, -0.2679, -0.9602, 0.0000, 1.0]])
    nt.assert_array_almost_equal(panda.jacobe([1, 2, 3, 4, 5, 6, 7]), Je, decimal=4)


def test_dhrobot_rne_traj_and_delete():
    """tests/test_DHRobot.py:1065-1090 test_rne_traj / test_rne_delete"""
    puma = rtb.models.DH.Puma560()
    z, o = np.zeros(6), np.ones(6)
    tr0 = [-0.0000, 31.6399, 6.0351, 0.0000, 0.0283, 0]
    tr1 = [32.4952, 60.8670, 17.7436, 1.4545, 1.2991, 0.7138]
    t0 = puma.rne(np.c_[puma.qn, puma.qn].T, np.c_[z, o].T, np.c_[z, o].T)
    nt.assert_array_almost_equal(t0[0, :], tr0, decimal=4)
    nt.assert_array_almost_equal(t0[1, :], tr1, decimal=4)
    a = puma.rne(puma.qn, z, z)
    puma.delete_rne()
    b = puma.rne(puma.qn, z, z)
    nt.assert_array_almost_equal(a, tr0, decimal=4)
    nt.assert_array_almost_equal(b, tr0, decimal=4)


# ------------------------------------------------------------------ tests/test_IK.py
@pytest.mark.parametrize("method", ["chan", "wampler", "sugihara"])
def test_ik_lm_reaches_the_pose(method):
    """tests/test_IK.py:186-251 (IK_LM chan / wampler / sugihara on the Panda): solve, then check the pose"""
    panda = rtb.models.Panda()
    Tep = panda.ets().eval([0, -0.3, 0, -2.2, 0, 2, np.pi / 4])
    k = {"chan": 1.0, "wampler": 1e-4, "sugihara": 1e-4}[method]  # wampler / sugihara want a small gain
    solver = rtb.IK_LM(method=method, k=k, seed=0)
    sol = solver.solve(panda.ets(), Tep)
    assert sol.success
    nt.assert_array_almost_equal(panda.ets().eval(sol.q), Tep, decimal=3)
    q, ok, its, searches, E = panda.ik_LM(Tep, method=method, k=k)
    assert ok == 1 and E < 1e-6
    nt.assert_array_almost_equal(panda.ets().eval(q), Tep, decimal=3)


def test_ik_solver_classes_on_a_trajectory():
    """tests/test_IK.py:451-492, 632-708 (IK_NR / IK_GN and trajectory input): q is (N, n), success the conjunction"""
    panda = rtb.models.Panda()
    Q = np.linspace([0, -0.3, 0, -2.2, 0, 2, 0.7], [0.4, -0.1, 0.3, -1.9, 0.2, 2.2, 0.9], 16)
    Tep = panda.ets().eval(Q)
    for solver in (rtb.IK_LM(seed=1), rtb.IK_NR(seed=1, pinv=True), rtb.IK_GN(seed=1, pinv=True)):
        sol = solver.solve(panda.ets(), Tep)
        assert sol.q.shape == (16, 7)
        assert sol.success and sol.searches >= 16
        nt.assert_array_almost_equal(panda.ets().eval(sol.q), Tep, decimal=2)


def test_sub_chain_ets_start_end():
    """BaseRobot.ets(start, end) (BaseRobot.py:1554-1652) on the serial Panda: the chain splits into link ranges whose
    poses multiply back to the whole, joints keep their robot-wide jindex (q stays the full joint vector)."""
    panda = rtb.models.Panda()
    Q = np.random.default_rng(12).uniform(-2, 2, (64, 7))
    L = panda.links
    head = panda.ets(end=L[3])                 # base .. link3
    tail = panda.ets(start=L[4], end=L[-1])    # link4 .. end-effector
    assert [et.jindex for et in tail if et.isjoint] == [4, 5, 6]
    assert panda.ets(end=L[3].name).n == head.n == 4
    T_head, T_tail = head.eval(Q[:, :4]), tail.eval(Q)
    nt.assert_allclose(T_head @ T_tail, panda.ets().eval(Q), rtol=1e-10, atol=1e-12)
    # Jacobian of a sub-chain = numerical derivative of its own pose
    q = Q[0]
    Jt = tail.jacob0(q)
    full = numjac(lambda x: tail.eval(x), q)
    nt.assert_array_almost_equal(Jt, full[:, 4:], decimal=5)
    nt.assert_array_almost_equal(panda.jacob0(q, end=L[3]), numjac(lambda x: head.eval(x), q[:4]), decimal=5)
    with pytest.raises(ValueError):
        panda.ets(end="no_such_link")
    # a path towards the base is the inverse of the forward range (reference _find_ets, BaseRobot.py:1457-1467)
    back = panda.ets(start=L[6], end=L[2])
    fwd = panda.ets(start=L[3], end=L[6])
    nt.assert_allclose(fwd.eval(Q) @ back.eval(Q), np.broadcast_to(np.eye(4), (64, 4, 4)), rtol=0, atol=1e-12)
    e = rtb.ET.Rz(jindex=2) * rtb.ET.tx(1) * rtb.ET.Rx(jindex=3, flip=True) * rtb.ET.tx(1)  # docstring example of ETS.inv
    nt.assert_allclose(e.eval(Q[:, :4]) @ e.inv().eval(Q[:, :4]), np.broadcast_to(np.eye(4), (64, 4, 4)), rtol=0, atol=1e-12)


def test_trajectory_ctraj_and_mstraj_reference_unit_tests():
    """reference tests/test_trajectory.py:175-204 (test_ctraj) and 592-664 (test_mstraj), replayed through the mirror."""
    import json

    from oracle import chains as ch

    T0, T1 = ch.transl(1, 2, 3), ch.transl(-1, -2, -3)
    T = rtb.ctraj(T0, T1, 3)
    assert len(T) == 3
    np.testing.assert_array_almost_equal(T[0].A, T0)
    np.testing.assert_array_almost_equal(T[2].A, T1)
    np.testing.assert_array_almost_equal(T[1].A, np.eye(4))
    T = rtb.ctraj(T0, T1, [1, 0, 0.5])
    assert len(T) == 3
    np.testing.assert_array_almost_equal(T[0].A, T1)
    np.testing.assert_array_almost_equal(T[1].A, T0)
    np.testing.assert_array_almost_equal(T[2].A, np.eye(4))
    T0, T1 = ch.trotx(-np.pi / 2), ch.trotx(np.pi / 2)
    T = rtb.ctraj(T0, T1, 3)
    np.testing.assert_array_almost_equal(T[0].A, T0)
    np.testing.assert_array_almost_equal(T[2].A, T1)
    np.testing.assert_array_almost_equal(T[1].A, np.eye(4))
    with pytest.raises(TypeError):
        rtb.ctraj(T0, T1, "hello")

    K = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))["mstraj"]
    via = np.array(K["via"])
    out = rtb.mstraj(via, dt=1, tacc=1, qdmax=[2, 1], q0=[4, 1])
    np.testing.assert_array_almost_equal(out.q, np.array(K["qdmax_case"]["q"]), decimal=4)
    out = rtb.mstraj(via, dt=1, tacc=1, tsegment=[2, 1, 3, 4], q0=[4, 1])
    np.testing.assert_array_almost_equal(out.q, np.array(K["tsegment_case"]["q"]), decimal=4)
    out = rtb.mstraj(via, dt=1, tacc=1, tsegment=[1, 2, 3, 4], q0=via[0, :])
    assert out.t.shape[0] == out.q.shape[0]
    assert isinstance(out.info, list) and len(out.info) == via.shape[0] + 1
    rtb.mstraj(via, dt=1, tacc=1, qdmax=[2, 1])
    rtb.mstraj(via, dt=1, tacc=1, qdmax=2)
    for kw in (dict(qdmax=[2, 1], q0=[1, 2, 3]), dict(qdmax=[2, 1], tsegment=[1, 2, 3, 4]), dict(), dict(tsegment=[3, 4]),
               dict(qdmax=[2, 1, 3]), dict(qdmax=[2, 1], qd0=[1, 2, 3], q0=[1, 2]), dict(qdmax=[2, 1], qdf=[1, 2, 3], q0=[1, 2])):
        with pytest.raises(ValueError):
            rtb.mstraj(via, dt=1, tacc=1, **kw)
    with pytest.raises(ValueError):
        rtb.mstraj(via, dt=1, tacc=[1, 2, 3, 4, 5], qdmax=[2, 1])


def test_robot_jtraj_between_poses():
    """Robot.jtraj (Robot.py:917-961): IK at both ends, quintic in between -- the end points of the trajectory reach the
    two poses, the interior is the quintic of tools/trajectory.jtraj (checked elsewhere), for the ETS and the DH class."""
    from oracle import oracle as orc

    for robot in (rtb.models.Panda(), rtb.models.Puma560()):
        C = orc.Chain(robot.ets().describe())
        qa, qb = (np.r_[0.1, -0.4, 0.2, -1.9, 0.1, 1.6, 0.5][:robot.n], np.r_[0.6, 0.1, -0.3, -1.4, 0.4, 1.2, -0.2][:robot.n])
        T1, T2 = C.fkine(qa[None])[0], C.fkine(qb[None])[0]
        tg = robot.jtraj(T1, T2, 50, seed=3)
        assert tg.q.shape == (50, robot.n)
        np.testing.assert_allclose(C.fkine(tg.q[:1])[0], T1, atol=5e-3)
        np.testing.assert_allclose(C.fkine(tg.q[-1:])[0], T2, atol=5e-3)
        np.testing.assert_allclose(tg.qd[0], 0, atol=1e-12)
        np.testing.assert_allclose(tg.qd[-1], 0, atol=1e-9)
        dv = robot.jtraj(T1, T2, 50, device=True, seed=3)
        assert dv.q.is_cuda
        np.testing.assert_allclose(dv.q.cpu().numpy(), tg.q, atol=1e-9)
