"""GPU tests of Robot.rne for rigid-body trees (reference Robot.py:1704-1903): the reference's own known answers
(tests/test_ERobot.py:100-200, Spong's two-link arm), the numpy restatement of its spatial-vector recursion on random
branched trees, a cross-check of the two dynamics formulations (an ETS robot built from the Puma560 DH table against
the DH recursion the compiled frne defines), and a robot ingested from URDF."""
import os
from math import cos, pi, sin

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import b2kin as rtb  # noqa: E402
from oracle import chains as ch  # noqa: E402
from oracle import oracle as orc  # noqa: E402

ET, ETS, Link, Robot = rtb.ET, rtb.ETS, rtb.Link, rtb.Robot
URDF_DIR = os.path.join(os.path.dirname(__file__), "golden", "urdf")


def dev(a, dt=np.float64):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).cuda()


def test_reference_known_answers_spong_two_link():
    """tests/test_ERobot.py:100-154 (test_invdyn) and 156-200 (test_invdyn_static), literal values."""
    for static_middle in (False, True):
        l1 = Link(ets=ETS(ET.Ry()), m=1, r=[0.5, 0, 0], name="l1")
        if static_middle:
            l2 = Link(ets=ETS(), m=0, r=[0, 0, 0], parent=l1, name="l2")
            l3 = Link(ets=ETS(ET.tx(1)) * ET.Ry(), m=1, r=[0.5, 0, 0], parent=l2, name="l3")
            robot = Robot([l1, l2, l3], name="simple 3 link")
        else:
            l2 = Link(ets=ETS(ET.tx(1)) * ET.Ry(), m=1, r=[0.5, 0, 0], parent=l1, name="l2")
            robot = Robot([l1, l2], name="simple 2 link")
        z = np.zeros(robot.n)
        np.testing.assert_array_almost_equal(robot.rne(z, z, z) / 9.81, np.r_[-2, -0.5])
        np.testing.assert_array_almost_equal(robot.rne(np.array([0.0, -pi / 2.0]), z, z) / 9.81, np.r_[-1.5, 0])
        np.testing.assert_array_almost_equal(robot.rne(np.array([-pi / 2, pi / 2]), z, z) / 9.81, np.r_[-0.5, -0.5])
        np.testing.assert_array_almost_equal(robot.rne(np.array([-pi / 2, 0]), z, z) / 9.81, np.r_[0, 0])
        robot.gravity = [0, 0, 0]
        q = np.array([0, -pi / 2])
        h = -0.5 * sin(q[1])
        np.testing.assert_array_almost_equal(robot.rne(q, np.array([0, 0]), z), np.r_[0, 0] * h)
        np.testing.assert_array_almost_equal(robot.rne(q, np.array([1, 0]), z), np.r_[0, -1] * h)
        np.testing.assert_array_almost_equal(robot.rne(q, np.array([0, 1]), z), np.r_[1, 0] * h)
        np.testing.assert_array_almost_equal(robot.rne(q, np.array([1, 1]), z), np.r_[3, -1] * h)
        d11, d12, d22 = 1.5 + cos(q[1]), 0.25 + 0.5 * cos(q[1]), 0.25
        np.testing.assert_array_almost_equal(robot.rne(q, z, np.array([0, 0])), np.r_[0, 0])
        np.testing.assert_array_almost_equal(robot.rne(q, z, np.array([1, 0])), np.r_[d11, d12])
        np.testing.assert_array_almost_equal(robot.rne(q, z, np.array([0, 1])), np.r_[d12, d22])
        np.testing.assert_array_almost_equal(robot.rne(q, z, np.array([1, 1])), np.r_[d11 + d12, d12 + d22])
        assert robot.rne_kernel_info().startswith("k_rne_spec<double,tree n=2>")


def random_robot(rng, n, branched):
    links = []
    for j in range(n):
        parent = None if j == 0 else links[int(rng.integers(0, len(links))) if branched else -1]
        consts = ETS()
        for _ in range(int(rng.integers(0, 3))):
            kind = int(rng.integers(0, 6))
            val = float(rng.choice([0.0, np.pi / 2, -np.pi / 2, rng.uniform(-1, 1)])) if kind < 3 else float(rng.uniform(-0.4, 0.4))
            consts = consts * getattr(ET, ("Rx", "Ry", "Rz", "tx", "ty", "tz")[kind])(val)
        joint = getattr(ET, ("Rx", "Ry", "Rz", "tx", "ty", "tz")[int(rng.integers(0, 6))])(flip=bool(rng.integers(0, 2)))
        links.append(Link(consts * joint, name=f"j{j}", parent=parent, m=float(rng.uniform(0.2, 3)),
                          r=(rng.uniform(-0.2, 0.2, 3) * rng.integers(0, 2, 3)).tolist()))
        if rng.random() < 0.3:  # a massless static link in between
            links.append(Link(ETS(ET.tx(float(rng.uniform(-0.2, 0.2)))), name=f"s{j}", parent=links[-1]))
    return Robot(links)


@pytest.mark.parametrize("n,branched", [(1, False), (4, False), (7, True), (12, True)])
def test_random_trees_against_the_spatial_vector_oracle(n, branched):
    rng = np.random.default_rng(300 + n)
    rob = random_robot(rng, n, branched)
    assert rob.n == n
    tree = rob.tree_description()
    for N in (1, 33, 1000):
        q, qd, qdd = rng.uniform(-3, 3, (N, n)), rng.normal(size=(N, n)), rng.normal(size=(N, n))
        for grav in (None, [0.5, -1.0, 3.0], [0, 0, 0]):
            want = orc.tree_rne(tree, q, qd, qdd, rob.gravity if grav is None else grav)
            got = rob.rne(dev(q), dev(qd), dev(qdd), gravity=grav)
            np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-9, atol=1e-9)
    q, qd, qdd = rng.uniform(-2, 2, (64, n)), rng.normal(size=(64, n)), rng.normal(size=(64, n))
    want = orc.tree_rne(tree, q, qd, qdd, rob.gravity)
    np.testing.assert_allclose(rob.rne(q, qd, qdd), want, rtol=1e-9, atol=1e-9)  # numpy in -> numpy out
    got32 = rob.rne(q.astype(np.float32), qd.astype(np.float32), qdd.astype(np.float32))
    assert got32.dtype == np.float32
    np.testing.assert_allclose(got32, want, rtol=2e-3, atol=2e-3 * max(1.0, np.abs(want).max()))
    one = rob.rne(q[0], qd[0], qdd[0])
    assert one.shape == (n,)
    np.testing.assert_allclose(one, want[0], rtol=1e-9, atol=1e-9)
    with pytest.raises(ValueError):
        rob.rne(q[:, :-1] if n > 1 else np.zeros((3, 2)), qd, qdd)


def test_ets_robot_built_from_the_puma_dh_table_reproduces_the_dh_recursion():
    """Two formulations, one robot: the DH recursion of ne.c (what the compiled frne computes; its restatement is pinned
    to it) and Featherstone's recursion on an ETS robot assembled from the same DH table.  Robot.rne uses mass and centre
    of mass only (Robot.py:1775-1783), so the comparison robot has no rotational / motor inertia and no friction."""
    dh = ch.puma560_links()
    for l in dh:
        l.update(I=[0, 0, 0, 0, 0, 0], Jm=0.0, B=0.0, Tc=[0.0, 0.0], G=1.0)
    L = ch.pack_rne(dh)
    links, prev = [], ETS()
    for j, l in enumerate(dh):
        post = ETS()
        for kind, v in (("tz", l["d"]), ("tx", l["a"]), ("Rx", l["alpha"])):
            if v != 0:
                post = post * getattr(ET, kind)(v)
        Tpost = np.eye(4)
        for et in post:
            Tpost = Tpost @ et.A()
        r_in_joint_frame = Tpost[:3, :3] @ np.asarray(l["r"], dtype=float) + Tpost[:3, 3]
        links.append(Link(prev * ET.Rz(), name=f"l{j}", parent=links[-1] if links else None, m=l["m"], r=r_in_joint_frame))
        prev = post
    links.append(Link(prev, name="tool", parent=links[-1]))
    rob = Robot(links)
    rng = np.random.default_rng(8)
    N = 2000
    q, qd, qdd = rng.uniform(-3, 3, (N, 6)), rng.normal(size=(N, 6)), rng.normal(size=(N, 6))
    want = orc.rne(6, 0, L, np.array([0, 0, 9.81]), q, qd, qdd)
    got = rob.rne(dev(q), dev(qd), dev(qdd)).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-9)
    # and the kinematics of that robot are the DH robot's
    Cd = orc.Chain(ch.dh_to_ets(ch.puma560_links()))
    np.testing.assert_allclose(rob.eval(dev(q[:100])).cpu().numpy(), Cd.fkine(q[:100]), rtol=1e-10, atol=1e-12)


def test_urdf_robot_dynamics_and_kinematics_on_the_device():
    rob = Robot.URDF(os.path.join(URDF_DIR, "two_arm.urdf"))
    tree = rob.tree_description()
    rng = np.random.default_rng(9)
    N = 500
    q, qd, qdd = rng.uniform(-1, 1, (N, 5)), rng.normal(size=(N, 5)), rng.normal(size=(N, 5))
    np.testing.assert_allclose(rob.rne(dev(q), dev(qd), dev(qdd)).cpu().numpy(), orc.tree_rne(tree, q, qd, qdd, rob.gravity),
                               rtol=1e-9, atol=1e-9)
    # the torso's own mass loads no joint; gravity load of the left shoulder (axis z, vertical) is zero
    tg = rob.rne(dev(q), dev(0 * q), dev(0 * q)).cpu().numpy()
    assert np.abs(tg[:, 0]).max() < 1e-12
    # kinematics across the branches on the device: left flange -> right hand
    e = rob.ets(start="l_flange", end="r_skew")
    C = orc.Chain(e.describe())
    np.testing.assert_allclose(e.eval(dev(q)).cpu().numpy(), C.fkine(q), rtol=1e-10, atol=1e-12)
    J = rob.jacob0(dev(q), end="r_skew")
    np.testing.assert_allclose(J.cpu().numpy(), orc.Chain(rob.ets(end="r_skew").describe()).jacob0(q), rtol=1e-10, atol=1e-12)


def _oracle_rne_fn(tree):
    return lambda a, b, c, g: orc.tree_rne(tree, np.atleast_2d(a), np.atleast_2d(b), np.atleast_2d(c), g)


@pytest.mark.parametrize("case", ["urdf", "random7", "puma_ets"])
def test_tree_robot_dynamics_operations_against_the_reference_loops(case):
    """Robot.inertia / gravload / itorque / coriolis / accel (DynamicsMixin on a tree robot: Python loops of self.rne in
    the reference, Dynamics.py:424-503, 700-915, 1418-1459) -- one generated kernel each, against the loops restated in
    oracle.dyn_* over the spatial-vector oracle.  fp64 to 1e-9, fp32 to its bar, host and device inputs, ragged batch."""
    rng = np.random.default_rng({"urdf": 21, "random7": 22, "puma_ets": 23}[case])
    if case == "urdf":
        rob = Robot.URDF(os.path.join(URDF_DIR, "two_arm.urdf"))
    elif case == "random7":
        rob = random_robot(rng, 7, True)
    else:
        dh = ch.puma560_links()
        links, prev = [], ETS()
        for j, l in enumerate(dh):
            post = ETS()
            for kind, v in (("tz", l["d"]), ("tx", l["a"]), ("Rx", l["alpha"])):
                if v != 0:
                    post = post * getattr(ET, kind)(v)
            Tpost = np.eye(4)
            for et in post:
                Tpost = Tpost @ et.A()
            r_in = Tpost[:3, :3] @ np.asarray(l["r"], dtype=float) + Tpost[:3, 3]
            links.append(Link(prev * ET.Rz(), name=f"l{j}", parent=links[-1] if links else None, m=l["m"] + 0.5, r=r_in + 0.05))
            prev = post
        rob = Robot(links)
    n = rob.n
    tree = rob.tree_description()
    f = _oracle_rne_fn(tree)
    N = 77  # two full tiles and a ragged tail
    q, qd, x = rng.uniform(-3, 3, (N, n)), rng.normal(size=(N, n)), rng.normal(size=(N, n))
    grav = [0.4, -0.3, -9.81]
    tol = dict(rtol=1e-9, atol=1e-9)
    M = orc.dyn_inertia(f, n, q)
    np.testing.assert_allclose(rob.inertia(dev(q)).cpu().numpy(), M, **tol)
    np.testing.assert_allclose(rob.gravload(dev(q)).cpu().numpy(), orc.dyn_gravload(f, n, q, rob.gravity), **tol)
    np.testing.assert_allclose(rob.gravload(dev(q), gravity=grav).cpu().numpy(), orc.dyn_gravload(f, n, q, np.asarray(grav)), **tol)
    np.testing.assert_allclose(rob.itorque(dev(q), dev(x)).cpu().numpy(), orc.dyn_itorque(f, n, q, x), **tol)
    np.testing.assert_allclose(rob.coriolis(dev(q), dev(qd)).cpu().numpy(), orc.dyn_coriolis(f, n, q, qd), **tol)
    # consistency the reference's formulas imply: M qdd = itorque, tau = M qdd + C qd + g
    np.testing.assert_allclose(np.einsum("kij,ki->kj", M, x), orc.dyn_itorque(f, n, q, x), rtol=1e-9, atol=1e-9)
    if np.linalg.cond(M).max() < 1e6:
        want = orc.dyn_accel(f, n, q, qd, x, np.asarray(grav))
        np.testing.assert_allclose(rob.accel(dev(q), dev(qd), dev(x), gravity=grav).cpu().numpy(), want, rtol=1e-6, atol=1e-7)
        got32 = rob.accel(q.astype(np.float32), qd.astype(np.float32), x.astype(np.float32), gravity=grav)
        assert got32.dtype == np.float32
        np.testing.assert_allclose(got32, want, rtol=5e-2, atol=5e-3 * max(1.0, np.abs(want).max()))
    else:
        assert case != "puma_ets", "the Puma-derived tree must have a regular inertia matrix"
    # numpy in -> numpy out, single state, fp32
    one = rob.inertia(q[0])
    assert isinstance(one, np.ndarray) and one.shape == (n, n)
    np.testing.assert_allclose(one, M[0], **tol)
    M32 = rob.inertia(q.astype(np.float32))
    assert M32.dtype == np.float32
    np.testing.assert_allclose(M32, M, rtol=2e-3, atol=2e-3 * max(1.0, np.abs(M).max()))
    assert rob.rne_kernel_info(op="coriolis").startswith(f"k_rne_spec<double,tree n={n},mode=4>")
    with pytest.raises(ValueError):
        rob.coriolis(q, qd[:-1])


def test_tree_robot_fdyn_follows_scipy_rk45_on_the_oracle():
    """Robot.fdyn (DynamicsMixin.fdyn through BaseRobot): the device-resident Dormand-Prince integrator around the tree's
    generated accel recursion takes the steps scipy's RK45 takes on the oracle's accel (the reference's procedure)."""
    rob = Robot.URDF(os.path.join(URDF_DIR, "two_arm.urdf"))
    n = rob.n
    tree = rob.tree_description()
    f = _oracle_rne_fn(tree)
    rng = np.random.default_rng(31)
    q0 = rng.uniform(-0.5, 0.5, n)
    M0 = orc.dyn_inertia(f, n, q0)
    if np.linalg.cond(M0).max() > 1e6:  # point-mass links on their own axes: not integrable, in the reference either
        rob = random_robot(np.random.default_rng(4), 5, True)
        n, tree = rob.n, rob.tree_description()
        f = _oracle_rne_fn(tree)
        q0 = rng.uniform(-0.5, 0.5, n)
        assert np.linalg.cond(orc.dyn_inertia(f, n, q0)).max() < 1e6
    acc = lambda q, qd, tau: orc.dyn_accel(f, n, q, qd, tau, rob.gravity)[0]  # noqa: E731
    sa = dict(rtol=1e-6, atol=1e-9)
    tg = rob.fdyn(0.25, q0, solver_args=sa)
    t, q, qd = orc.fdyn(acc, n, 0.25, q0, solver_args=sa)
    assert tg.t.shape == t.shape, (tg.t.shape, t.shape)
    np.testing.assert_allclose(tg.t, t, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(tg.q, q, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(tg.qd, qd, rtol=1e-5, atol=1e-7)
    tau = rng.normal(size=n)
    tg = rob.fdyn(0.2, q0, Q=tau, solver_args=sa, dt=0.02)
    t, q, qd = orc.fdyn(acc, n, 0.2, q0, torque_fn=lambda t, q, qd: tau, solver_args=sa, dt=0.02)
    np.testing.assert_allclose(tg.q, q, rtol=1e-6, atol=1e-8)
    Q0 = q0 + rng.uniform(-0.2, 0.2, (64, n))
    ens = rob.fdyn(0.2, dev(Q0), solver_args=sa, dt=0.02)
    assert ens.q.is_cuda and ens.q.shape == (64, 10, n)
    for i in (0, 63):
        t, q, qd = orc.fdyn(acc, n, 0.2, Q0[i], solver_args=sa, dt=0.02)
        np.testing.assert_allclose(ens.q[i].cpu().numpy(), q, rtol=1e-6, atol=1e-8)
