"""Pin the CPU oracle (oracle/oracle_kin.c) before anything trusts it.

(a) against the literal golden vectors of the reference's own tests
    (tests/golden/reference_kats.json, each with its reference file:line);
(b) against fixtures produced by the compiled reference itself
    (tests/golden/*.npz, generator tests/golden/make_golden.py);
(c) where oracle/_ref is present (build container / shipped to the GPU box), live against
    the compiled reference on fresh random inputs.
"""
import json
import os

import numpy as np
import pytest

from oracle import chains as ch
from oracle import oracle as orc
from oracle import ref_driver as ref

G = os.path.join(os.path.dirname(__file__), "golden")
KAT = json.load(open(os.path.join(G, "reference_kats.json")))


def load_desc(z, prefix=""):
    return {k: z[prefix + k] for k in ("isjoint", "axis", "flip", "jindex", "T", "qlim")}


def opt(a):
    return None if a.size == 0 else a


# ------------------------------------------------------------------ (a) literal KATs
def test_kat_panda_fkine():
    k = KAT["panda_fkine"]
    T = orc.Chain(ch.panda_ets()).fkine(np.array(k["q"]))[0]
    np.testing.assert_array_almost_equal(T, np.array(k["T"]), decimal=k["decimal"])


def test_kat_panda_jacob0():
    k = KAT["panda_jacob0"]
    J = orc.Chain(ch.panda_ets()).jacob0(np.array(k["q"]))[0]
    np.testing.assert_array_almost_equal(J, np.array(k["J"]), decimal=k["decimal"])


def test_kat_puma_rne():
    k = KAT["puma560_rne"]
    L = ch.pack_rne(ch.puma560_links())
    for c in k["cases"]:
        g = np.array(c.get("gravity", [0, 0, -9.81]), dtype=float)
        tau = orc.rne(6, 0, L, -g, ch.PUMA_QN, np.full(6, float(c["qd"])), np.full(6, float(c["qdd"])),
                      fext=c.get("fext"))[0]
        np.testing.assert_array_almost_equal(tau, np.array(c["tau"], dtype=float), decimal=k["decimal"])


def test_kat_panda_ik_converges():
    C = orc.Chain(ch.panda_ets())
    qr = np.array(KAT["panda_ik"]["qr"])
    Tep = C.fkine(qr)
    for method, kk in (("chan", 1.0), ("sugihara", 0.1), ("wampler", 0.01)):
        for sem in (0, 1):
            q, s, it, sr, E = C.ik_lm(Tep, q0=None, method=method, k=kk, seed=0, semantics=sem)
            assert s[0] == 1 and E[0] < 1e-5
            assert np.abs(C.fkine(q)[0] - Tep[0]).max() < 5e-3


def test_jacobe_identity():
    """jacobe == tr2jac(T^T) @ jacob0 (reference tests/test_ETS.py:365-398)."""
    C = orc.Chain(ch.panda_ets())
    q = np.array(KAT["panda_jacob0"]["q"])
    T = C.fkine(q)[0]
    R = T[:3, :3]
    blk = np.zeros((6, 6)); blk[:3, :3] = R.T; blk[3:, 3:] = R.T
    np.testing.assert_allclose(C.jacobe(q)[0], blk @ C.jacob0(q)[0], atol=1e-14)


def test_jacob0_numerical_derivative():
    """reference tests/test_jacob.py:27-39 style: J0 linear rows vs finite differences of FK."""
    C = orc.Chain(ch.dh_to_ets(ch.puma560_links()))
    rng = np.random.default_rng(5)
    q = rng.uniform(-2, 2, 6)
    J = C.jacob0(q)[0]
    h = 1e-7
    for j in range(6):
        dq = np.zeros(6); dq[j] = h
        dp = (C.fkine(q + dq)[0][:3, 3] - C.fkine(q - dq)[0][:3, 3]) / (2 * h)
        np.testing.assert_allclose(J[:3, j], dp, atol=1e-6)


def test_dh_ets_equals_dh_link_products():
    """DH->ETS expansion (DHLink.py:173-225) against the closed-form link matrix (DHLink.py:633-673)."""
    for links, mdh, tool in ((ch.ur10_links(), False, None), (ch.puma560_links(), False, None),
                             (ch.panda_mdh_links(), True, ch.panda_mdh_tool())):
        C = orc.Chain(ch.dh_to_ets(links, mdh=mdh, tool=tool))
        rng = np.random.default_rng(9)
        for _ in range(5):
            q = rng.uniform(-3, 3, len(links))
            T = np.eye(4)
            for L, qj in zip(links, q):
                T = T @ ch.dh_A(L, qj, mdh)
            if tool is not None:
                T = T @ tool
            np.testing.assert_allclose(C.fkine(q)[0], T, atol=1e-13)


# ------------------------------------------------------------------ (b) fixtures from the compiled reference
@pytest.mark.parametrize("name", ["panda_fkj.npz", "ur10_fkj.npz"])
def test_fixture_fkj(name):
    z = np.load(os.path.join(G, name))
    C = orc.Chain(load_desc(z))
    np.testing.assert_allclose(C.fkine(z["Q"]), z["Tfk"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(C.jacob0(z["Q"]), z["J0"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(C.jacobe(z["Q"]), z["Je"], rtol=0, atol=1e-13)


def test_fixture_random_chains():
    z = np.load(os.path.join(G, "random_fkj.npz"))
    for c in range(int(z["nchains"])):
        p = f"c{c}_"
        C = orc.Chain(load_desc(z, p))
        base, tool = opt(z[p + "base"]), opt(z[p + "tool"])
        np.testing.assert_allclose(C.fkine(z[p + "Q"], base, tool), z[p + "Tfk"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(C.jacob0(z[p + "Q"], tool), z[p + "J0"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(C.jacobe(z[p + "Q"], tool), z[p + "Je"], rtol=0, atol=1e-12)


def test_fixture_rne():
    z = np.load(os.path.join(G, "puma_rne.npz"))
    g = z["gravity"]
    a = (6, 0, z["L"])
    np.testing.assert_allclose(orc.rne(*a, -g, z["q"], z["qd"], z["qdd"]), z["tau"], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(orc.rne(*a, -g, z["q"], z["qd"], z["qdd"], z["fext"]), z["tau_fext"], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(orc.rne(*a, np.zeros(3), z["q"], z["qd"], z["qdd"]), z["tau_zerog"], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(orc.rne(*a, -z["g2"], z["q"], z["qd"], z["qdd"], z["fext"]), z["tau_g2"], rtol=1e-13, atol=1e-13)
    z = np.load(os.path.join(G, "panda_mdh_rne.npz"))
    np.testing.assert_allclose(orc.rne(7, 1, z["L"], -z["gravity"], z["q"], z["qd"], z["qdd"], z["fext"]),
                               z["tau_fext"], rtol=1e-13, atol=1e-13)
    z = np.load(os.path.join(G, "random_rne.npz"))
    for c in range(int(z["ncases"])):
        p = f"c{c}_"
        n = z[p + "q"].shape[1]
        tau = orc.rne(n, int(z[p + "mdh"]), z[p + "L"], -z[p + "gravity"], z[p + "q"], z[p + "qd"], z[p + "qdd"], z[p + "fext"])
        np.testing.assert_allclose(tau, z[p + "tau"], rtol=1e-12, atol=1e-11)


def test_dynamics_fanouts_kat_and_fixture():
    """Dynamics fan-outs (reference Dynamics.py loops) on top of the oracle's rne: the reference's own
    KATs (tests/test_DHRobot.py:1092-1200) and the fixture generated through the compiled frne."""
    k = KAT["puma560_dynamics"]
    L = ch.pack_rne(ch.puma560_links())
    g = np.array([0, 0, -9.81])
    f = lambda q, qd, qdd, grav: orc.rne(6, 0, L, -np.asarray(grav, dtype=float), q, qd, qdd)  # noqa: E731
    Lnf = orc.nofriction_L(L)
    fnf = lambda q, qd, qdd, grav: orc.rne(6, 0, Lnf, -np.asarray(grav, dtype=float), q, qd, qdd)  # noqa: E731
    qn = ch.PUMA_QN
    dec = k["decimal"]
    np.testing.assert_array_almost_equal(orc.dyn_inertia(f, 6, qn)[0], np.array(k["inertia"]), decimal=dec)
    np.testing.assert_array_almost_equal(orc.dyn_gravload(f, 6, qn, g)[0], np.array(k["gravload"], dtype=float), decimal=dec)
    np.testing.assert_array_almost_equal(orc.dyn_itorque(f, 6, qn, k["itorque"]["qdd"])[0], k["itorque"]["taui"], decimal=dec)
    np.testing.assert_array_almost_equal(orc.dyn_coriolis(fnf, 6, qn, k["coriolis"]["qd"])[0], np.array(k["coriolis"]["C"], dtype=float), decimal=dec)
    np.testing.assert_array_almost_equal(orc.dyn_accel(f, 6, qn, k["accel"]["qd"], k["accel"]["torque"], g)[0], k["accel"]["qdd"], decimal=dec)
    z = np.load(os.path.join(G, "puma_dynamics.npz"))
    np.testing.assert_allclose(orc.dyn_inertia(f, 6, z["q"]), z["inertia"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(orc.dyn_coriolis(fnf, 6, z["q"], z["qd"]), z["coriolis"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(orc.dyn_accel(f, 6, z["q"], z["qd"], z["torque"], g), z["accel"], rtol=1e-10, atol=1e-10)


def test_hessian_and_manipulability_restatements():
    """oracle.hessian (methods.cpp:16-32) against the compiled fknm.ETS_hessian0/e fixture and the literal
    golden of the reference's own test (tests/test_ETS.py:718-1116, transcribed by make_golden.py)."""
    z = np.load(os.path.join(G, "panda_hessian.npz"))
    for k in range(z["Q"].shape[0]):
        np.testing.assert_allclose(orc.hessian(z["J0"][k]), z["H0"][k], rtol=0, atol=1e-14)
        np.testing.assert_allclose(orc.hessian(z["Je"][k]), z["He"][k], rtol=0, atol=1e-14)
    assert z["kat_hessian0_q1"].shape == (7, 6, 7)
    np.testing.assert_array_almost_equal(orc.hessian(z["J0"][0]), z["kat_hessian0_q1"], decimal=6)
    J = z["J0"][3]
    assert abs(orc.yoshikawa(J) - np.sqrt(abs(np.linalg.det(J @ J.T)))) < 1e-15
    assert abs(orc.yoshikawa(J[:, :6]) - abs(np.linalg.det(J[:, :6]))) < 1e-15
    # manipulability Jacobian (ETS.py:1672-1685) on the reference's J and H: the reference test's literal golden
    kat = KAT["panda_jacobm"]
    np.testing.assert_allclose(z["Q"][0], kat["q"])
    np.testing.assert_array_almost_equal(orc.jacobm(z["J0"][0], z["H0"][0]).ravel(), kat["Jm"], decimal=kat["decimal"])
    # Jacobian time derivative (Robot.py:1099) against a central difference of the reference's own jacob0
    C = orc.Chain(load_desc(z))
    q, qd, h = z["Q"][1], np.array([0.1, -0.2, 0.3, -0.4, 0.5, -0.6, 0.7]), 1e-6
    num = (C.jacob0((q + h * qd)[None])[0] - C.jacob0((q - h * qd)[None])[0]) / (2 * h)
    np.testing.assert_allclose(orc.jacob_dot(z["H0"][1], qd), num, atol=1e-8)


def test_fixture_angle_axis():
    z = np.load(os.path.join(G, "angle_axis.npz"))
    for Te, Tep, e in zip(z["Te"], z["Tep"], z["e"]):
        np.testing.assert_allclose(orc.angle_axis(Te, Tep), e, rtol=0, atol=1e-15)


def test_fixture_ik():
    z = np.load(os.path.join(G, "panda_ik.npz"))
    C = orc.Chain(load_desc(z))
    for tag, method, k in (("chan1", "chan", 1.0), ("chan01", "chan", 0.1), ("sugi", "sugihara", 1e-4)):
        q, s, it, sr, E = C.ik_lm(z["Tep"], q0=z["q0"], slimit=1, joint_limits=False, k=k, method=method)
        assert (s == z[tag + "_success"]).all()
        assert (it == z[tag + "_it"]).all()
        assert (sr == z[tag + "_search"]).all()
        ok = s == 1
        np.testing.assert_allclose(q[ok], z[tag + "_q"][ok], atol=1e-8)
        np.testing.assert_allclose(E[ok], z[tag + "_E"][ok], atol=1e-12)
    # joint-limit rejection after the fmod wrap (ik.cpp:50-52)
    q, s, it, sr, E = C.ik_lm(z["Tep"], q0=z["q0"], slimit=1, joint_limits=True, k=1.0, method="chan")
    assert (s == z["jl_success"]).all() and (it == z["jl_it"]).all() and (sr == z["jl_search"]).all()
    # masked
    q, s, it, sr, E = C.ik_lm(z["Tep"], q0=z["q0"], slimit=1, joint_limits=False, mask=z["mask"], k=1.0)
    assert (s == z["mask_success"]).mean() > 0.97  # the masked problem is rank deficient: LU vs inverse may split ties
    # restart statistics (reference RNG is unseeded libc rand: compare outcome rates only)
    q, s, it, sr, E = C.ik_lm(z["Tep"], q0=None, joint_limits=True, k=1.0, seed=11)
    assert s.mean() == z["rs_success"].mean() == 1.0
    assert abs(sr.mean() - z["rs_search"].mean()) < 0.5


IKNR_CASES = [("nr_d", "nr", 0.1, False), ("nr_d_mask", "nr", 0.1, True), ("nr", "nr", 0.0, False),
              ("gn", "gn", 0.0, False), ("gn_mask", "gn", 0.0, True), ("nr_inv", "nr", 0.0, False)]


def test_fixture_ik_nr_gn():
    """Newton-Raphson / Gauss-Newton restatements vs fknm.IK_NR_c / IK_GN_c.  The oracle takes the
    pseudo-inverse through a one-sided Jacobi SVD, the reference through Eigen's JacobiSVD / BDCSVD:
    damped runs reproduce the reference row for row; undamped Newton steps are chaotic near singular
    configurations, so those are held to outcome statistics and to the well-started half of the batch."""
    z = np.load(os.path.join(G, "ik_nr_gn.npz"))
    for name in ("panda", "ur10", "ur3"):
        C = orc.Chain(load_desc(z, name + "_"))
        Tep, q0 = z[name + "_Tep"], z[name + "_q0"]
        half = len(Tep) // 2
        for tag, meth, damp, masked in IKNR_CASES:
            if f"{name}_{tag}_q" not in z:
                continue
            q, s, it, sr, E = C.ik_lm(Tep, q0, 30, 1, 1e-6, False, z["mask"] if masked else None, damp, meth)
            rs, rit, rq = z[f"{name}_{tag}_success"], z[f"{name}_{tag}_it"], z[f"{name}_{tag}_q"]
            if damp > 0:
                assert (s == rs).all() and (it == rit).all(), (name, tag)
                ok = s == 1
                np.testing.assert_allclose(q[ok], rq[ok], atol=1e-8)
                np.testing.assert_allclose(E[ok], z[f"{name}_{tag}_E"][ok], atol=1e-12)
            else:
                assert abs(s.mean() - rs.mean()) < 0.06, (name, tag, s.mean(), rs.mean())
                near = ((s == rs) & (it == rit))[:half]
                assert near.mean() >= 0.9, (name, tag, near.mean())


def test_rng_stream_is_stable():
    """The restart RNG is part of the product's contract: pin a few draws."""
    u = [orc.rand_u01(0, 0, 0, 0), orc.rand_u01(1, 2, 3, 4), orc.rand_u01(2**63, 10**9, 99, 6)]
    assert all(0.0 <= x < 1.0 for x in u)
    np.testing.assert_allclose(u, [orc.rand_u01(0, 0, 0, 0), orc.rand_u01(1, 2, 3, 4), orc.rand_u01(2**63, 10**9, 99, 6)])
    xs = np.array([orc.rand_u01(7, r, s, j) for r in range(20) for s in range(5) for j in range(7)])
    assert abs(xs.mean() - 0.5) < 0.05 and len(np.unique(xs)) == xs.size


# ------------------------------------------------------------------ (c) live against the compiled reference
@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_live_against_compiled_reference():
    rng = np.random.default_rng(42)
    for t in range(6):
        d = ch.random_chain(rng, n_joints=int(rng.integers(1, 10)))
        C, R = orc.Chain(d), ref.RefETS(d)
        Q = rng.uniform(-4, 4, (40, C.n))
        tool = ch.trotx(0.3) @ ch.transl(0.1, 0.2, 0.3)
        base = ch.trotz(-0.7) @ ch.transl(1, 2, 3)
        np.testing.assert_allclose(C.fkine(Q, base, tool), R.fkine_rows(Q, base, tool), atol=1e-13)
        np.testing.assert_allclose(C.jacob0(Q, tool), R.jacob0(Q, tool), atol=1e-13)
        np.testing.assert_allclose(C.jacobe(Q, tool), R.jacobe(Q, tool), atol=1e-13)
    # the reference's own batch entry (config 1: Panda, batch 1024)
    d = ch.panda_ets()
    Q = np.random.default_rng(0).uniform(-np.pi, np.pi, (1024, 7))
    np.testing.assert_allclose(orc.Chain(d).fkine(Q), ref.RefETS(d).fkine(Q), atol=1e-14)


def test_jtraj_restatement_properties():
    """oracle.jtraj (tools/trajectory.py:730-775) against the reference test's properties
    (tests/test_trajectory.py:420-520) and the polynomial's defining boundary conditions."""
    q1 = np.r_[1, 2, 3, 4, 5, 6].astype(float)
    tv, q, qd, qdd = orc.jtraj(q1, -q1, 11)
    assert q.shape == qd.shape == qdd.shape == (11, 6)
    assert np.allclose(q[0], q1) and np.allclose(q[-1], -q1) and np.allclose(q[5], 0)
    assert np.allclose(qd[0], 0) and np.allclose(qd[-1], 0) and np.allclose(qdd[[0, 5, -1]], 0)
    t = np.linspace(0, 3, 301)
    tv, q, qd, qdd = orc.jtraj(q1, -q1, t, 0.3 * q1, -0.1 * q1)
    np.testing.assert_allclose(qd[0], 0.3 * q1, atol=1e-12)
    np.testing.assert_allclose(qd[-1], -0.1 * q1, atol=1e-10)
    np.testing.assert_allclose(np.gradient(q, t, axis=0)[5:-5], qd[5:-5], atol=2e-3)


def test_quintic_trapezoidal_restatements():
    """oracle.quintic / trapezoidal against the reference tests' properties (tests/test_trajectory.py:19-153)."""
    t, s, sd, sdd = orc.quintic(1, 2, 11)
    assert np.all(np.diff(s) > 0) and abs(s[0] - 1) < 1e-9 and abs(s[-1] - 2) < 1e-9 and abs(s[5] - 1.5) < 1e-9
    assert abs(sd[0]) < 1e-9 and abs(sd[-1]) < 1e-9 and abs(sdd[[0, 5, -1]]).max() < 1e-9 and abs(sdd.sum()) < 1e-9
    t, s, sd, sdd = orc.quintic(1, 2, 11, -1, 1)
    assert abs(sd[0] + 1) < 1e-9 and abs(sd[-1] - 1) < 1e-9 and abs(sdd[0]) < 1e-9 and abs(sdd[-1]) < 1e-9
    t, s, sd, sdd, tb = orc.trapezoidal(1, 2, 11)
    assert np.all(np.diff(s) > 0) and abs(s[0] - 1) < 1e-12 and abs(s[-1] - 2) < 1e-12 and abs(s[5] - 1.5) < 1e-12
    assert abs(sd[0]) < 1e-12 and abs(sd[-1]) < 1e-12 and abs(sd[5] - 0.15) < 1e-12  # V = 1.5 (qf-q0)/T
    with pytest.raises(ValueError):
        orc.trapezoidal(1, 2, 11, 0.01)
    with pytest.raises(ValueError):
        orc.trapezoidal(1, 2, 11, 1.0)


def test_pose_representation_restatements_are_self_consistent():
    """tr2rpy / tr2eul / trlog / rotvelxform come from spatialmath (not under the reference tree): the oracle's
    restatements are pinned by identities instead -- round trips through the defining products, and the analytical
    Jacobian being the derivative of the representation along the chain."""
    from oracle import chains as ch

    rng = np.random.default_rng(3)
    rx, ry, rz = (lambda a: ch.trotx(a)[:3, :3]), (lambda a: ch.troty(a)[:3, :3]), (lambda a: ch.trotz(a)[:3, :3])
    for _ in range(50):
        r, p, y = rng.uniform(-3, 3), rng.uniform(-1.5, 1.5), rng.uniform(-3, 3)
        np.testing.assert_allclose(orc.tr2rpy(rz(y) @ ry(p) @ rx(r), "zyx"), [r, p, y], atol=1e-12)
        np.testing.assert_allclose(orc.tr2rpy(rx(y) @ ry(p) @ rz(r), "xyz"), [r, p, y], atol=1e-12)
        ph, th, ps = rng.uniform(-3, 3), rng.uniform(0.05, 3.0), rng.uniform(-3, 3)
        np.testing.assert_allclose(orc.tr2eul(rz(ph) @ ry(th) @ rz(ps)), [ph, th, ps], atol=1e-12)
    # gimbal lock: roll := 0 and the product is still reproduced
    for p in (np.pi / 2, -np.pi / 2):
        R = rz(0.3) @ ry(p) @ rx(0.7)
        g = orc.tr2rpy(R, "zyx")
        assert g[0] == 0
        np.testing.assert_allclose(rz(g[2]) @ ry(g[1]) @ rx(g[0]), R, atol=1e-12)
    C = orc.Chain(ch.panda_ets())
    Q = rng.uniform(-2, 2, (20, 7))
    T, J = C.fkine(Q), C.jacob0(Q)
    h = 1e-6
    for rep in ("rpy/xyz", "rpy/zyx", "eul", "exp"):
        Ja = orc.jacob0_analytical(T, J, rep)
        np.testing.assert_array_equal(Ja[:, :3], J[:, :3])
        for k in range(len(Q)):
            for j in range(7):
                dq = np.zeros(7); dq[j] = h
                gp = orc.r2x(C.fkine(Q[k] + dq)[0, :3, :3], rep)
                gm = orc.r2x(C.fkine(Q[k] - dq)[0, :3, :3], rep)
                d = gp - gm
                if rep != "exp":
                    d = (d + np.pi) % (2 * np.pi) - np.pi
                if np.abs(d).max() < 1e-3:  # skip samples that straddle a branch cut of the representation
                    np.testing.assert_allclose(Ja[k, 3:, j], d / (2 * h), rtol=2e-5, atol=2e-6, err_msg=f"{rep} row {k} joint {j}")
    # Cartesian interpolation: end points, linear translation, constant angular rate along the shorter arc
    T0, T1 = T[0], T[1]
    s = np.linspace(0, 1, 11)
    P = orc.ctraj_poses(T0, T1, s)
    np.testing.assert_allclose(P[0], T0, atol=1e-12)
    np.testing.assert_allclose(P[-1], T1, atol=1e-12)
    np.testing.assert_allclose(P[:, :3, 3], T0[:3, 3] + s[:, None] * (T1[:3, 3] - T0[:3, 3]), atol=1e-12)
    ang = np.array([np.linalg.norm(orc.trlog(T0[:3, :3].T @ p[:3, :3])) for p in P])
    np.testing.assert_allclose(ang, s * ang[-1], atol=1e-10)
    assert ang[-1] <= np.pi + 1e-12
    # mstraj: passes through the neighbourhood of the via points, starts at q0, ends at the last via point, dt grid
    via = np.array([[0.0, 0.0], [1.0, 0.5], [1.0, 2.0], [-0.5, 2.0]])
    t, q, arrive = orc.mstraj(via, dt=0.1, tacc=0.4, qdmax=[1.0, 0.8])
    np.testing.assert_allclose(np.diff(t), 0.1)
    np.testing.assert_allclose(q[-1], via[-1], atol=1e-12)
    assert np.abs(q[0] - via[0]).max() < 0.1 and (np.diff(arrive) > 0).all()
    assert (np.abs(np.diff(q, axis=0)).max(axis=0) / 0.1 <= 1.25 * np.array([1.0, 0.8])).all()  # blends overshoot the cruise speed a little


def test_trajectory_restatements_reproduce_the_reference_kats():
    """mstraj / ctraj literals of the reference's own tests (tests/test_trajectory.py:175-204, 592-631)."""
    import json
    import os

    K = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))["mstraj"]
    via = np.array(K["via"])
    c = K["qdmax_case"]
    np.testing.assert_array_almost_equal(orc.mstraj(via, dt=c["dt"], tacc=c["tacc"], qdmax=c["qdmax"], q0=c["q0"])[1], np.array(c["q"]), decimal=4)
    c = K["tsegment_case"]
    np.testing.assert_array_almost_equal(orc.mstraj(via, dt=c["dt"], tacc=c["tacc"], tsegment=c["tsegment"], q0=c["q0"])[1], np.array(c["q"]), decimal=4)
    from oracle import chains as ch

    s3 = orc.trapezoidal(0, 1, 3)[1]
    for T0, T1 in ((ch.transl(1, 2, 3), ch.transl(-1, -2, -3)), (ch.trotx(-np.pi / 2), ch.trotx(np.pi / 2))):
        P = orc.ctraj_poses(T0, T1, s3)
        np.testing.assert_array_almost_equal(P[0], T0)
        np.testing.assert_array_almost_equal(P[2], T1)
        np.testing.assert_array_almost_equal(P[1], np.eye(4))
