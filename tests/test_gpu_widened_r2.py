"""GPU parity of the rows widened in round 2: singular-value manipulability measures, ..."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import b2kin as rtb  # noqa: E402
from oracle import chains as ch  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def dev(a, dt=np.float64):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).cuda()


def host(t):
    return t.cpu().numpy()


def test_manipulability_minsingular_and_invcondition():
    """ETS.manipulability(method="minsingular" | "invcondition") (ETS.py:1789-1796) against numpy's svd / cond on the
    oracle's Jacobians: Panda (6x7: wide), a 4-joint chain (6x4: tall), axis subsets, a singular configuration."""
    rng = np.random.default_rng(12)
    panda = rtb.models.Panda().ets()
    C = orc.Chain(panda.describe())
    Q = rng.uniform(-2.5, 2.5, (500, 7))
    Q[0] = 0.0  # Panda at zero is singular
    J = C.jacob0(Q)
    for axes, mask in (("all", [True] * 6), ("trans", [True] * 3 + [False] * 3), ("rot", [False] * 3 + [True] * 3),
                       ([True, False, True, False, True, True], [True, False, True, False, True, True])):
        for method in ("minsingular", "invcondition"):
            got = host(panda.manipulability(dev(Q), method=method, axes=axes))
            np.testing.assert_allclose(got, orc.manip_svd(J, mask, method), rtol=1e-9, atol=1e-12, err_msg=f"{axes} {method}")
    assert panda.manipulability(Q[0], method="minsingular") < 1e-10
    # from a given J, numpy in -> numpy out, one configuration -> float
    m = panda.manipulability(J=J[5], method="invcondition")
    assert isinstance(m, float) and abs(m - orc.manip_svd(J[5], kind="invcondition")[0]) < 1e-12
    ET = rtb.ET
    e4 = ET.Rz() * ET.tx(0.3) * ET.Ry() * ET.tz(0.2) * ET.Rx() * ET.tx(0.1) * ET.Rz()
    C4 = orc.Chain(e4.describe())
    Q4 = rng.uniform(-2, 2, (300, 4))
    for method in ("minsingular", "invcondition"):
        np.testing.assert_allclose(host(e4.manipulability(dev(Q4), method=method)), orc.manip_svd(C4.jacob0(Q4), kind=method),
                                   rtol=1e-9, atol=1e-12)
    got32 = host(panda.manipulability(dev(Q[1:], np.float32), method="minsingular"))
    np.testing.assert_allclose(got32, orc.manip_svd(J[1:], kind="minsingular"), rtol=2e-3, atol=2e-5)
    with pytest.raises(ValueError):
        panda.manipulability(Q[1], method="asada")
