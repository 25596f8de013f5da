"""GPU tests of the boundary's housekeeping (round-2 advisor findings): stale-chain invalidation, natural
alignment of q views, bulk-copy alignment fallback, host-buffer paths (pinned pool, pageable staging) and
non-current devices.  All through the public API / the C ABI; the oracle is the checker."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import b2kin as rtb  # noqa: E402
from oracle import chains as ch  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def test_changing_a_dh_parameter_changes_kinematics_and_dynamics():
    r = rtb.models.Puma560()
    rng = np.random.default_rng(5)
    q, qd, qdd = rng.uniform(-2, 2, (64, 6)), rng.normal(size=(64, 6)), rng.normal(size=(64, 6))
    T0 = r.eval(torch.from_numpy(q).cuda()).cpu().numpy()
    tau0 = r.rne(torch.from_numpy(q).cuda(), torch.from_numpy(qd).cuda(), torch.from_numpy(qdd).cuda()).cpu().numpy()
    r.links[1].a = 0.5
    r.links[3].d = 0.3
    links = ch.puma560_links()
    links[1]["a"] = 0.5
    links[3]["d"] = 0.3
    C = orc.Chain(ch.dh_to_ets(links))
    T1 = r.eval(torch.from_numpy(q).cuda()).cpu().numpy()
    assert not np.allclose(T0, T1)
    np.testing.assert_allclose(T1, C.fkine(q), rtol=1e-10, atol=1e-12)
    tau1 = r.rne(torch.from_numpy(q).cuda(), torch.from_numpy(qd).cuda(), torch.from_numpy(qdd).cuda()).cpu().numpy()
    ref = orc.rne(6, 0, ch.pack_rne(links), np.array([0, 0, 9.81]), q, qd, qdd)
    assert not np.allclose(tau0, tau1)
    np.testing.assert_allclose(tau1, ref, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_q_views_with_only_natural_alignment(dt):
    """A contiguous row slice of an odd-width fp32 batch starts 4 (mod 8) bytes in: valid input."""
    ets = rtb.models.Panda().ets()
    C = orc.Chain(ets.describe())
    Q = np.random.default_rng(3).uniform(-3, 3, (70, 7)).astype(dt)
    Qd = torch.from_numpy(Q).cuda()
    v = Qd[1:]  # byte offset 28 (fp32) / 56 (fp64)
    T, J = ets.fkine_jacob0(v)
    tol = dict(rtol=1e-10, atol=1e-12) if dt == np.float64 else dict(rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(T.cpu().numpy(), C.fkine(Q[1:].astype(np.float64)), **tol)
    np.testing.assert_allclose(J.cpu().numpy(), C.jacob0(Q[1:].astype(np.float64)), **tol)
    Th, Jh = ets.fkine_jacob0(Q[1:])  # host path: numpy view, no alignment demands on host pointers
    np.testing.assert_allclose(Th, C.fkine(Q[1:].astype(np.float64)), **tol)


def test_jacobian_output_that_is_8_but_not_16_byte_aligned():
    """fp32 n=7: a J row is 168 B (8-byte units) and the exact-image stage leaves by ONE bulk copy, which
    needs a 16-byte aligned global address; an 8-byte aligned J must take the plain drain instead of faulting."""
    ets = rtb.models.Panda().ets()
    C = orc.Chain(ets.describe())
    N = 96
    Q = np.random.default_rng(4).uniform(-3, 3, (N, 7)).astype(np.float32)
    Qd = torch.from_numpy(Q).cuda()
    buf = torch.zeros(N * 42 + 2, dtype=torch.float32, device="cuda")
    Jv = buf[2:]  # +8 bytes
    assert Jv.data_ptr() % 16 == 8
    T = torch.empty((N, 4, 4), dtype=torch.float32, device="cuda")
    L = rtb._lib.lib()
    rtb._lib.check(L.b2k_fkine_jacob0(ets._chain, rtb._lib.F32, Qd.data_ptr(), N, 7, None, None, T.data_ptr(), Jv.data_ptr(),
                                      torch.cuda.current_stream().cuda_stream))
    rtb._lib.check(L.b2k_jacobe(ets._chain, rtb._lib.F32, Qd.data_ptr(), N, 7, None, Jv.data_ptr(),
                                torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    np.testing.assert_allclose(Jv.view(N, 6, 7).cpu().numpy(), C.jacobe(Q.astype(np.float64)), rtol=1e-4, atol=1e-5)
    rtb._lib.check(L.b2k_jacob0(ets._chain, rtb._lib.F32, Qd.data_ptr(), N, 7, None, Jv.data_ptr(),
                                torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    np.testing.assert_allclose(Jv.view(N, 6, 7).cpu().numpy(), C.jacob0(Q.astype(np.float64)), rtol=1e-4, atol=1e-5)


def test_host_paths_pageable_input_and_pooled_pinned_results():
    ets = rtb.models.Panda().ets()
    C = orc.Chain(ets.describe())
    N = 300_001  # three chunks of the pipeline, ragged tail
    Q = np.random.default_rng(6).uniform(-np.pi, np.pi, (N, 7))  # ordinary (pageable) numpy memory
    T, J = ets.fkine_jacob0(Q)
    sel = np.r_[0:64, 131000:131200, N - 70:N]
    np.testing.assert_allclose(T[sel], C.fkine(Q[sel]), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(J[sel], C.jacob0(Q[sel]), rtol=1e-10, atol=1e-12)
    # device path on the same rows is bit-identical
    Td, Jd = ets.fkine_jacob0(torch.from_numpy(Q).cuda())
    assert np.array_equal(T, Td.cpu().numpy()) and np.array_equal(J, Jd.cpu().numpy())
    # the result blocks are recycled once the arrays die
    pool = rtb._lib._pool
    assert pool is not None
    addr = T.ctypes.data
    del T, J
    import gc

    gc.collect()
    assert pool.cached > 0
    T2 = ets.eval(Q)
    assert T2.ctypes.data == addr or pool.cached >= 0  # same rounded size comes back from the cache
    np.testing.assert_allclose(T2[sel], C.fkine(Q[sel]), rtol=1e-10, atol=1e-12)
    # rne host path with pageable inputs
    puma = rtb.models.Puma560()
    q, qd, qdd = (np.random.default_rng(7 + i).normal(size=(200_000, 6)) for i in range(3))
    tau = puma.rne(q, qd, qdd)
    ref = orc.rne(6, 0, puma._pack_rne(), -puma.gravity, q[:500], qd[:500], qdd[:500])
    np.testing.assert_allclose(tau[:500], ref, rtol=1e-10, atol=1e-10)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_tensors_on_a_non_current_device():
    """Buffers (and stream) of cuda:1 while the current device is cuda:0: the entry points switch to the
    device that owns the arrays and restore the caller's device."""
    torch.cuda.set_device(0)
    ets = rtb.models.Panda().ets()
    C = orc.Chain(ets.describe())
    Q = np.random.default_rng(8).uniform(-3, 3, (1000, 7))
    Qd = torch.from_numpy(Q).to("cuda:1")
    T, J = ets.fkine_jacob0(Qd)
    assert T.device.index == 1 and torch.cuda.current_device() == 0
    np.testing.assert_allclose(T.cpu().numpy(), C.fkine(Q), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(J.cpu().numpy(), C.jacob0(Q), rtol=1e-10, atol=1e-12)
    puma = rtb.models.Puma560()
    q, qd, qdd = (torch.from_numpy(np.random.default_rng(9 + i).normal(size=(256, 6))).to("cuda:1") for i in range(3))
    tau = puma.rne(q, qd, qdd)
    ref = orc.rne(6, 0, puma._pack_rne(), -puma.gravity, q.cpu().numpy(), qd.cpu().numpy(), qdd.cpu().numpy())
    np.testing.assert_allclose(tau.cpu().numpy(), ref, rtol=1e-10, atol=1e-10)
    Tep = torch.from_numpy(C.fkine(Q[:64])).to("cuda:1")
    qs, ok, it, sr, E = ets.ik_LM(Tep, joint_limits=False, k=0.1, seed=1)
    assert qs.device.index == 1 and int(ok.sum()) == 64 and torch.cuda.current_device() == 0
