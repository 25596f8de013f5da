#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ from the reference's OWN native code.

Run in the build container (where /root/reference exists):

    make -C oracle ref            # compiles fknm/frne from /root/reference into oracle/_ref/
    python tests/golden/make_golden.py

Every output array below is produced by the compiled reference modules
(oracle/ref_driver.py -> oracle/_ref/fknm*.so, frne*.so); inputs are seeded numpy
draws.  The fixtures are small (a few hundred rows each) and committed, so the
GPU box -- which has no /root/reference -- can replay them.

Fixture            reference entry point                      BASELINE config
panda_fkj.npz      fknm.ETS_fkine / ETS_jacob0 / ETS_jacobe    1, 2
ur10_fkj.npz       same, UR10 DH->ETS chain                    5
random_fkj.npz     same, 12 random chains (all ET kinds, flips, base, tool)
puma_rne.npz       frne.frne, Puma560 standard DH               3
panda_mdh_rne.npz  frne.frne, Panda modified DH
random_rne.npz     frne.frne, random DH/MDH links incl. prismatic
panda_ik.npz       fknm.IK_LM_c (explicit q0, slimit=1; and with restarts)
ik_nr_gn.npz       fknm.IK_NR_c / IK_GN_c (explicit q0, slimit=1): Panda n=7, UR10 n=6, a 3-joint chain
                   (`python make_golden.py iknr` regenerates this file alone)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import chains as ch  # noqa: E402
from oracle import ref_driver as ref  # noqa: E402


def desc_arrays(d, prefix=""):
    return {prefix + k: np.asarray(v) for k, v in d.items() if k != "n"}


def fkj(desc, Q, base=None, tool=None):
    R = ref.RefETS(desc)
    return dict(Tfk=R.fkine_rows(Q, base, tool), J0=R.jacob0(Q, tool), Je=R.jacobe(Q, tool))


def ik_nr_gn(out):
    """Newton-Raphson / Gauss-Newton fixtures.  With pinv_damping > 0 the iteration is well conditioned and
    the outputs are reproducible to rounding; undamped runs are chaotic near singular configurations, so
    consumers compare those statistically (success rate, iteration histogram)."""
    pack = {}
    ur = ch.dh_to_ets(ch.ur10_links())
    m3 = len(ur["isjoint"])
    # first three joints of the UR10 chain: cut after the ET that precedes joint 4
    jpos = [i for i in range(m3) if ur["isjoint"][i]]
    cut = jpos[3]
    ur3 = {k: (np.asarray(v)[:cut] if k != "n" else 3) for k, v in ur.items()}
    for name, d, n in (("panda", ch.panda_ets(), 7), ("ur10", ur, 6), ("ur3", ur3, 3)):
        R = ref.RefETS(d)
        rng = np.random.default_rng(11)
        N = 160
        qs = rng.uniform(-2.6, 2.6, (N, n))
        Tep = R.fkine_rows(qs)
        q0 = qs + rng.normal(0, 0.35, (N, n))
        q0[N // 2:] = rng.uniform(-np.pi, np.pi, (N - N // 2, n))  # second half: unrelated starts
        pack.update({f"{name}_{k}": v for k, v in desc_arrays(d).items()})
        pack.update({f"{name}_qs": qs, f"{name}_Tep": Tep, f"{name}_q0": q0})
        mask = np.array([1, 1, 1, 0, 0, 0.0])
        runs = [("nr_d", ref.ik_nr, dict(pinv_damping=0.1)), ("nr", ref.ik_nr, dict(pinv_damping=0.0)),
                ("gn", ref.ik_gn, {}), ("gn_mask", ref.ik_gn, dict(mask=mask)),
                ("nr_d_mask", ref.ik_nr, dict(pinv_damping=0.1, mask=mask))]
        if n == 6:
            runs.append(("nr_inv", ref.ik_nr, dict(pinv=False)))
        for tag, fn, kw in runs:
            q, s, it, sr, E = fn(R, Tep, q0=q0, ilimit=30, slimit=1, tol=1e-6, joint_limits=False, **kw)
            pack.update({f"{name}_{tag}_q": q, f"{name}_{tag}_success": s, f"{name}_{tag}_it": it, f"{name}_{tag}_E": E})
    pack["mask"] = mask
    np.savez(os.path.join(out, "ik_nr_gn.npz"), **pack)


def main():
    out = HERE
    if len(sys.argv) > 1 and sys.argv[1] == "iknr":
        ik_nr_gn(out)
        return
    # ---------------- Panda (config 1/2 inputs: default_rng(0).uniform(-pi,pi,(N,7)))
    d = ch.panda_ets()
    Q = np.random.default_rng(0).uniform(-np.pi, np.pi, (1024, 7))[:256]
    Q[-1] = [1.4, 0.2, 1.8, 0.7, 0.1, 3.1, 2.9]  # the KAT configuration
    Q[-2] = 0.0
    Q[-3] = [0, -0.3, 0, -2.2, 0, 2.0, np.pi / 4]  # qr
    Q[-4] = 1e6 * np.array([1, -1, 0.5, -0.5, 0.25, -0.25, 0.125])  # huge angles (range reduction)
    np.savez(os.path.join(out, "panda_fkj.npz"), Q=Q, **desc_arrays(d), **fkj(d, Q))

    # ---------------- UR10 DH -> ETS
    d = ch.dh_to_ets(ch.ur10_links())
    Q = np.random.default_rng(3).uniform(-np.pi, np.pi, (256, 6))
    np.savez(os.path.join(out, "ur10_fkj.npz"), Q=Q, **desc_arrays(d), **fkj(d, Q))

    # ---------------- random chains
    pack = {}
    rng = np.random.default_rng(1234)
    nch = 12
    for c in range(nch):
        n = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 6, 7][c]
        d = ch.random_chain(rng, n_joints=n, with_flips=(c % 2 == 0), extra_consts=(c != 5))
        Q = rng.uniform(-3.5, 3.5, (48, n))
        base = ch.trotz(rng.uniform(-3, 3)) @ ch.trotx(rng.uniform(-3, 3)) @ ch.transl(*rng.uniform(-1, 1, 3)) if c % 3 else None
        tool = ch.troty(rng.uniform(-3, 3)) @ ch.transl(*rng.uniform(-0.3, 0.3, 3)) if c % 4 != 1 else None
        r = fkj(d, Q, base, tool)
        pack.update(desc_arrays(d, f"c{c}_"))
        pack[f"c{c}_Q"] = Q
        pack[f"c{c}_base"] = np.zeros((0,)) if base is None else base
        pack[f"c{c}_tool"] = np.zeros((0,)) if tool is None else tool
        for k, v in r.items():
            pack[f"c{c}_{k}"] = v
    pack["nchains"] = np.array(nch)
    np.savez(os.path.join(out, "random_fkj.npz"), **pack)

    # ---------------- Puma560 RNE
    links = ch.puma560_links()
    L = ch.pack_rne(links)
    g = np.array([0, 0, -9.81])
    R = ref.RefRNE(6, 0, L, g)
    rng = np.random.default_rng(1)
    ql = np.array([l["qlim"] for l in links])
    N = 256
    q = rng.uniform(ql[:, 0], ql[:, 1], (N, 6))
    qd = rng.normal(size=(N, 6))
    qdd = rng.normal(size=(N, 6))
    qd[-32:-16] = 0.0  # Coulomb qd==0 branch (ne.c:487-490)
    qd[-16:, ::2] = 0.0
    fext = np.array([1, 2, 3, 1, 2, 3.0])
    g2 = np.array([1.0, -2.0, 3.0])
    np.savez(os.path.join(out, "puma_rne.npz"), L=L, gravity=g, q=q, qd=qd, qdd=qdd, fext=fext, g2=g2,
             tau=R.rne(q, qd, qdd), tau_fext=R.rne(q, qd, qdd, fext=fext),
             tau_zerog=R.rne(q, qd, qdd, gravity=[0, 0, 0]), tau_g2=R.rne(q, qd, qdd, gravity=g2, fext=fext))

    # ---------------- Panda MDH RNE
    Lm = ch.pack_rne(ch.panda_mdh_links())
    Rm = ref.RefRNE(7, 1, Lm, g)
    q = rng.uniform(-2.8, 2.8, (N, 7))
    qd = rng.normal(size=(N, 7))
    qdd = rng.normal(size=(N, 7))
    np.savez(os.path.join(out, "panda_mdh_rne.npz"), L=Lm, gravity=g, q=q, qd=qd, qdd=qdd, fext=fext,
             tau=Rm.rne(q, qd, qdd), tau_fext=Rm.rne(q, qd, qdd, fext=fext))

    # ---------------- random DH / MDH links incl. prismatic joints
    pack = {}
    rng = np.random.default_rng(77)
    ncase = 8
    for c in range(ncase):
        n = [1, 2, 3, 4, 6, 7, 9, 5][c]
        mdh = c % 2
        links = []
        for j in range(n):
            A = rng.normal(size=(3, 3))
            links.append(dict(
                sigma=int(rng.random() < 0.35), theta=float(rng.uniform(-1, 1)), d=float(rng.uniform(-0.5, 0.5)),
                alpha=float(rng.choice([0.0, np.pi / 2, -np.pi / 2, 0.3])), a=float(rng.uniform(-0.5, 0.5)),
                offset=float(rng.choice([0.0, 0.4])), m=float(rng.uniform(0, 5)), r=rng.uniform(-0.2, 0.2, 3),
                I=A @ A.T * 0.1, Jm=float(rng.uniform(0, 1e-3)), G=float(rng.uniform(-100, 100)),
                B=float(rng.uniform(0, 1e-3)), Tc=[float(rng.uniform(0, 0.5)), float(-rng.uniform(0, 0.5))]))
        L = ch.pack_rne(links)
        gg = rng.normal(size=3) * 5
        Rr = ref.RefRNE(n, mdh, L, gg)
        q = rng.uniform(-3, 3, (32, n)); qd = rng.normal(size=(32, n)); qdd = rng.normal(size=(32, n))
        qd[:4] = 0
        fx = rng.normal(size=6)
        pack[f"c{c}_L"] = L; pack[f"c{c}_mdh"] = np.array(mdh); pack[f"c{c}_gravity"] = gg
        pack[f"c{c}_q"] = q; pack[f"c{c}_qd"] = qd; pack[f"c{c}_qdd"] = qdd; pack[f"c{c}_fext"] = fx
        pack[f"c{c}_tau"] = Rr.rne(q, qd, qdd, fext=fx)
    pack["ncases"] = np.array(ncase)
    np.savez(os.path.join(out, "random_rne.npz"), **pack)

    # ---------------- dynamics fan-outs (Dynamics.py loops over the compiled frne), Puma560
    from oracle import oracle as orc_mod

    links = ch.puma560_links()
    L = ch.pack_rne(links)
    Rf = ref.RefRNE(6, 0, L, g)
    Rnf = ref.RefRNE(6, 0, orc_mod.nofriction_L(L), g)
    f_fric = lambda q, qd, qdd, grav: Rf.rne(q, qd, qdd, gravity=grav)      # noqa: E731
    f_nofr = lambda q, qd, qdd, grav: Rnf.rne(q, qd, qdd, gravity=grav)     # noqa: E731
    rng = np.random.default_rng(11)
    Nd = 24
    q = rng.uniform(-2.5, 2.5, (Nd, 6)); qd = rng.normal(size=(Nd, 6)); qdd = rng.normal(size=(Nd, 6))
    torque = rng.normal(size=(Nd, 6)) * 5
    q[0] = ch.PUMA_QN
    np.savez(os.path.join(out, "puma_dynamics.npz"), L=L, gravity=g, q=q, qd=qd, qdd=qdd, torque=torque,
             inertia=orc_mod.dyn_inertia(f_fric, 6, q), gravload=orc_mod.dyn_gravload(f_fric, 6, q, g),
             itorque=orc_mod.dyn_itorque(f_fric, 6, q, qdd), coriolis=orc_mod.dyn_coriolis(f_nofr, 6, q, qd),
             accel=orc_mod.dyn_accel(f_fric, 6, q, qd, torque, g))

    # ---------------- manipulator Hessians: compiled fknm + the literal golden of the reference's own test
    d = ch.panda_ets()
    R = ref.RefETS(d)
    Qh = np.random.default_rng(5).uniform(-np.pi, np.pi, (40, 7))
    Qh[0] = [1.4, 0.2, 1.8, 0.7, 0.1, 3.1, 2.9]
    toolh = ch.trotx(0.3) @ ch.transl(0.1, 0.2, 0.3)
    kat = np.zeros((0,))
    tpath = "/root/reference/tests/test_ETS.py"
    if os.path.exists(tpath):  # transcribe `ans` of test_hessian0 (tests/test_ETS.py:718-1111) mechanically
        src = open(tpath).read()
        a = src.index("def test_hessian0(self)")
        a = src.index("ans = np.array(", a) + len("ans = ")
        depth, b = 0, a + len("np.array")
        while True:
            ch_ = src[b]
            depth += ch_ == "("
            depth -= ch_ == ")"
            b += 1
            if depth == 0:
                break
        ans = eval(src[a:b], {"np": np})
        kat = np.stack([ans[:, :, i] for i in range(7)])  # ans_new of tests/test_ETS.py:1113-1116
    np.savez(os.path.join(out, "panda_hessian.npz"), Q=Qh, tool=toolh, **desc_arrays(d), J0=R.jacob0(Qh), Je=R.jacobe(Qh),
             H0=R.hessian0(Qh), He=R.hessiane(Qh), H0_tool=R.hessian0(Qh, toolh), kat_hessian0_q1=kat)

    # ---------------- Panda IK (config 4 protocol: reachable targets Tep = FK(q*))
    d = ch.panda_ets()
    R = ref.RefETS(d)
    rng = np.random.default_rng(2)
    N = 192
    qs = rng.uniform(-np.pi, np.pi, (N, 7))
    Tep = R.fkine_rows(qs)
    q0 = rng.uniform(-np.pi, np.pi, (N, 7))
    pack = dict(qs=qs, Tep=Tep, q0=q0, **desc_arrays(d))
    for tag, method, k in (("chan1", "chan", 1.0), ("chan01", "chan", 0.1), ("sugi", "sugihara", 1e-4),
                           ("wamp", "wampler", 1e-2)):
        q, s, it, sr, E = R.ik_lm(Tep, q0=q0, ilimit=30, slimit=1, tol=1e-6, joint_limits=False, k=k, method=method)
        pack.update({f"{tag}_q": q, f"{tag}_success": s, f"{tag}_it": it, f"{tag}_search": sr, f"{tag}_E": E})
    # explicit q0 + joint limits (exercises the fmod wrap + limit rejection, ik.cpp:50-52)
    q, s, it, sr, E = R.ik_lm(Tep, q0=q0, ilimit=30, slimit=1, tol=1e-6, joint_limits=True, k=1.0, method="chan")
    pack.update(jl_q=q, jl_success=s, jl_it=it, jl_search=sr, jl_E=E)
    # masked solve (position only)
    mask = np.array([1, 1, 1, 0, 0, 0.0])
    q, s, it, sr, E = R.ik_lm(Tep, q0=q0, ilimit=30, slimit=1, tol=1e-6, joint_limits=False, mask=mask, k=1.0, method="chan")
    pack.update(mask=mask, mask_q=q, mask_success=s, mask_it=it, mask_E=E)
    # restarts (unseeded libc rand in the reference: outcome statistics only)
    q, s, it, sr, E = R.ik_lm(Tep, q0=None, ilimit=30, slimit=100, tol=1e-6, joint_limits=True, k=1.0, method="chan")
    pack.update(rs_success=s, rs_it=it, rs_search=sr, rs_E=E, rs_q=q)
    np.savez(os.path.join(out, "panda_ik.npz"), **pack)

    ik_nr_gn(out)

    # angle-axis corner cases (ik.cpp:261-277)
    Ts = [np.eye(4), ch.trotx(np.pi), ch.troty(np.pi) @ ch.transl(1, 2, 3), ch.trotz(np.pi), ch.trotx(1e-7),
          ch.trotz(3.1415) @ ch.trotx(0.2), ch.trotx(0.3) @ ch.troty(-1.2) @ ch.transl(0.1, 0.2, 0.3)]
    pairs = [(a, b) for a in Ts for b in Ts]
    Te = np.stack([p[0] for p in pairs]); Tp = np.stack([p[1] for p in pairs])
    e = np.stack([ref.angle_axis(a, b) for a, b in pairs])
    np.savez(os.path.join(out, "angle_axis.npz"), Te=Te, Tep=Tp, e=e)
    print("golden fixtures written to", out)
    for f in sorted(os.listdir(out)):
        if f.endswith(".npz"):
            print(f"  {f}: {os.path.getsize(os.path.join(out, f))} bytes")


if __name__ == "__main__":
    main()
