// b2k_rne_gen.cpp -- robot-specialised code generation for the recursive Newton-Euler kernels.
//
// The generic kernels (b2k_rne.cuh) multiply through every link parameter of the 24-double table
// whether it is zero or not: for a Puma560 that is 886 FP64 instructions per row, and the kernel is
// FP64-pipe bound (profiles/r01_rne_v3.txt).  Most of a real arm's table is structure, not numbers:
// alpha in {0, +-pi/2}, a or d zero, centre of mass on an axis, diagonal inertia, no tip wrench, gravity
// along one base axis, and the first links of the chain carry sparse velocities (w_0 = 0).  This file
// is the chain compiler for dynamics: it runs the Luh-Walker-Paul recursion (reference ne.c:137-457)
// SYMBOLICALLY over scalars that are either exactly zero, a known constant or a run-time value, and
// emits straight-line CUDA C in which a term whose factor is zero never appears, a factor of +-1 is
// a sign, constants are folded, and every sum of products is a chain of FMAs.  The text is compiled
// for sm_100a at run time (NVRTC, b2k_rne_spec.cu) into a kernel for THIS robot; the same text
// compiles as plain C++ for the host, which is how tests/ check it against the oracle without a GPU.
//
// The dynamics fan-outs of the reference's DynamicsMixin (Dynamics.py: inertia, gravload, itorque,
// coriolis, accel) are the same recursion evaluated with symbolic unit / zero inputs, so e.g. the
// column i of the inertia matrix costs only the links i..n-1 and nothing of the velocity terms.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <functional>
#include <map>
#include <string>
#include <vector>

#include "b2k_rne_gen.h"

namespace {

// ------------------------------------------------------------------ symbolic scalars
struct Opd { // an operand an instruction can take directly: a run-time variable or a constant-bank entry
    bool isc = false;
    double c = 0.0; // constant value (isc)
    int id = -1;    // variable id (!isc)
    bool neg = false;
};
struct Sym {
    enum Kind { Z, O, P } k = Z; // zero | operand | pending product x*y (sign in x)
    Opd x, y;
};

struct Gen {
    std::vector<std::string> code;
    std::vector<std::string> names; // variable id -> spelling
    std::vector<double> consts;     // constant bank
    std::map<unsigned long long, int> cslot;
    int n_mul = 0, n_fma = 0, n_add = 0;

    int new_var()
    {
        names.push_back("t" + std::to_string((int)names.size()));
        return (int)names.size() - 1;
    }
    int named_var(const std::string &s)
    {
        names.push_back(s);
        return (int)names.size() - 1;
    }
    static Sym zero() { return Sym(); }
    static Sym cst(double c)
    {
        Sym s;
        if (c == 0.0) return s;
        s.k = Sym::O; s.x.isc = true; s.x.c = c;
        return s;
    }
    Sym var(const std::string &name)
    {
        Sym s;
        s.k = Sym::O; s.x.id = named_var(name);
        return s;
    }
    std::string spell(const Opd &o, bool flip = false)
    {
        bool neg = o.neg != flip;
        if (o.isc) {
            double v = o.c;
            if (v < 0) { v = -v; neg = !neg; }
            unsigned long long bits;
            memcpy(&bits, &v, 8);
            auto it = cslot.find(bits);
            int slot;
            if (it == cslot.end()) { slot = (int)consts.size(); consts.push_back(v); cslot[bits] = slot; }
            else slot = it->second;
            return std::string(neg ? "-" : "") + "C[" + std::to_string(slot) + "]";
        }
        return std::string(neg ? "-" : "") + names[o.id];
    }
    Sym emit(const std::string &rhs)
    {
        int id = new_var();
        code.push_back("    const real " + names[id] + " = " + rhs + ";");
        Sym s;
        s.k = Sym::O; s.x.id = id;
        return s;
    }
    // a value an instruction can read: pending products are multiplied out
    Opd operand(const Sym &a)
    {
        if (a.k == Sym::O) return a.x;
        Sym m = emit(spell(a.x) + " * " + spell(a.y));
        n_mul++;
        return m.x;
    }
    Sym neg(Sym a)
    {
        if (a.k == Sym::Z) return a;
        if (a.x.isc) a.x.c = -a.x.c;
        else a.x.neg = !a.x.neg;
        return a;
    }
    Sym mul(const Sym &a, const Sym &b)
    {
        if (a.k == Sym::Z || b.k == Sym::Z) return zero();
        Opd x = operand(a), y = operand(b);
        if (x.isc && x.neg) { x.c = -x.c; x.neg = false; }
        if (y.isc && y.neg) { y.c = -y.c; y.neg = false; }
        if (x.isc && y.isc) return cst(x.c * y.c);
        if (y.isc) std::swap(x, y); // constant first
        if (x.isc && (x.c == 1.0 || x.c == -1.0)) {
            Sym s;
            s.k = Sym::O; s.x = y;
            if (x.c < 0) s = neg(s);
            return s;
        }
        Sym s;
        s.k = Sym::P;
        if (x.isc) std::swap(x, y); // keep the variable in x, the constant in y
        s.x = x; s.y = y;
        if (s.y.neg) { s.y.neg = false; s.x.neg = !s.x.neg; }
        if (s.y.isc && s.y.c < 0) { s.y.c = -s.y.c; s.x.neg = !s.x.neg; }
        return s;
    }
    // n-ary sum: constants folded, one chain of FMAs over the products, plain additions last
    Sym sum(const std::vector<Sym> &terms)
    {
        double K = 0.0;
        std::vector<Opd> plain;
        std::vector<Sym> prods;
        for (const Sym &t : terms) {
            if (t.k == Sym::Z) continue;
            if (t.k == Sym::P) { prods.push_back(t); continue; }
            if (t.x.isc) K += t.x.c;
            else plain.push_back(t.x);
        }
        if (K != 0.0) { Opd o; o.isc = true; o.c = K; plain.push_back(o); }
        if (plain.empty() && prods.empty()) return zero();
        if (plain.size() + prods.size() == 1) {
            if (!prods.empty()) return prods[0];
            Sym s; s.k = Sym::O; s.x = plain[0];
            return s;
        }
        Opd acc;
        size_t pi = 0, ai = 0;
        if (!plain.empty()) { acc = plain[0]; ai = 1; }
        else { acc = operand(prods[0]); pi = 1; }
        for (; pi < prods.size(); pi++) {
            acc = emit("fma(" + spell(prods[pi].x) + ", " + spell(prods[pi].y) + ", " + spell(acc) + ")").x;
            n_fma++;
        }
        for (; ai < plain.size(); ai++) {
            const Opd &b = plain[ai];
            const bool bneg = b.isc ? ((b.c < 0) != b.neg) : b.neg;
            Opd babs = b;
            babs.neg = false;
            if (babs.isc) babs.c = fabs(babs.c);
            std::string rhs;
            rhs = spell(acc) + (bneg ? " - " : " + ") + spell(babs);
            acc = emit(rhs).x;
            n_add++;
        }
        Sym s; s.k = Sym::O; s.x = acc;
        return s;
    }
    Sym add(const Sym &a, const Sym &b) { return sum({a, b}); }
    Sym sub(const Sym &a, const Sym &b) { return sum({a, neg(b)}); }
    // make sure a value that will be read several times is not a pending product
    Sym fix(const Sym &a)
    {
        if (a.k != Sym::P) return a;
        Sym s; s.k = Sym::O; s.x = operand(a);
        return s;
    }
};

struct V3 { Sym x, y, z; };

struct Vops {
    Gen &g;
    explicit Vops(Gen &gg) : g(gg) {}
    V3 zero() { return {Gen::zero(), Gen::zero(), Gen::zero()}; }
    V3 cst(const double *v) { return {Gen::cst(v[0]), Gen::cst(v[1]), Gen::cst(v[2])}; }
    V3 fix(const V3 &a) { return {g.fix(a.x), g.fix(a.y), g.fix(a.z)}; }
    V3 add(const V3 &a, const V3 &b) { return {g.add(a.x, b.x), g.add(a.y, b.y), g.add(a.z, b.z)}; }
    V3 scale(const Sym &s, const V3 &a) { return {g.mul(s, a.x), g.mul(s, a.y), g.mul(s, a.z)}; }
    // a x b + c (+ d), each component one FMA chain
    V3 cross_acc(const V3 &a, const V3 &b, const V3 *c = nullptr, const V3 *d = nullptr)
    {
        V3 A = fix(a), B = fix(b);
        auto comp = [&](const Sym &p, const Sym &q, const Sym &r, const Sym &s, const Sym *e, const Sym *f) {
            std::vector<Sym> t = {g.mul(p, q), g.neg(g.mul(r, s))};
            if (e) t.push_back(*e);
            if (f) t.push_back(*f);
            return g.sum(t);
        };
        return {comp(A.y, B.z, A.z, B.y, c ? &c->x : nullptr, d ? &d->x : nullptr),
                comp(A.z, B.x, A.x, B.z, c ? &c->y : nullptr, d ? &d->y : nullptr),
                comp(A.x, B.y, A.y, B.x, c ? &c->z : nullptr, d ? &d->z : nullptr)};
    }
    // planar rotations: Rz(th) v and Rz(-th) v with run-time (s, c); Rx(al) v and Rx(-al) v with (sa, ca)
    V3 rotz(const Sym &s, const Sym &c, const V3 &v, bool transpose, const V3 *acc = nullptr)
    {
        V3 a = fix(v);
        const Sym sx = transpose ? s : g.neg(s);
        // [c -s; s c] (or its transpose [c s; -s c]) on (x, y)
        return {g.sum({g.mul(c, a.x), g.mul(sx, a.y), acc ? acc->x : Gen::zero()}),
                g.sum({g.mul(g.neg(sx), a.x), g.mul(c, a.y), acc ? acc->y : Gen::zero()}),
                acc ? g.add(a.z, acc->z) : a.z};
    }
    V3 rotx(const Sym &s, const Sym &c, const V3 &v, bool transpose, const V3 *acc = nullptr)
    {
        V3 a = fix(v);
        const Sym sx = transpose ? s : g.neg(s);
        return {acc ? g.add(a.x, acc->x) : a.x,
                g.sum({g.mul(c, a.y), g.mul(sx, a.z), acc ? acc->y : Gen::zero()}),
                g.sum({g.mul(g.neg(sx), a.y), g.mul(c, a.z), acc ? acc->z : Gen::zero()})};
    }
    V3 matvec(const double *I /* row-major 3x3, read column-major like vmath.c mat_vect_mult */, const V3 &v)
    {
        V3 a = fix(v);
        auto row = [&](int r) {
            return g.sum({g.mul(Gen::cst(I[r]), a.x), g.mul(Gen::cst(I[3 + r]), a.y), g.mul(Gen::cst(I[6 + r]), a.z)});
        };
        return {row(0), row(1), row(2)};
    }
};

double snap(double v)
{ // sin / cos of alpha = k pi/2 come out of libm as 6.1e-17 or 1 - 1e-16: structure, not numbers
    const double r = nearbyint(v);
    if (fabs(v - r) < 4e-16 && fabs(r) <= 1.0) return r;
    return v;
}

struct Link {
    Sym st, ct, sa, ca; // joint rotation (run time; constants of theta for a prismatic link), twist (constants)
    bool pris = false;
    V3 ps;              // p*: (a, d sin(alpha), d cos(alpha)) for DH, (a, -d sin(alpha), d cos(alpha)) for MDH; d = q + offset if prismatic
    double r[3], I[9], m, c_jm, c_b, c_tcp, c_tcm;
};

struct Inputs { // one evaluation of the recursion: per-joint velocity / acceleration symbols, base acceleration, tip wrench
    std::vector<Sym> qd, qdd;
    V3 grav, ftip, ntip;
    bool friction = true;
};

// One Luh-Walker-Paul recursion over symbolic inputs; returns the joint torques as symbols.
// Standard DH: reference ne.c:245-347 (forward), 409-457 (backward); modified DH: ne.c:137-240, 358-403;
// joint torque with the actuator terms: ne.c:464-491.  All-revolute chains.
std::vector<Sym> recursion(Gen &g, const std::vector<Link> &L, bool mdh, const Inputs &in)
{
    Vops v(g);
    const int N = (int)L.size();
    std::vector<V3> Fm(N), Nm(N);
    V3 w = v.zero(), wd = v.zero(), acc = v.zero();
    auto RT = [&](const Link &l, const V3 &a, const V3 *add = nullptr) { // R^T a (+ add)
        if (!mdh) return v.rotx(l.sa, l.ca, v.rotz(l.st, l.ct, a, true), true, add); // R = Rz(th) Rx(al)
        return v.rotz(l.st, l.ct, v.rotx(l.sa, l.ca, a, true), true, add);          // R = Rx(al) Rz(th)
    };
    auto R = [&](const Link &l, const V3 &a, const V3 *add = nullptr) { // R a (+ add)
        if (!mdh) return v.rotz(l.st, l.ct, v.rotx(l.sa, l.ca, a, false), false, add);
        return v.rotx(l.sa, l.ca, v.rotz(l.st, l.ct, a, false), false, add);
    };
    for (int j = 0; j < N; j++) {
        const Link &l = L[j];
        const V3 ps = l.ps, rc = v.cst(l.r);
        const Sym qd = in.qd[j], qdd = in.qdd[j];
        V3 wn, wdn, accn;
        if (l.pris) { // translational joint along z_{j-1} (ne.c:183-225 MDH, 290-333 DH): no joint rate in w / wd
            const V3 qdv = {Gen::zero(), Gen::zero(), qd}, qddv = {Gen::zero(), Gen::zero(), qdd};
            const Sym two = Gen::cst(2.0);
            if (mdh) {
                if (j == 0) { // sic, ne.c:187-204: the base acceleration is taken over unrotated
                    wn = qdv; wdn = qddv; accn = in.grav;
                } else {
                    wn = RT(l, w);
                    wdn = RT(l, wd);
                    const V3 wxp = v.cross_acc(w, ps);
                    const V3 a = v.cross_acc(wd, ps, &acc);
                    const V3 t = RT(l, v.cross_acc(w, wxp, &a));
                    const V3 c2 = v.scale(two, v.cross_acc(v.fix(wn), qdv));
                    accn = v.add(v.add(t, c2), qddv);
                }
            } else {
                V3 base;
                if (j == 0) { wn = v.zero(); wdn = v.zero(); base = RT(l, v.add(qddv, in.grav)); }
                else { wn = v.fix(RT(l, w)); wdn = v.fix(RT(l, wd)); base = RT(l, v.add(qddv, acc)); }
                const V3 a1 = v.cross_acc(wdn, ps, &base);
                const V3 c2 = v.scale(two, v.cross_acc(wn, RT(l, qdv)));
                const V3 a2 = v.add(a1, c2);
                const V3 wxp = v.cross_acc(wn, ps);
                accn = v.cross_acc(wn, wxp, &a2);
            }
        } else if (mdh) {
            if (j == 0) {
                wn = {Gen::zero(), Gen::zero(), qd};
                wdn = {Gen::zero(), Gen::zero(), qdd};
                accn = RT(l, in.grav);
            } else {
                const V3 t1 = v.fix(RT(l, w));
                wn = {t1.x, t1.y, g.add(t1.z, qd)};
                const V3 t3 = RT(l, wd);
                // t1 x (0,0,qd) + t3 + (0,0,qdd)
                wdn = {g.sum({g.mul(t1.y, qd), t3.x}), g.sum({g.neg(g.mul(t1.x, qd)), t3.y}), g.add(t3.z, qdd)};
                const V3 wxp = v.cross_acc(w, ps);
                const V3 a = v.cross_acc(wd, ps, &acc);
                accn = RT(l, v.cross_acc(w, wxp, &a));
            }
        } else {
            const V3 wz = {w.x, w.y, g.add(w.z, qd)};
            wn = RT(l, wz);
            // wd + (0,0,qdd) + w x (0,0,qd)
            const V3 wdz = {g.sum({g.mul(w.y, qd), wd.x}), g.sum({g.neg(g.mul(w.x, qd)), wd.y}), g.add(wd.z, qdd)};
            wdn = RT(l, wdz);
            wn = v.fix(wn);
            wdn = v.fix(wdn);
            const V3 rg = RT(l, j == 0 ? in.grav : acc);
            const V3 t2 = v.cross_acc(wn, ps);
            const V3 a = v.cross_acc(wdn, ps, &rg);
            accn = v.cross_acc(wn, t2, &a);
        }
        w = v.fix(wn); wd = v.fix(wdn); acc = v.fix(accn);
        const V3 wxr = v.cross_acc(w, rc);
        const V3 a1 = v.cross_acc(wd, rc, &acc);
        const V3 abar = v.cross_acc(w, wxr, &a1);
        Fm[j] = v.fix(v.scale(Gen::cst(l.m), abar));
        const V3 Iw = v.matvec(l.I, w), Iwd = v.matvec(l.I, wd);
        Nm[j] = v.fix(v.cross_acc(w, Iw, &Iwd));
    }
    std::vector<Sym> tau(N);
    V3 f = v.zero(), nn = v.zero();
    for (int j = N - 1; j >= 0; j--) {
        const Link &l = L[j];
        const V3 rc = v.cst(l.r);
        V3 fj, nj;
        const V3 rxF = v.cross_acc(rc, Fm[j], &Nm[j]); // r x F + N
        if (mdh) {
            if (j == N - 1) {
                fj = v.add(in.ftip, Fm[j]);
                nj = v.add(in.ntip, rxF);
            } else {
                const Link &ln = L[j + 1];
                const V3 psn = ln.ps;
                const V3 Rf = v.fix(R(ln, f));
                fj = v.add(Rf, Fm[j]);
                const V3 pxf = v.cross_acc(psn, Rf, &rxF);
                nj = R(ln, nn, &pxf);
            }
        } else {
            const V3 ps = l.ps;
            if (j == N - 1) {
                fj = v.fix(v.add(in.ftip, Fm[j]));
                nj = v.cross_acc(ps, fj, &rxF, &in.ntip);
            } else {
                const Link &ln = L[j + 1];
                fj = v.fix(R(ln, f, &Fm[j]));
                const V3 pxf = v.cross_acc(ps, fj, &rxF);
                nj = R(ln, nn, &pxf);
            }
        }
        f = v.fix(fj); nn = v.fix(nj);
        // torque about (force along, for a prismatic joint) the joint axis z_{j-1} seen from frame j:
        // (0, sa, ca) for DH, (0, 0, 1) for MDH
        std::vector<Sym> t;
        const V3 &load = l.pris ? f : nn;
        if (mdh) t.push_back(load.z);
        else { t.push_back(g.mul(load.y, l.sa)); t.push_back(g.mul(load.z, l.ca)); }
        t.push_back(g.mul(Gen::cst(l.c_jm), in.qdd[j]));
        if (in.friction) t.push_back(g.mul(Gen::cst(l.c_b), in.qd[j]));
        Sym tq = g.fix(g.sum(t));
        if (in.friction && (l.c_tcp != 0.0 || l.c_tcm != 0.0) && in.qd[j].k != Sym::Z) {
            // Coulomb friction, asymmetric, none at rest (ne.c:487-490)
            const Opd qdo = g.operand(in.qd[j]);
            if (qdo.isc) {
                const double c = qdo.c > 0 ? l.c_tcp : (qdo.c < 0 ? l.c_tcm : 0.0);
                tq = g.fix(g.add(tq, Gen::cst(c)));
            } else { // tq + Tc+ [qd > 0] + Tc- [qd < 0]: two FMAs against 0 / 1 masks, no branch
                const std::string q = g.spell(qdo);
                std::string e = g.spell(g.operand(tq));
                if (l.c_tcm != 0.0) { Opd m; m.isc = true; m.c = l.c_tcm; e = "fma(" + g.spell(m) + ", step_neg(" + q + "), " + e + ")"; g.n_fma++; }
                if (l.c_tcp != 0.0) { Opd p; p.isc = true; p.c = l.c_tcp; e = "fma(" + g.spell(p) + ", step_pos(" + q + "), " + e + ")"; g.n_fma++; }
                tq = g.emit(e);
            }
        }
        tau[j] = tq;
    }
    return tau;
}

std::string assign(Gen &g, const std::string &lhs, const Sym &s)
{
    if (s.k == Sym::Z) return "    " + lhs + " = (real)0;";
    return "    " + lhs + " = " + g.spell(g.operand(s)) + ";";
}

// The operations of DynamicsMixin as evaluations of ONE recursion over symbolic inputs (`rec`: per-joint velocity and
// acceleration symbols in q order, base acceleration, tip wrench -> joint torques): shared by the DH recursion and the
// spatial-vector recursion of rigid-body trees.  Fills the statements that store the result (`tail`) and the signature.
typedef std::function<std::vector<Sym>(const Inputs &)> RecFn;
int assemble(Gen &g, int N, const b2k_gen_opts &o, const RecFn &rec, std::vector<std::string> &tail, const char *&sig, std::string &err)
{
    auto jvec = [&](const char *name) {
        std::vector<Sym> s(N);
        for (int j = 0; j < N; j++) s[j] = g.var(std::string(name) + "[" + std::to_string(j) + "]");
        return s;
    };
    auto gvec = [&]() {
        V3 v;
        v.x = (o.grav_mask & 1) ? g.var("grav[0]") : Gen::zero();
        v.y = (o.grav_mask & 2) ? g.var("grav[1]") : Gen::zero();
        v.z = (o.grav_mask & 4) ? g.var("grav[2]") : Gen::zero();
        return v;
    };
    const std::vector<Sym> zeros(N, Gen::zero());
    Vops vo(g);
    Inputs in;
    in.grav = vo.zero(); in.ftip = vo.zero(); in.ntip = vo.zero();
    if (o.mode == B2K_GEN_RNE) {
        sig = "rne_row(const real *C, const real *grav, const real *fext, const real *st, const real *ct, const real *qq, "
              "const real *qd, const real *qdd, real *out)";
        in.qd = jvec("qd"); in.qdd = jvec("qdd");
        in.grav = gvec();
        if (o.has_fext) {
            in.ftip = {g.var("fext[0]"), g.var("fext[1]"), g.var("fext[2]")};
            in.ntip = {g.var("fext[3]"), g.var("fext[4]"), g.var("fext[5]")};
        }
        std::vector<Sym> tau = rec(in);
        for (int j = 0; j < N; j++) tail.push_back(assign(g, "out[" + std::to_string(j) + "]", tau[j]));
    } else if (o.mode == B2K_GEN_GRAVLOAD) { // rne(q, 0, 0, g)  Dynamics.py:912-915
        sig = "rne_row(const real *C, const real *grav, const real *fext, const real *st, const real *ct, const real *qq, "
              "const real *in1, const real *in2, real *out)";
        in.qd = zeros; in.qdd = zeros; in.grav = gvec();
        std::vector<Sym> tau = rec(in);
        for (int j = 0; j < N; j++) tail.push_back(assign(g, "out[" + std::to_string(j) + "]", tau[j]));
    } else if (o.mode == B2K_GEN_ITORQUE) { // rne(q, 0, qdd, g = 0)  Dynamics.py:1456-1459
        sig = "rne_row(const real *C, const real *grav, const real *fext, const real *st, const real *ct, const real *qq, "
              "const real *in1, const real *in2, real *out)";
        in.qd = zeros; in.qdd = jvec("in1");
        std::vector<Sym> tau = rec(in);
        for (int j = 0; j < N; j++) tail.push_back(assign(g, "out[" + std::to_string(j) + "]", tau[j]));
    } else if (o.mode == B2K_GEN_INERTIA) { // row i of M = rne(q, 0, e_i, g = 0)  Dynamics.py:752-758
        sig = "rne_row(const real *C, const real *grav, const real *fext, const real *st, const real *ct, const real *qq, "
              "const real *in1, const real *in2, real *out)";
        for (int i = 0; i < N; i++) {
            in.qd = zeros; in.qdd = zeros;
            in.qdd[i] = Gen::cst(1.0);
            std::vector<Sym> tau = rec(in);
            for (int k = 0; k < N; k++) tail.push_back(assign(g, "out[" + std::to_string(i * N + k) + "]", tau[k]));
        }
    } else if (o.mode == B2K_GEN_CORIOLIS) { // Dynamics.py:825-857 on the friction-free robot
        sig = "rne_row(const real *C, const real *grav, const real *fext, const real *st, const real *ct, const real *qq, "
              "const real *in1, const real *in2, real *out)";
        in.friction = false;
        std::vector<Sym> qd = jvec("in1");
        std::vector<std::vector<Sym>> Csq(N, std::vector<Sym>(N));
        std::vector<std::vector<std::vector<Sym>>> terms(N, std::vector<std::vector<Sym>>(N));
        for (int i = 0; i < N; i++) { // centripetal: joint i alone at unit speed
            in.qd = zeros; in.qdd = zeros;
            in.qd[i] = Gen::cst(1.0);
            std::vector<Sym> tau = rec(in);
            for (int k = 0; k < N; k++) Csq[k][i] = g.fix(tau[k]);
        }
        const Sym half = Gen::cst(0.5);
        for (int i = 0; i < N; i++)
            for (int j = i + 1; j < N; j++) { // Coriolis: joints i and j at unit speed
                in.qd = zeros; in.qdd = zeros;
                in.qd[i] = Gen::cst(1.0); in.qd[j] = Gen::cst(1.0);
                std::vector<Sym> tau = rec(in);
                for (int k = 0; k < N; k++) {
                    const Sym t = g.fix(g.sum({tau[k], g.neg(Csq[k][j]), g.neg(Csq[k][i])}));
                    const Sym th = g.fix(g.mul(half, t));
                    terms[k][j].push_back(g.mul(th, qd[i]));
                    terms[k][i].push_back(g.mul(th, qd[j]));
                }
            }
        for (int k = 0; k < N; k++)
            for (int i = 0; i < N; i++) {
                terms[k][i].push_back(g.mul(Csq[k][i], qd[i]));
                tail.push_back(assign(g, "out[" + std::to_string(k * N + i) + "]", g.sum(terms[k][i])));
            }
    } else if (o.mode == B2K_GEN_ACCEL) {
        // tau0 = rne(q, qd, 0, g) with friction; M rows with unit accelerations, no gravity: out = [M (n*n) | torque - tau0 (n)]
        // (the kernel wrapper solves the n x n system; Dynamics.py:490-503)
        sig = "rne_row(const real *C, const real *grav, const real *fext, const real *st, const real *ct, const real *qq, "
              "const real *in1, const real *in2, real *out)";
        in.qd = jvec("in1"); in.qdd = zeros; in.grav = gvec();
        std::vector<Sym> tau0 = rec(in);
        std::vector<Sym> tq = jvec("in2");
        for (int k = 0; k < N; k++) tail.push_back(assign(g, "out[" + std::to_string(N * N + k) + "]", g.sub(tq[k], tau0[k])));
        in.grav = vo.zero();
        for (int i = 0; i < N; i++) {
            in.qd = zeros; in.qdd = zeros;
            in.qdd[i] = Gen::cst(1.0);
            std::vector<Sym> tau = rec(in);
            for (int k = 0; k < N; k++) tail.push_back(assign(g, "out[" + std::to_string(i * N + k) + "]", tau[k]));
        }
    } else {
        err = "unknown generator mode";
        return -1;
    }
    return 0;
}

} // namespace

int b2k_rne_generate(const b2k_rne_s *r, const b2k_gen_opts &o, b2k_gen_out &out)
{
    const int N = r->n;
    Gen g;
    std::vector<Link> L(N);
    for (int j = 0; j < N; j++) {
        const double *l = r->L[j];
        const double alpha = l[0], A = l[1], theta = l[2], D = l[3], G = l[20];
        const double sa = snap(sin(alpha)), ca = snap(cos(alpha));
        L[j].pris = ((int)l[4] != 0);
        L[j].sa = Gen::cst(sa);
        L[j].ca = Gen::cst(ca);
        Sym d;
        if (L[j].pris) { // the joint variable is the link offset d = q + offset (qq = q + offset in the kernel), theta is fixed
            L[j].st = Gen::cst(snap(sin(theta)));
            L[j].ct = Gen::cst(snap(cos(theta)));
            d = g.var("qq[" + std::to_string(j) + "]");
        } else {
            L[j].st = g.var("st[" + std::to_string(j) + "]");
            L[j].ct = g.var("ct[" + std::to_string(j) + "]");
            d = Gen::cst(D);
        }
        L[j].ps = {Gen::cst(A), g.fix(g.mul(d, Gen::cst(r->mdh ? -sa : sa))), g.fix(g.mul(d, Gen::cst(ca)))};
        L[j].m = l[6];
        for (int k = 0; k < 3; k++) L[j].r[k] = l[7 + k];
        for (int k = 0; k < 9; k++) L[j].I[k] = l[10 + k];
        L[j].c_jm = G * G * l[19];
        L[j].c_b = G * G * l[21];
        L[j].c_tcp = fabs(G) * l[22];
        L[j].c_tcm = fabs(G) * l[23];
    }
    std::vector<std::string> tail;
    const char *sig = nullptr;
    if (assemble(g, N, o, [&](const Inputs &in) { return recursion(g, L, r->mdh != 0, in); }, tail, sig, out.error)) return -1;
    std::string src = std::string("__device__ __forceinline__ void ") + sig + "\n{\n";
    for (const std::string &s : g.code) src += s + "\n";
    for (const std::string &s : tail) src += s + "\n";
    src += "}\n";
    out.source = src;
    out.consts = g.consts;
    if (out.consts.empty()) out.consts.push_back(0.0);
    out.n_mul = g.n_mul; out.n_fma = g.n_fma; out.n_add = g.n_add;
    out.error.clear();
    return 0;
}

// ------------------------------------------------------------------ rigid-body trees (Robot.rne)
// Featherstone's recursion as the reference's Robot.rne runs it (Robot.py:1704-1903) with spatial vectors in
// [linear; angular] order (spatialmath SpatialVelocity / SpatialForce): per joint group j with T_j = C_j J_j(q),
//   v_j = X_j v_p + s_j qd_j            a_j = X_j a_p + s_j qdd_j + v_j x s_j qd_j      (root: a_p = -gravity)
//   f_j = I_j a_j + v_j x* (I_j v_j)    tau_j = s_j . f_j      f_p += X_j^T f_j,        X_j = Ad(T_j^-1)
// written out on 3-vectors and evaluated symbolically: constant rotations with 0 / +-1 entries become permutations,
// zero inertia entries vanish, the root's velocities are one-component.  Reference quirks kept: the motion subspace
// s_j ignores a joint's flip (ET.s, ET.py:592-608), the group inertia is the plain sum of SpatialInertia(m, r) of its
// links without their rotational inertia, torques come out in group order.
namespace {
struct TreeOps {
    Gen &g;
    Vops v;
    explicit TreeOps(Gen &gg) : g(gg), v(gg) {}
    static double snap01(double x)
    {
        const double r = nearbyint(x);
        return (fabs(x - r) < 4e-16 && fabs(r) <= 1.0) ? r : x;
    }
    V3 constR(const double *C, const V3 &u, bool transpose)
    { // Rc u or Rc^T u, Rc = the rotation part of the 3x4 constant
        V3 a = v.fix(u);
        auto e = [&](int i, int j) { return Gen::cst(snap01(transpose ? C[j * 4 + i] : C[i * 4 + j])); };
        auto row = [&](int i) { return g.sum({g.mul(e(i, 0), a.x), g.mul(e(i, 1), a.y), g.mul(e(i, 2), a.z)}); };
        return {row(0), row(1), row(2)};
    }
    V3 jointR(int axis, const Sym &s, const Sym &c, const V3 &u, bool transpose)
    {
        if (axis > 2) return u; // prismatic: no rotation
        V3 a = v.fix(u);
        const Sym sx = transpose ? s : g.neg(s);
        // rotation about axis k mixes the other two components (cyclic order)
        Sym *comp[3] = {&a.x, &a.y, &a.z};
        const int i = (axis + 1) % 3, j = (axis + 2) % 3;
        V3 r = a;
        Sym *out[3] = {&r.x, &r.y, &r.z};
        *out[i] = g.sum({g.mul(c, *comp[i]), g.mul(sx, *comp[j])});
        *out[j] = g.sum({g.mul(g.neg(sx), *comp[i]), g.mul(c, *comp[j])});
        return r;
    }
    V3 mat3(const double *I6, int r0, int c0, const V3 &u)
    { // 3x3 block (rows r0.., cols c0..) of the 6x6 constant times u
        V3 a = v.fix(u);
        auto row = [&](int i) {
            return g.sum({g.mul(Gen::cst(I6[(r0 + i) * 6 + c0]), a.x), g.mul(Gen::cst(I6[(r0 + i) * 6 + c0 + 1]), a.y),
                          g.mul(Gen::cst(I6[(r0 + i) * 6 + c0 + 2]), a.z)});
        };
        return {row(0), row(1), row(2)};
    }
};
} // namespace

int b2k_tree_generate(const b2k_tree_s *t, const b2k_gen_opts &o, b2k_gen_out &out)
{
    const int N = t->n;
    Gen g;
    TreeOps T(g);
    Vops &v = T.v;
    std::vector<Sym> st(N), ct(N), qq(N);
    for (int j = 0; j < N; j++) {
        const std::string k = "[" + std::to_string(t->jindex[j]) + "]";
        st[j] = g.var("st" + k); ct[j] = g.var("ct" + k); qq[j] = g.var("qq" + k);
        if (t->flip[j]) { st[j] = g.neg(st[j]); qq[j] = g.neg(qq[j]); } // eta = -q
    }
    auto unit = [&](int axis, const Sym &x) { // x * e_axis
        V3 u = v.zero();
        (axis % 3 == 0 ? u.x : (axis % 3 == 1 ? u.y : u.z)) = x;
        return u;
    };
    // one evaluation of the recursion: in.qd / in.qdd are indexed like q (jindex), the torques come out in group order
    // (Robot.py:1890-1899); in.grav is a_grav = MINUS the gravity vector (the caller negates), linear part only
    auto rec = [&](const Inputs &in) {
        const V3 &agrav = in.grav;
        std::vector<V3> w(N), vl(N), fl(N), fa(N), p(N);
        std::vector<V3> aw(N), av(N);
        for (int j = 0; j < N; j++) {
            const double *C = t->C[j];
            const int ax = t->axis[j];
            const bool rev = ax < 3;
            const Sym &qdj = in.qd[t->jindex[j]], &qddj = in.qdd[t->jindex[j]];
            auto RT = [&](const V3 &u) { return T.jointR(ax, st[j], ct[j], T.constR(C, u, true), true); };
            // p = pc + Rc (e_axis * eta) for a prismatic joint
            const double pc[3] = {C[3], C[7], C[11]};
            p[j] = v.cst(pc);
            if (!rev) p[j] = v.fix(v.add(p[j], T.constR(C, unit(ax, qq[j]), false)));
            const V3 vJw = rev ? unit(ax, qdj) : v.zero(), vJv = rev ? v.zero() : unit(ax, qdj);
            const V3 aJw = rev ? unit(ax, qddj) : v.zero(), aJv = rev ? v.zero() : unit(ax, qddj);
            const int pa = t->parent[j];
            if (pa < 0) {
                w[j] = vJw; vl[j] = vJv;
                aw[j] = aJw;
                av[j] = v.fix(v.add(RT(agrav), aJv));
            } else {
                w[j] = v.fix(v.add(RT(w[pa]), vJw));
                vl[j] = v.fix(v.add(RT(v.cross_acc(w[pa], p[j], &vl[pa])), vJv));
                const V3 wxJw = v.cross_acc(w[j], vJw);
                aw[j] = v.fix(v.add(v.add(RT(aw[pa]), aJw), wxJw));
                const V3 t1 = RT(v.cross_acc(aw[pa], p[j], &av[pa]));
                const V3 t2 = v.cross_acc(w[j], vJv, &aJv);
                const V3 t3 = v.cross_acc(vl[j], vJw, &t2);
                av[j] = v.fix(v.add(t1, t3));
            }
            const double *I6 = t->I6[j];
            const V3 hl = v.fix(v.add(T.mat3(I6, 0, 0, vl[j]), T.mat3(I6, 0, 3, w[j])));
            const V3 ha = v.fix(v.add(T.mat3(I6, 3, 0, vl[j]), T.mat3(I6, 3, 3, w[j])));
            const V3 il = v.add(T.mat3(I6, 0, 0, av[j]), T.mat3(I6, 0, 3, aw[j]));
            const V3 ia = v.add(T.mat3(I6, 3, 0, av[j]), T.mat3(I6, 3, 3, aw[j]));
            fl[j] = v.fix(v.cross_acc(w[j], hl, &il));
            const V3 t4 = v.cross_acc(vl[j], hl, &ia);
            fa[j] = v.fix(v.cross_acc(w[j], ha, &t4));
        }
        std::vector<Sym> tau(N);
        for (int j = N - 1; j >= 0; j--) {
            const int ax = t->axis[j];
            const V3 &f = ax < 3 ? fa[j] : fl[j];
            tau[j] = g.fix(ax % 3 == 0 ? f.x : (ax % 3 == 1 ? f.y : f.z));
            const int pa = t->parent[j];
            if (pa >= 0) {
                const double *C = t->C[j];
                auto R = [&](const V3 &u) { return T.constR(C, T.jointR(ax, st[j], ct[j], u, false), false); };
                const V3 Rf = v.fix(R(fl[j]));
                fl[pa] = v.fix(v.add(fl[pa], Rf));
                const V3 Rn = R(fa[j]);
                const V3 pxf = v.cross_acc(p[j], Rf, &Rn);
                fa[pa] = v.fix(v.add(fa[pa], pxf));
            }
        }
        return tau;
    };
    std::vector<std::string> tail;
    const char *sig = nullptr;
    b2k_gen_opts oo = o;
    oo.has_fext = 0; // Robot.rne has no tip wrench
    if (assemble(g, N, oo, rec, tail, sig, out.error)) return -1;
    std::string src = std::string("__device__ __forceinline__ void ") + sig + "\n{\n";
    for (const std::string &s : g.code) src += s + "\n";
    for (const std::string &s : tail) src += s + "\n";
    src += "}\n";
    out.source = src;
    out.consts = g.consts;
    if (out.consts.empty()) out.consts.push_back(0.0);
    out.n_mul = g.n_mul; out.n_fma = g.n_fma; out.n_add = g.n_add;
    out.error.clear();
    return 0;
}
