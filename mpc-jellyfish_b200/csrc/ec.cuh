// ec.cuh -- G1 group law for y^2 = x^3 + b (a = 0) in extended Jacobian "XYZZ" coordinates.
//
// Replaces the curve arithmetic ark-ec 0.4 runs inside `VariableBaseMSM::msm_bigint`
// (called at primitives/src/pcs/univariate_kzg/mod.rs:110,151).  A point is (X, Y, ZZ, ZZZ)
// with x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; identity <=> ZZ = 0.  The mixed addition
// (XYZZ += affine) costs 8M + 2S and needs no inversion, which is what the bucket
// accumulation wants; formulas: EFD "xyzz" add-2008-s, madd-2008-s, dbl-2008-s-1, mdbl-2008-s-1.
// Affine points are x || y with identity encoded as (0, 0) ((0,0) is not on either curve).
#pragma once
#include "field.cuh"

namespace jf {

template <class Fq> struct Affine {
    Fp<Fq> x, y;
    JF_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
    static JF_HD Affine inf() { return Affine{Fp<Fq>::zero(), Fp<Fq>::zero()}; }
};

template <class Fq> struct XYZZ {
    using F = Fp<Fq>;
    F x, y, zz, zzz;

    JF_HD bool is_inf() const { return zz.is_zero(); }
    static JF_HD XYZZ inf() { return XYZZ{F::zero(), F::zero(), F::zero(), F::zero()}; }
    static JF_HD XYZZ from_affine(const Affine<Fq> &p) {
        if (p.is_inf()) return inf();
        return XYZZ{p.x, p.y, F::one(), F::one()};
    }

    // 2 * (affine p), p != identity      (mdbl-2008-s-1)
    static JF_HD XYZZ dbl_affine(const Affine<Fq> &p) {
        XYZZ r;
        F u = F::dbl(p.y);
        F v = F::sqr(u);
        F w = F::mul(u, v);
        F s = F::mul(p.x, v);
        F xx = F::sqr(p.x);
        F m = F::add(F::dbl(xx), xx);
        r.x = F::sub(F::sqr(m), F::dbl(s));
        r.y = F::mul_sub(m, F::sub(s, r.x), w, p.y);
        r.zz = v;
        r.zzz = w;
        return r;
    }

    // 2 * this      (dbl-2008-s-1)
    JF_HD XYZZ dbl() const {
        if (is_inf()) return *this;
        XYZZ r;
        F u = F::dbl(y);
        F v = F::sqr(u);
        F w = F::mul(u, v);
        F s = F::mul(x, v);
        F xx = F::sqr(x);
        F m = F::add(F::dbl(xx), xx);
        r.x = F::sub(F::sqr(m), F::dbl(s));
        r.y = F::mul_sub(m, F::sub(s, r.x), w, y);
        r.zz = F::mul(v, zz);
        r.zzz = F::mul(w, zzz);
        return r;
    }

    // this += affine q (q != identity).  Handles this == identity, this == q, this == -q.
    JF_HD void add_affine(const Affine<Fq> &q) {
        if (is_inf()) {
            x = q.x; y = q.y; zz = F::one(); zzz = F::one();
            return;
        }
        F u2 = F::mul(q.x, zz);
        F s2 = F::mul(q.y, zzz);
        F p = F::sub(u2, x);
        F r = F::sub(s2, y);
        if (p.is_zero()) {
            if (r.is_zero()) *this = dbl_affine(q);
            else *this = inf();
            return;
        }
        F pp = F::sqr(p);
        F ppp = F::mul(p, pp);
        F q1 = F::mul(x, pp);
        F x3 = F::sub(F::sub(F::sqr(r), ppp), F::dbl(q1));
        F y3 = F::mul_sub(r, F::sub(q1, x3), y, ppp);
        x = x3;
        y = y3;
        zz = F::mul(zz, pp);
        zzz = F::mul(zzz, ppp);
    }

    // this += o   (add-2008-s)
    JF_HD void add(const XYZZ &o) {
        if (o.is_inf()) return;
        if (is_inf()) { *this = o; return; }
        F u1 = F::mul(x, o.zz);
        F u2 = F::mul(o.x, zz);
        F s1 = F::mul(y, o.zzz);
        F s2 = F::mul(o.y, zzz);
        F p = F::sub(u2, u1);
        F r = F::sub(s2, s1);
        if (p.is_zero()) {
            if (r.is_zero()) *this = dbl();
            else *this = inf();
            return;
        }
        F pp = F::sqr(p);
        F ppp = F::mul(p, pp);
        F q1 = F::mul(u1, pp);
        F x3 = F::sub(F::sub(F::sqr(r), ppp), F::dbl(q1));
        F y3 = F::mul_sub(r, F::sub(q1, x3), s1, ppp);
        x = x3;
        y = y3;
        zz = F::mul(F::mul(zz, o.zz), pp);
        zzz = F::mul(F::mul(zzz, o.zzz), ppp);
    }

    JF_HD XYZZ neg() const {
        XYZZ r = *this;
        r.y = F::neg(y);
        return r;
    }

    // `into_affine`: identity -> (0, 0)
    JF_HD Affine<Fq> to_affine() const {
        if (is_inf()) return Affine<Fq>::inf();
        F zi = F::inv(zzz);  // 1/ZZZ;  ZZ^3 = ZZZ^2  =>  1/ZZ = ZZ^2 / ZZZ^2 = (ZZ * zi)^2
        F t = F::mul(zi, zz);
        F izz = F::sqr(t);
        return Affine<Fq>{F::mul(x, izz), F::mul(y, zi)};
    }
};

// scalar * affine by double-and-add (small / one-off uses: SRS generation, tests)
template <class Fq> JF_HD XYZZ<Fq> scalar_mul(const Affine<Fq> &p, const uint32_t *k, int words) {
    XYZZ<Fq> acc = XYZZ<Fq>::inf();
    if (p.is_inf()) return acc;
    for (int i = words * 32 - 1; i >= 0; i--) {
        acc = acc.dbl();
        if ((k[i >> 5] >> (i & 31)) & 1) acc.add_affine(p);
    }
    return acc;
}

// ---- curve descriptors ------------------------------------------------------------------
struct Bn254G1 {
    using Fq = Bn254Fq;
    using Fr = Bn254Fr;
    static JF_HD Affine<Fq> generator() {  // (1, 2)
        return Affine<Fq>{Fp<Fq>::one(), Fp<Fq>::dbl(Fp<Fq>::one())};
    }
};

struct Bls12381G1 {
    using Fq = Bls12381Fq;
    using Fr = Bls12381Fr;
    static JF_HD Affine<Fq> generator() {
        Affine<Fq> g;
        const uint32_t gx[12] = {0xdb22c6bbu, 0xfb3af00au, 0xf97a1aefu, 0x6c55e83fu, 0x171bac58u, 0xa14e3a3fu,
                                 0x9774b905u, 0xc3688c4fu, 0x4fa9ac0fu, 0x2695638cu, 0x3197d794u, 0x17f1d3a7u};
        const uint32_t gy[12] = {0x46c5e7e1u, 0x0caa2329u, 0xa2888ae4u, 0xd03cc744u, 0x2c04b3edu, 0x00db18cbu,
                                 0xd5d00af6u, 0xfcf5e095u, 0x741d8ae4u, 0xa09e30edu, 0xe3aaa0f1u, 0x08b3f481u};
        for (int i = 0; i < 12; i++) { g.x.v[i] = gx[i]; g.y.v[i] = gy[i]; }
        g.x = Fp<Fq>::to_mont(g.x);
        g.y = Fp<Fq>::to_mont(g.y);
        return g;
    }
};

}  // namespace jf
