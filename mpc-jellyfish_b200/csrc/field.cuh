// field.cuh -- Montgomery-form prime-field arithmetic for sm_100a (K1 of SURVEY.md §2.3).
//
// Replaces ark-ff 0.4 `Fp<MontBackend<_, N>>` (un-vendored dependency of the reference,
// plonk/Cargo.toml:15) for the four fields on the hot path.  Memory layout is ark-ff's:
// little-endian limbs of a*R mod p, R = 2^(32*N32) = 2^(64*N64), always fully reduced,
// so a `&[Fr]` can be handed to the GPU without conversion.
//
// Arithmetic runs on the integer-multiply (IMAD) pipe with 32-bit limbs.  The multiplier is
// an operand-scanning Montgomery product over two interleaved accumulators (even / odd
// columns) so that every 32x32->64 partial product is one IMAD.WIDE.U32 with carry-in/out
// and the per-row shift is a register renaming, not a move.  The chains themselves live in
// the generated mont_chains.cuh.  Everything is also compilable for the host (the chains
// carry a portable C fallback) which is how the algorithm is unit-tested without a GPU and
// how the library does its few serial host-side field operations.
#pragma once
#include <stdint.h>

#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#endif
#include "mont_chains.cuh"

#define JF_HD __host__ __device__ __forceinline__

namespace jf {

// ---- field descriptors ---------------------------------------------------------------------
// P* modulus limbs, R* = R mod p (Montgomery one), RR* = R^2 mod p, INV = -p^-1 mod 2^32.
#define JF_LIMB_DECL8(pfx, a0, a1, a2, a3, a4, a5, a6, a7)                                      \
    static constexpr uint32_t pfx##0 = a0, pfx##1 = a1, pfx##2 = a2, pfx##3 = a3, pfx##4 = a4,  \
                              pfx##5 = a5, pfx##6 = a6, pfx##7 = a7;
#define JF_LIMB_DECL12(pfx, a0, a1, a2, a3, a4, a5, a6, a7, a8, a9, a10, a11)                   \
    JF_LIMB_DECL8(pfx, a0, a1, a2, a3, a4, a5, a6, a7)                                          \
    static constexpr uint32_t pfx##8 = a8, pfx##9 = a9, pfx##10 = a10, pfx##11 = a11;

#define JF_ARR8(F, pfx) {F::pfx##0, F::pfx##1, F::pfx##2, F::pfx##3, F::pfx##4, F::pfx##5, F::pfx##6, F::pfx##7}
#define JF_ARR12(F, pfx)                                                                         \
    {F::pfx##0, F::pfx##1, F::pfx##2, F::pfx##3, F::pfx##4, F::pfx##5, F::pfx##6, F::pfx##7,     \
     F::pfx##8, F::pfx##9, F::pfx##10, F::pfx##11}

struct Bn254Fr {
    static constexpr bool P_OPAQUE = false;
    static constexpr int N = 8;
    static constexpr int BITS = 254;
    static constexpr int TWO_ADICITY = 28;
    static constexpr uint32_t GENERATOR = 5;
    static constexpr uint32_t INV = 0xefffffffu;
    JF_LIMB_DECL8(P, 0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u)
    JF_LIMB_DECL8(R, 0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u)
    JF_LIMB_DECL8(RR, 0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u)
};

struct Bn254Fq {
    static constexpr bool P_OPAQUE = false;
    static constexpr int N = 8;
    static constexpr int BITS = 254;
    static constexpr uint32_t INV = 0xe4866389u;
    JF_LIMB_DECL8(P, 0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u)
    JF_LIMB_DECL8(R, 0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u)
    JF_LIMB_DECL8(RR, 0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u, 0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u)
};

struct Bls12381Fr {
    static constexpr bool P_OPAQUE = true;  // INV = 0xffffffff must stay opaque to ptxas: see mont_inv()
    static constexpr int N = 8;
    static constexpr int BITS = 255;
    static constexpr int TWO_ADICITY = 32;
    static constexpr uint32_t GENERATOR = 7;
    static constexpr uint32_t INV = 0xffffffffu;
    JF_LIMB_DECL8(P, 0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u)
    JF_LIMB_DECL8(R, 0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau, 0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u)
    JF_LIMB_DECL8(RR, 0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu, 0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u)
};

struct Bls12381Fq {
    static constexpr bool P_OPAQUE = false;
    static constexpr int N = 12;
    static constexpr int BITS = 381;
    static constexpr uint32_t INV = 0xfffcfffdu;
    JF_LIMB_DECL12(P, 0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u, 0xf38512bfu, 0x64774b84u,
                   0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau)
    JF_LIMB_DECL12(R, 0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu, 0x53c758bau, 0x5f489857u, 0x70525745u, 0x77ce5853u,
                   0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u)
    JF_LIMB_DECL12(RR, 0x1c341746u, 0xf4df1f34u, 0x09d104f1u, 0x0a76e6a6u, 0x4c95b6d5u, 0x8de5476cu, 0x939d83c0u, 0x67eb88a9u,
                   0xb519952du, 0x9a793e85u, 0x92cae3aau, 0x11988fe5u)
};

template <class F> struct FieldConsts;  // P / R / RR as arrays

template <class F, int N = F::N> struct Limbs;
template <class F> struct Limbs<F, 8> {
    static JF_HD void p(uint32_t (&o)[8]) { const uint32_t t[8] = JF_ARR8(F, P); for (int i = 0; i < 8; i++) o[i] = t[i]; }
    static JF_HD void r(uint32_t (&o)[8]) { const uint32_t t[8] = JF_ARR8(F, R); for (int i = 0; i < 8; i++) o[i] = t[i]; }
    static JF_HD void rr(uint32_t (&o)[8]) { const uint32_t t[8] = JF_ARR8(F, RR); for (int i = 0; i < 8; i++) o[i] = t[i]; }
    static constexpr uint32_t top() { return F::P7; }
};
template <class F> struct Limbs<F, 12> {
    static JF_HD void p(uint32_t (&o)[12]) { const uint32_t t[12] = JF_ARR12(F, P); for (int i = 0; i < 12; i++) o[i] = t[i]; }
    static JF_HD void r(uint32_t (&o)[12]) { const uint32_t t[12] = JF_ARR12(F, R); for (int i = 0; i < 12; i++) o[i] = t[i]; }
    static JF_HD void rr(uint32_t (&o)[12]) { const uint32_t t[12] = JF_ARR12(F, RR); for (int i = 0; i < 12; i++) o[i] = t[i]; }
    static constexpr uint32_t top() { return F::P11; }
};

#ifdef __CUDACC__
static __constant__ uint32_t jf_c_bls12381_fr_inv = Bls12381Fr::INV;
#endif
// -p^-1 mod 2^32.  For Bls12381Fr it is 0xffffffff: as an immediate ptxas rewrites m = x * INV into a
// negation, pushes the sign through the m * p products and no longer fuses their (lo, hi) pairs into
// IMAD.WIDE.  Reading the constant from constant memory keeps m an ordinary product.
template <class F> JF_HD uint32_t mont_inv() {
#ifdef __CUDA_ARCH__
    if constexpr (F::P_OPAQUE) return jf_c_bls12381_fr_inv;
#endif
    return F::INV;
}
template <class F> JF_HD void mad_p_pair(uint32_t (&odd_acc)[F::N], uint32_t (&even_acc)[F::N], uint32_t m) {
    chain_mad_odd_p<F>(odd_acc, m);
    chain_mad_even_p<F>(even_acc, m, odd_acc[F::N - 1]);
}

// ---- the field element -------------------------------------------------------------------
template <class F> struct Fp {
    static constexpr int N = F::N;
    uint32_t v[N];

    static JF_HD Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = 0;
        return r;
    }
    static JF_HD Fp one() {
        Fp r;
        Limbs<F>::r(r.v);
        return r;
    }
    static JF_HD Fp r_squared() {
        Fp r;
        Limbs<F>::rr(r.v);
        return r;
    }
    static JF_HD Fp from_u32(uint32_t x) {  // small integer -> Montgomery form
        Fp t = zero();
        t.v[0] = x;
        return mul(t, r_squared());
    }

    JF_HD bool is_zero() const {
        uint32_t t = 0;
#pragma unroll
        for (int i = 0; i < N; i++) t |= v[i];
        return t == 0;
    }
    JF_HD bool operator==(const Fp &o) const {
        uint32_t t = 0;
#pragma unroll
        for (int i = 0; i < N; i++) t |= v[i] ^ o.v[i];
        return t == 0;
    }
    JF_HD bool operator!=(const Fp &o) const { return !(*this == o); }

    // r = a if r < p else r - p   (input < 2p)
    static JF_HD void reduce_once(Fp &a) {
        uint32_t t[N], borrow;
        chain_sub_p<F>(t, a.v, borrow);
#pragma unroll
        for (int i = 0; i < N; i++) a.v[i] = borrow ? a.v[i] : t[i];
    }

    static JF_HD Fp add(const Fp &a, const Fp &b) {
        Fp r;
        uint32_t c;
        chain_add(r.v, a.v, b.v, c);  // 2p < 2^(32N): no carry out
        reduce_once(r);
        return r;
    }
    static JF_HD Fp sub(const Fp &a, const Fp &b) {
        Fp r;
        uint32_t borrow, pm[N], t[N];
        chain_sub(t, a.v, b.v, borrow);
        Limbs<F>::p(pm);
#pragma unroll
        for (int i = 0; i < N; i++) pm[i] &= borrow;
        chain_add_masked(r.v, t, pm);
        return r;
    }
    static JF_HD Fp neg(const Fp &a) {
        if (a.is_zero()) return a;
        Fp r, pp;
        uint32_t borrow;
        Limbs<F>::p(pp.v);
        chain_sub(r.v, pp.v, a.v, borrow);
        return r;
    }
    static JF_HD Fp dbl(const Fp &a) { return add(a, a); }

    // Montgomery product a*b/R mod p.  See DESIGN.md "K1" for the even/odd accumulator scheme.
    static JF_HD Fp mul(const Fp &a, const Fp &b) {
        uint32_t x[N], y[N];
        // row 0: plain wide products, no carries between the pairs
#pragma unroll
        for (int j = 0; j < N; j += 2) {
            uint64_t e = (uint64_t)a.v[j] * b.v[0];
            uint64_t o = (uint64_t)a.v[j + 1] * b.v[0];
            x[j] = (uint32_t)e;
            x[j + 1] = (uint32_t)(e >> 32);
            y[j] = (uint32_t)o;
            y[j + 1] = (uint32_t)(o >> 32);
        }
        {
            uint32_t m = x[0] * mont_inv<F>();
            mad_p_pair<F>(y, x, m);
        }
#pragma unroll
        for (int i = 1; i < N; i += 2) {
            row(x, y, a.v, b.v[i]);                 // even acc = y, odd acc = x
            if (i + 1 < N) row(y, x, a.v, b.v[i + 1]);  // even acc = x, odd acc = y
        }
        // N even: after the last (odd-numbered) row the even accumulator is y, the odd one is x
        Fp r;
        chain_merge(r.v, y, x);
        reduce_once(r);
        return r;
    }
    // Montgomery square.  Row i of the operand-scanning product only needs the terms with j >= i when the others were
    // added doubled by the rows before it: row i multiplies a_i with d = (a_i, 2 a^{>i}) limb by limb, 36 (78 for 12 limbs)
    // wide products instead of 64 (144); the skipped ones become carry propagation (chain_*_from<I>, gen_chains.py).
    // The doubled operand costs one bit of headroom in the top limb, so this needs p < 2^(32 N - 2) (both Fq and BN254
    // Fr; BLS12-381 Fr, 255 bits, falls back to mul).  Enabled by JF_DEDICATED_SQR (the Makefile sets it): bit-exact on the
    // host (tests/test_host_field.py) and on the device (tests/test_gpu_field.py); 2^20 MSM 3.24 -> 3.17 ms.
    static constexpr bool SQR_OK = Limbs<F>::top() < (1u << 30);
    static JF_HD Fp sqr(const Fp &a) {
#ifdef JF_DEDICATED_SQR
        if constexpr (SQR_OK) return sqr_dedicated(a);
#endif
        return mul(a, a);
    }
    // a b + c d with ONE Montgomery reduction: every row adds both partial products before its reduction row, 2 x 64 + 72
    // wide products instead of 2 x (64 + 72).  (a b + c d + M p) / R < 1.5 p for p < R / 4, so one conditional subtraction
    // still fully reduces; like the squaring this needs the two spare top bits (SQR_OK; tests/test_field_model.py).
    // y3 = r (q - x3) - y ppp of the point formulas has this shape (ec.cuh, behind JF_FUSED_MULSUB: host-verified, not yet
    // run on a GPU, so the Makefile leaves it off).
    static JF_HD Fp mul_add(const Fp &a, const Fp &b, const Fp &c, const Fp &d) {
        if constexpr (!SQR_OK) return add(mul(a, b), mul(c, d));
        uint32_t x[N], y[N];
#pragma unroll
        for (int j = 0; j < N; j += 2) {
            uint64_t e = (uint64_t)a.v[j] * b.v[0];
            uint64_t o = (uint64_t)a.v[j + 1] * b.v[0];
            x[j] = (uint32_t)e;
            x[j + 1] = (uint32_t)(e >> 32);
            y[j] = (uint32_t)o;
            y[j + 1] = (uint32_t)(o >> 32);
        }
        chain_mad_odd(y, c.v, d.v[0]);
        chain_mad_even(x, c.v, d.v[0], y[N - 1]);
        {
            uint32_t m = x[0] * mont_inv<F>();
            mad_p_pair<F>(y, x, m);
        }
#pragma unroll
        for (int i = 1; i < N; i += 2) {
            row2(x, y, a.v, b.v[i], c.v, d.v[i]);
            if (i + 1 < N) row2(y, x, a.v, b.v[i + 1], c.v, d.v[i + 1]);
        }
        Fp r;
        chain_merge(r.v, y, x);
        reduce_once(r);
        return r;
    }
    // a b - c d
    static JF_HD Fp mul_sub(const Fp &a, const Fp &b, const Fp &c, const Fp &d) {
#ifdef JF_FUSED_MULSUB
        if constexpr (SQR_OK) return mul_add(a, b, c, neg(d));
#endif
        return sub(mul(a, b), mul(c, d));
    }
    static JF_HD void row2(uint32_t (&prev_e)[N], uint32_t (&prev_o)[N], const uint32_t (&a)[N], uint32_t bi, const uint32_t (&c)[N],
                           uint32_t di) {
        chain_shift_mad_odd(prev_e, prev_o[0], a, bi);
        chain_mad_even(prev_o, a, bi, prev_e[N - 1]);
        chain_mad_odd(prev_e, c, di);
        chain_mad_even(prev_o, c, di, prev_e[N - 1]);
        uint32_t m = prev_o[0] * mont_inv<F>();
        mad_p_pair<F>(prev_e, prev_o, m);
    }
    template <int I> static JF_HD void sqr_rows(uint32_t (&even)[N], uint32_t (&odd)[N], const uint32_t (&a)[N], const uint32_t (&a2)[N]) {
        if constexpr (I < N) {
            uint32_t d[N];
#pragma unroll
            for (int k = 0; k < N; k++) d[k] = a2[k];
            d[I] = a[I];
            if constexpr (I + 1 < N) d[I + 1] = a[I + 1] << 1;
            // `even` held the even columns (limb 0 zero after the previous reduction), `odd` the odd ones
            chain_shift_mad_odd_from<I>(even, odd[0], d, a[I]);
            chain_mad_even_from<I>(odd, d, a[I], even[N - 1]);
            uint32_t m = odd[0] * mont_inv<F>();
            mad_p_pair<F>(even, odd, m);
            sqr_rows<I + 1>(odd, even, a, a2);  // the accumulators swap roles
        }
    }
    static JF_HD Fp sqr_dedicated(const Fp &a) {
        uint32_t x[N], y[N], a2[N], d[N];
        a2[0] = a.v[0] << 1;
#pragma unroll
        for (int k = 1; k < N; k++) a2[k] = (a.v[k] << 1) | (a.v[k - 1] >> 31);
#pragma unroll
        for (int k = 0; k < N; k++) d[k] = a2[k];
        d[0] = a.v[0];
        d[1] = a.v[1] << 1;
        // row 0: plain wide products, no carries between the pairs
#pragma unroll
        for (int j = 0; j < N; j += 2) {
            uint64_t e = (uint64_t)d[j] * a.v[0];
            uint64_t o = (uint64_t)d[j + 1] * a.v[0];
            x[j] = (uint32_t)e;
            x[j + 1] = (uint32_t)(e >> 32);
            y[j] = (uint32_t)o;
            y[j + 1] = (uint32_t)(o >> 32);
        }
        {
            uint32_t m = x[0] * mont_inv<F>();
            mad_p_pair<F>(y, x, m);
        }
        sqr_rows<1>(x, y, a.v, a2);  // row 1: even acc = x? see mul(): row(x, y, ..) first
        // N even: after the last (odd-numbered) row the even accumulator is y, the odd one is x
        Fp r;
        chain_merge(r.v, y, x);
        reduce_once(r);
        return r;
    }

    // One operand-scanning row: `prev_e` held the even columns of the running sum (its limb 0
    // is zero after the previous reduction), `prev_o` the odd ones.  After the call prev_o has
    // become the even accumulator and prev_e (shifted down 64 bits) the odd one.
    static JF_HD void row(uint32_t (&prev_e)[N], uint32_t (&prev_o)[N], const uint32_t (&a)[N], uint32_t bi) {
        chain_shift_mad_odd(prev_e, prev_o[0], a, bi);
        chain_mad_even(prev_o, a, bi, prev_e[N - 1]);
        uint32_t m = prev_o[0] * mont_inv<F>();
        mad_p_pair<F>(prev_e, prev_o, m);
    }

    static JF_HD Fp to_mont(const Fp &a) { return mul(a, r_squared()); }
    static JF_HD Fp from_mont(const Fp &a) {
        Fp o = zero();
        o.v[0] = 1;
        return mul(a, o);
    }

    // a^e, e a plain little-endian integer of `words` 32-bit words
    static JF_HD Fp pow(const Fp &a, const uint32_t *e, int words) {
        Fp acc = one();
        for (int i = words * 32 - 1; i >= 0; i--) {
            acc = sqr(acc);
            if ((e[i >> 5] >> (i & 31)) & 1) acc = mul(acc, a);
        }
        return acc;
    }
    static JF_HD Fp pow_u64(const Fp &a, uint64_t e) {
        uint32_t w[2] = {(uint32_t)e, (uint32_t)(e >> 32)};
        return pow(a, w, 2);
    }
    // Fermat inverse a^(p-2); inv(0) = 0
    static JF_HD Fp inv(const Fp &a) {
        uint32_t e[N];
        Limbs<F>::p(e);
        uint32_t borrow = 2;
        for (int i = 0; i < N && borrow; i++) {
            uint32_t t = e[i];
            e[i] = t - borrow;
            borrow = t < borrow ? 1 : 0;
        }
        return pow(a, e, N);
    }

    JF_HD Fp operator+(const Fp &o) const { return add(*this, o); }
    JF_HD Fp operator-(const Fp &o) const { return sub(*this, o); }
    JF_HD Fp operator*(const Fp &o) const { return mul(*this, o); }
};

}  // namespace jf
