/* jf_oracle.c -- CPU restatement of the KZG-commit MSM and radix-2 NTT hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product never does.
 *
 * PARITY UNPINNED by the reference: mpc-jellyfish has no golden vectors for MSM / NTT and
 * its arithmetic lives in un-vendored crates (ark-ff / ark-ec 0.4.2 / ark-poly 0.4.2,
 * ark-bn254 / ark-bls12-381 0.4.0; plonk/Cargo.toml:13-19, primitives/Cargo.toml:12-28) that
 * cannot be built here (no Rust).  This file restates their published algorithms and is
 * itself checked against oracle/pyref.py (exact big-int) and the public constants /
 * known-answer points in tests/golden/.
 *
 * What follows which reference call site:
 *   jfo_msm        E::G1::msm_bigint(..).into_affine()   primitives/src/pcs/univariate_kzg/mod.rs:108-111,151-155
 *   jfo_ntt        Radix2EvaluationDomain::{fft,ifft}_in_place and get_coset(GENERATOR)
 *                  relation/src/constraint_system.rs:1172,1189,1221,1240,1257;
 *                  plonk/src/proof_system/prover.rs:545,552-567,672
 *   jfo_gen_srs    gen_srs_for_testing                   primitives/src/pcs/univariate_kzg/srs.rs:118-153
 *   jfo_field_op   ark-ff Fp arithmetic, into_bigint     primitives/src/pcs/univariate_kzg/mod.rs:390-395
 *
 * Memory layout = ark-ff's: little-endian u64 limbs, Montgomery form unless stated.
 * Affine point = x || y, identity = (0, 0).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

#define NL 4
#define FN(x) CAT(f4_, x)
#include "mont_tmpl.h"
#define EC(x) CAT(e4_, x)
#include "ec_tmpl.h"
#undef EC
#undef FN
#undef NL

#define NL 6
#define FN(x) CAT(f6_, x)
#include "mont_tmpl.h"
#define EC(x) CAT(e6_, x)
#include "ec_tmpl.h"
#undef EC
#undef FN
#undef NL

/* field ids (shared with include/jf_b200.h) */
enum { JFO_BN254_FR = 0, JFO_BN254_FQ = 1, JFO_BLS12_381_FR = 2, JFO_BLS12_381_FQ = 3 };
enum { JFO_BN254 = 0, JFO_BLS12_381 = 1 };

static const uint64_t MOD_BN254_FR[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const uint64_t MOD_BN254_FQ[4] = {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const uint64_t MOD_BLS_FR[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
static const uint64_t MOD_BLS_FQ[6] = {0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                                       0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL};

static f4_params P_BN254_FR, P_BN254_FQ, P_BLS_FR;
static f6_params P_BLS_FQ;
static int g_init = 0;

/* x -> 2x mod p on plain limbs */
static void dbl_mod(uint64_t *x, const uint64_t *p, int n) {
    uint64_t carry = 0;
    for (int i = 0; i < n; i++) {
        uint64_t nc = x[i] >> 63;
        x[i] = (x[i] << 1) | carry;
        carry = nc;
    }
    int ge = 1;
    if (!carry) {
        for (int i = n - 1; i >= 0; i--) {
            if (x[i] > p[i]) { ge = 1; break; }
            if (x[i] < p[i]) { ge = 0; break; }
        }
    }
    if (carry || ge) {
        unsigned __int128 b = 0;
        for (int i = 0; i < n; i++) {
            unsigned __int128 t = (unsigned __int128)x[i] - p[i] - (uint64_t)b;
            x[i] = (uint64_t)t;
            b = (t >> 64) & 1;
        }
    }
}

static void derive(const uint64_t *p, int n, uint64_t *r, uint64_t *r2, uint64_t *inv) {
    uint64_t x[6] = {1, 0, 0, 0, 0, 0};
    for (int i = 0; i < 64 * n; i++) dbl_mod(x, p, n);
    memcpy(r, x, 8 * n);
    for (int i = 0; i < 64 * n; i++) dbl_mod(x, p, n);
    memcpy(r2, x, 8 * n);
    uint64_t y = 1; /* Newton: y = p^-1 mod 2^64 */
    for (int i = 0; i < 6; i++) y *= 2 - p[0] * y;
    *inv = (uint64_t)0 - y;
}

static void init_all(void) {
    if (g_init) return;
    memcpy(P_BN254_FR.p, MOD_BN254_FR, 32);
    derive(MOD_BN254_FR, 4, P_BN254_FR.r, P_BN254_FR.r2, &P_BN254_FR.inv);
    memcpy(P_BN254_FQ.p, MOD_BN254_FQ, 32);
    derive(MOD_BN254_FQ, 4, P_BN254_FQ.r, P_BN254_FQ.r2, &P_BN254_FQ.inv);
    memcpy(P_BLS_FR.p, MOD_BLS_FR, 32);
    derive(MOD_BLS_FR, 4, P_BLS_FR.r, P_BLS_FR.r2, &P_BLS_FR.inv);
    memcpy(P_BLS_FQ.p, MOD_BLS_FQ, 48);
    derive(MOD_BLS_FQ, 6, P_BLS_FQ.r, P_BLS_FQ.r2, &P_BLS_FQ.inv);
    g_init = 1;
}

static const f4_params *field4(int field) {
    init_all();
    switch (field) {
    case JFO_BN254_FR: return &P_BN254_FR;
    case JFO_BN254_FQ: return &P_BN254_FQ;
    case JFO_BLS12_381_FR: return &P_BLS_FR;
    default: return NULL;
    }
}

static int nthreads(int threads) {
#ifdef _OPENMP
    return threads > 0 ? threads : omp_get_max_threads();
#else
    (void)threads;
    return 1;
#endif
}

int jfo_max_threads(void) { return nthreads(0); }

/* ------------------------------------------------------------------------------------
 * element-wise field ops; op: 0 mul, 1 add, 2 sub, 3 sqr, 4 inv, 5 to_mont, 6 from_mont, 7 neg
 * ---------------------------------------------------------------------------------- */
int jfo_field_op(int field, int op, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n) {
    init_all();
    if (field == JFO_BLS12_381_FQ) {
        const f6_params *P = &P_BLS_FQ;
        for (size_t i = 0; i < n; i++) {
            const uint64_t *x = a + 6 * i, *y = b ? b + 6 * i : NULL;
            uint64_t *o = out + 6 * i;
            switch (op) {
            case 0: f6_mul(P, o, x, y); break;
            case 1: f6_add(P, o, x, y); break;
            case 2: f6_sub(P, o, x, y); break;
            case 3: f6_sqr(P, o, x); break;
            case 4: f6_inv(P, o, x); break;
            case 5: f6_to_mont(P, o, x); break;
            case 6: f6_from_mont(P, o, x); break;
            case 7: f6_neg(P, o, x); break;
            default: return -1;
            }
        }
        return 0;
    }
    const f4_params *P = field4(field);
    if (!P) return -1;
    for (size_t i = 0; i < n; i++) {
        const uint64_t *x = a + 4 * i, *y = b ? b + 4 * i : NULL;
        uint64_t *o = out + 4 * i;
        switch (op) {
        case 0: f4_mul(P, o, x, y); break;
        case 1: f4_add(P, o, x, y); break;
        case 2: f4_sub(P, o, x, y); break;
        case 3: f4_sqr(P, o, x); break;
        case 4: f4_inv(P, o, x); break;
        case 5: f4_to_mont(P, o, x); break;
        case 6: f4_from_mont(P, o, x); break;
        case 7: f4_neg(P, o, x); break;
        default: return -1;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------
 * Radix-2 NTT with ark-poly 0.4.2 semantics (natural order in and out, in place).
 *
 *   forward : x[j] *= offset^j (j < in_len), zero-pad to n, out[i] = sum_j x[j] w^(ij)
 *   inverse : c[j] = offset^-j n^-1 sum_i e[i] w^(-ij)
 *
 * with w = two_adic_root^(2^(two_adicity - log_n)).  Structure follows ark-poly's
 * fft_helper_in_place: forward = in-order -> out-of-order DIF passes then bit-reverse,
 * inverse = bit-reverse then out-of-order -> in-order DIT passes; butterflies of one pass
 * run in parallel (rayon there, OpenMP here) over a precomputed root table.
 * ---------------------------------------------------------------------------------- */
static void two_adic_root(const f4_params *P, int field, uint64_t *root) {
    /* GENERATOR^((p-1)/2^s): GENERATOR = 5 (BN254 Fr, s = 28), 7 (BLS12-381 Fr, s = 32) */
    uint64_t g[4] = {field == JFO_BN254_FR ? 5u : 7u, 0, 0, 0}, gm[4], e[4], one[4] = {1, 0, 0, 0};
    int s = field == JFO_BN254_FR ? 28 : 32;
    f4_to_mont(P, gm, g);
    f4_sub_raw(e, P->p, one);
    for (int k = 0; k < s; k++) { /* e >>= 1 */
        for (int i = 0; i < 4; i++) e[i] = (e[i] >> 1) | (i < 3 ? e[i + 1] << 63 : 0);
    }
    f4_pow(P, root, gm, e);
}

static size_t bitrev(size_t x, unsigned bits) {
    size_t r = 0;
    for (unsigned i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

/* x[j] = (x ? x[j] : 1) * c * base^j for j < n, chunked across threads (ark-poly's
 * distribute_powers_and_mul_by_const is chunk-parallel in the same way) */
static void powers_apply(const f4_params *P, uint64_t *x, int mul_into, const uint64_t *c, const uint64_t *base,
                         size_t n, int T) {
    size_t chunk = (n + (size_t)T - 1) / (size_t)T;
    if (chunk < 1024) chunk = 1024;
    size_t nchunks = (n + chunk - 1) / chunk;
#pragma omp parallel for num_threads(T) schedule(static)
    for (size_t ci = 0; ci < nchunks; ci++) {
        size_t lo = ci * chunk, hi = lo + chunk < n ? lo + chunk : n;
        uint64_t g[4], e[4] = {lo, 0, 0, 0};
        f4_pow(P, g, base, e);
        f4_mul(P, g, g, c);
        for (size_t j = lo; j < hi; j++) {
            if (mul_into) f4_mul(P, x + 4 * j, x + 4 * j, g);
            else f4_copy(x + 4 * j, g);
            f4_mul(P, g, g, base);
        }
    }
}

int jfo_ntt(int field, uint64_t *data, size_t in_len, unsigned log_n, int inverse,
            const uint64_t *coset_offset, size_t batch, size_t stride_elems, int threads) {
    const f4_params *P = field4(field);
    if (!P || (field != JFO_BN254_FR && field != JFO_BLS12_381_FR)) return -1;
    int s = field == JFO_BN254_FR ? 28 : 32;
    if ((int)log_n > s) return -2;
    size_t n = (size_t)1 << log_n;
    if (in_len > n || (batch > 1 && stride_elems < n)) return -1;
    int T = nthreads(threads);

    uint64_t w[4], off[4];
    two_adic_root(P, field, w);
    for (int k = 0; k < s - (int)log_n; k++) f4_sqr(P, w, w);
    if (inverse) f4_inv(P, w, w);
    int has_off = coset_offset != NULL && !f4_eq(coset_offset, P->r);
    if (has_off) {
        f4_copy(off, coset_offset);
        if (inverse) f4_inv(P, off, off);
    }
    /* root table w^k, k < n/2 */
    size_t half = n > 1 ? n / 2 : 1;
    uint64_t *roots = (uint64_t *)malloc(32 * half);
    powers_apply(P, roots, 0, P->r, w, half, T);

    for (size_t b = 0; b < batch; b++) {
        uint64_t *x = data + 4 * b * stride_elems;
        if (!inverse) {
            if (has_off) { /* distribute_powers over the given coefficients */
                powers_apply(P, x, 1, P->r, off, in_len, T);
            }
            for (size_t j = in_len; j < n; j++) f4_zero(x + 4 * j);
            for (size_t gap = n / 2; gap >= 1; gap >>= 1) { /* DIF */
                size_t step = n / (2 * gap);
#pragma omp parallel for num_threads(T) schedule(static)
                for (size_t t = 0; t < n / 2; t++) {
                    size_t blk = t / gap, k = t % gap;
                    uint64_t *lo = x + 4 * (blk * 2 * gap + k), *hi = lo + 4 * gap;
                    uint64_t d[4];
                    f4_sub(P, d, lo, hi);
                    f4_add(P, lo, lo, hi);
                    f4_mul(P, hi, d, roots + 4 * (k * step));
                }
            }
        } else {
            for (size_t j = in_len; j < n; j++) f4_zero(x + 4 * j);
        }
#pragma omp parallel for num_threads(T) schedule(static)
        for (size_t i = 0; i < n; i++) { /* derange */
            size_t j = bitrev(i, log_n);
            if (i < j) {
                uint64_t t[4];
                f4_copy(t, x + 4 * i);
                f4_copy(x + 4 * i, x + 4 * j);
                f4_copy(x + 4 * j, t);
            }
        }
        if (inverse) {
            for (size_t gap = 1; gap < n; gap <<= 1) { /* DIT */
                size_t step = n / (2 * gap);
#pragma omp parallel for num_threads(T) schedule(static)
                for (size_t t = 0; t < n / 2; t++) {
                    size_t blk = t / gap, k = t % gap;
                    uint64_t *lo = x + 4 * (blk * 2 * gap + k), *hi = lo + 4 * gap;
                    uint64_t v[4];
                    f4_mul(P, v, hi, roots + 4 * (k * step));
                    f4_sub(P, hi, lo, v);
                    f4_add(P, lo, lo, v);
                }
            }
            /* distribute_powers_and_mul_by_const(offset^-1, n^-1) */
            uint64_t ninv[4] = {n, 0, 0, 0};
            f4_to_mont(P, ninv, ninv);
            f4_inv(P, ninv, ninv);
            powers_apply(P, x, 1, ninv, has_off ? off : P->r, n, T);
        }
    }
    free(roots);
    return 0;
}

/* ------------------------------------------------------------------------------------
 * MSM = msm_bigint(bases, scalars).into_affine(); scalars canonical (4 limbs)
 * ---------------------------------------------------------------------------------- */
int jfo_msm(int curve, const uint64_t *points, size_t n_points, const uint64_t *scalars, size_t n_scalars,
            uint64_t *out_xy, int *out_inf, int threads) {
    init_all();
    size_t n = n_points < n_scalars ? n_points : n_scalars;
    int T = nthreads(threads);
    if (curve == JFO_BN254) {
        e4_jac r;
        e4_aff a;
        e4_msm(&P_BN254_FQ, &r, (const e4_aff *)points, scalars, n, 254, T);
        e4_jac_to_aff(&P_BN254_FQ, &a, &r);
        memcpy(out_xy, &a, sizeof a);
        *out_inf = e4_jac_is_inf(&r);
        return 0;
    }
    if (curve == JFO_BLS12_381) {
        e6_jac r;
        e6_aff a;
        e6_msm(&P_BLS_FQ, &r, (const e6_aff *)points, scalars, n, 255, T);
        e6_jac_to_aff(&P_BLS_FQ, &a, &r);
        memcpy(out_xy, &a, sizeof a);
        *out_inf = e6_jac_is_inf(&r);
        return 0;
    }
    return -1;
}

/* Number of windows the restated msm_bigint uses (= its maximum parallelism) */
int jfo_msm_windows(size_t n, int scalar_bits) {
    int c;
    if (n < 32) c = 3;
    else {
        int lg = 0;
        while (((size_t)1 << lg) < n) lg++;
        c = lg * 69 / 100 + 2;
    }
    return (scalar_bits + c - 1) / c;
}

static void gen_affine(int curve, e4_aff *g4, e6_aff *g6) {
    if (curve == JFO_BN254) {
        uint64_t one[4] = {1, 0, 0, 0}, two[4] = {2, 0, 0, 0};
        f4_to_mont(&P_BN254_FQ, g4->x, one);
        f4_to_mont(&P_BN254_FQ, g4->y, two);
    } else {
        static const uint64_t gx[6] = {0xfb3af00adb22c6bbULL, 0x6c55e83ff97a1aefULL, 0xa14e3a3f171bac58ULL,
                                       0xc3688c4f9774b905ULL, 0x2695638c4fa9ac0fULL, 0x17f1d3a73197d794ULL};
        static const uint64_t gy[6] = {0x0caa232946c5e7e1ULL, 0xd03cc744a2888ae4ULL, 0x00db18cb2c04b3edULL,
                                       0xfcf5e095d5d00af6ULL, 0xa09e30ed741d8ae4ULL, 0x08b3f481e3aaa0f1ULL};
        f6_to_mont(&P_BLS_FQ, g6->x, gx);
        f6_to_mont(&P_BLS_FQ, g6->y, gy);
    }
}

/* out[i] = scalars[i] * G  (scalars canonical, 4 limbs each); affine Montgomery out */
int jfo_fixed_base_mul(int curve, const uint64_t *scalars, size_t n, uint64_t *out, int threads) {
    init_all();
    int T = nthreads(threads);
    e4_aff g4;
    e6_aff g6;
    if (curve != JFO_BN254 && curve != JFO_BLS12_381) return -1;
    gen_affine(curve, &g4, &g6);
#pragma omp parallel for num_threads(T) schedule(dynamic, 16)
    for (size_t i = 0; i < n; i++) {
        if (curve == JFO_BN254) {
            e4_jac r;
            e4_scalar_mul(&P_BN254_FQ, &r, &g4, scalars + 4 * i, 4);
            e4_jac_to_aff(&P_BN254_FQ, (e4_aff *)(out + 8 * i), &r);
        } else {
            e6_jac r;
            e6_scalar_mul(&P_BLS_FQ, &r, &g6, scalars + 4 * i, 4);
            e6_jac_to_aff(&P_BLS_FQ, (e6_aff *)(out + 12 * i), &r);
        }
    }
    return 0;
}

/* powers_of_g[i] = beta^i * G, i < n  (srs.rs:118-153 with g = the curve generator) */
int jfo_gen_srs(int curve, const uint64_t *beta, size_t n, uint64_t *out, int threads) {
    init_all();
    const f4_params *F = curve == JFO_BN254 ? &P_BN254_FR : &P_BLS_FR;
    uint64_t *pw = (uint64_t *)malloc(32 * (n ? n : 1));
    uint64_t cur[4], bm[4];
    f4_to_mont(F, bm, beta);
    f4_copy(cur, F->r);
    for (size_t i = 0; i < n; i++) {
        f4_from_mont(F, pw + 4 * i, cur);
        f4_mul(F, cur, cur, bm);
    }
    int rc = jfo_fixed_base_mul(curve, pw, n, out, threads);
    free(pw);
    return rc;
}

/* Horner evaluation p(x); coeffs and x Montgomery (DensePolynomial::evaluate) */
int jfo_poly_eval(int field, const uint64_t *coeffs, size_t n, const uint64_t *x, uint64_t *out) {
    const f4_params *P = field4(field);
    if (!P) return -1;
    uint64_t acc[4];
    f4_zero(acc);
    for (size_t i = n; i-- > 0;) {
        f4_mul(P, acc, acc, x);
        f4_add(P, acc, acc, coeffs + 4 * i);
    }
    f4_copy(out, acc);
    return 0;
}

/* SplitMix64 rejection sampler shared with oracle/pyref.py::random_field_elems */
int jfo_random_field_elems(int field, size_t n, uint64_t seed, int montgomery, uint64_t *out) {
    const f4_params *P = field4(field);
    if (!P) return -1;
    int bits = field == JFO_BLS12_381_FR ? 255 : 254;
    uint64_t topmask = ((uint64_t)1 << (bits - 192)) - 1;
    uint64_t s = seed;
    size_t k = 0;
    while (k < n) {
        uint64_t v[4];
        for (int i = 0; i < 4; i++) {
            s += 0x9E3779B97F4A7C15ULL;
            uint64_t z = s;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
            v[i] = z ^ (z >> 31);
        }
        v[3] &= topmask;
        if (f4_geq(v, P->p)) continue;
        if (montgomery) f4_to_mont(P, out + 4 * k, v);
        else f4_copy(out + 4 * k, v);
        k++;
    }
    return 0;
}
