/* Montgomery field template for the CPU oracle (TEST INFRASTRUCTURE ONLY).
 *
 * Included several times by jf_oracle.c with
 *     #define NL   <number of u64 limbs>       (4: BN254 Fr/Fq, BLS12-381 Fr; 6: BLS12-381 Fq)
 *     #define FN(x) <prefix>##x
 * Restates ark-ff 0.4 `Fp<MontBackend<_, N>>`: little-endian u64 limbs, value kept as
 * a*R mod p with R = 2^(64 N), CIOS multiplication, results always fully reduced to [0,p).
 * ark-ff itself is not in /root/reference (un-vendored crate, plonk/Cargo.toml:15).
 */

typedef struct {
    uint64_t p[NL];    /* modulus */
    uint64_t r[NL];    /* R mod p  (Montgomery one) */
    uint64_t r2[NL];   /* R^2 mod p */
    uint64_t inv;      /* -p^-1 mod 2^64 */
} FN(params);

static inline int FN(geq)(const uint64_t *a, const uint64_t *b) {
    for (int i = NL - 1; i >= 0; i--) {
        if (a[i] > b[i]) return 1;
        if (a[i] < b[i]) return 0;
    }
    return 1;
}

static inline int FN(is_zero)(const uint64_t *a) {
    uint64_t t = 0;
    for (int i = 0; i < NL; i++) t |= a[i];
    return t == 0;
}

static inline int FN(eq)(const uint64_t *a, const uint64_t *b) {
    uint64_t t = 0;
    for (int i = 0; i < NL; i++) t |= a[i] ^ b[i];
    return t == 0;
}

static inline void FN(copy)(uint64_t *o, const uint64_t *a) {
    for (int i = 0; i < NL; i++) o[i] = a[i];
}

static inline void FN(zero)(uint64_t *o) {
    for (int i = 0; i < NL; i++) o[i] = 0;
}

static inline uint64_t FN(sub_raw)(uint64_t *o, const uint64_t *a, const uint64_t *b) {
    unsigned __int128 borrow = 0;
    for (int i = 0; i < NL; i++) {
        unsigned __int128 t = (unsigned __int128)a[i] - b[i] - (uint64_t)borrow;
        o[i] = (uint64_t)t;
        borrow = (t >> 64) & 1;
    }
    return (uint64_t)borrow;
}

static inline uint64_t FN(add_raw)(uint64_t *o, const uint64_t *a, const uint64_t *b) {
    unsigned __int128 carry = 0;
    for (int i = 0; i < NL; i++) {
        unsigned __int128 t = (unsigned __int128)a[i] + b[i] + (uint64_t)carry;
        o[i] = (uint64_t)t;
        carry = t >> 64;
    }
    return (uint64_t)carry;
}

static inline void FN(add)(const FN(params) *P, uint64_t *o, const uint64_t *a, const uint64_t *b) {
    uint64_t t[NL];
    uint64_t c = FN(add_raw)(t, a, b);
    if (c || FN(geq)(t, P->p)) FN(sub_raw)(t, t, P->p);
    FN(copy)(o, t);
}

static inline void FN(sub)(const FN(params) *P, uint64_t *o, const uint64_t *a, const uint64_t *b) {
    uint64_t t[NL];
    if (FN(sub_raw)(t, a, b)) FN(add_raw)(t, t, P->p);
    FN(copy)(o, t);
}

static inline void FN(neg)(const FN(params) *P, uint64_t *o, const uint64_t *a) {
    if (FN(is_zero)(a)) { FN(zero)(o); return; }
    FN(sub_raw)(o, P->p, a);
}

static inline void FN(dbl)(const FN(params) *P, uint64_t *o, const uint64_t *a) { FN(add)(P, o, a, a); }

/* CIOS Montgomery product: o = a*b/R mod p */
static inline void FN(mul)(const FN(params) *P, uint64_t *o, const uint64_t *a, const uint64_t *b) {
    uint64_t t[NL + 2];
    for (int i = 0; i < NL + 2; i++) t[i] = 0;
    for (int i = 0; i < NL; i++) {
        unsigned __int128 c = 0;
        for (int j = 0; j < NL; j++) {
            c += (unsigned __int128)a[j] * b[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[NL];
        t[NL] = (uint64_t)c;
        t[NL + 1] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * P->inv;
        c = (unsigned __int128)m * P->p[0] + t[0];
        c >>= 64;
        for (int j = 1; j < NL; j++) {
            c += (unsigned __int128)m * P->p[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[NL];
        t[NL - 1] = (uint64_t)c;
        t[NL] = t[NL + 1] + (uint64_t)(c >> 64);
    }
    if (t[NL] || FN(geq)(t, P->p)) FN(sub_raw)(t, t, P->p);
    FN(copy)(o, t);
}

static inline void FN(sqr)(const FN(params) *P, uint64_t *o, const uint64_t *a) { FN(mul)(P, o, a, a); }

static inline void FN(to_mont)(const FN(params) *P, uint64_t *o, const uint64_t *a) { FN(mul)(P, o, a, P->r2); }

static inline void FN(from_mont)(const FN(params) *P, uint64_t *o, const uint64_t *a) {
    uint64_t one[NL];
    FN(zero)(one);
    one[0] = 1;
    FN(mul)(P, o, a, one);
}

/* o = a^e (e: NL-limb little-endian plain integer); a, o in Montgomery form */
static inline void FN(pow)(const FN(params) *P, uint64_t *o, const uint64_t *a, const uint64_t *e) {
    uint64_t acc[NL], base[NL];
    FN(copy)(acc, P->r);
    FN(copy)(base, a);
    for (int i = 0; i < NL * 64; i++) {
        if ((e[i / 64] >> (i % 64)) & 1) FN(mul)(P, acc, acc, base);
        FN(sqr)(P, base, base);
    }
    FN(copy)(o, acc);
}

static inline void FN(pow_u64)(const FN(params) *P, uint64_t *o, const uint64_t *a, uint64_t e) {
    uint64_t ee[NL];
    FN(zero)(ee);
    ee[0] = e;
    FN(pow)(P, o, a, ee);
}

/* Fermat inversion a^(p-2); inv(0) = 0 */
static inline void FN(inv)(const FN(params) *P, uint64_t *o, const uint64_t *a) {
    uint64_t e[NL], two[NL];
    FN(zero)(two);
    two[0] = 2;
    FN(sub_raw)(e, P->p, two);
    FN(pow)(P, o, a, e);
}
