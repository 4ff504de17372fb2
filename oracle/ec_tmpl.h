/* Short-Weierstrass (a = 0) G1 template for the CPU oracle (TEST INFRASTRUCTURE ONLY).
 *
 *     #define NL, FN(x)  as for mont_tmpl.h (base field Fq)
 *     #define EC(x)      <curve prefix>##x
 *
 * Restates ark-ec 0.4 `short_weierstrass::{Affine, Projective}`: `Projective` is Jacobian
 * (x = X/Z^2, y = Y/Z^3), identity has Z = 0, `into_affine` maps identity to the
 * (0, 0, infinity = true) record.  Formulas are the EFD ones ark-ec documents
 * (dbl-2009-l, madd-2007-bl, add-2007-bl).  ark-ec is not in /root/reference.
 */

typedef struct { uint64_t x[NL], y[NL], z[NL]; } EC(jac);
typedef struct { uint64_t x[NL], y[NL]; } EC(aff); /* (0,0) == identity */

static inline int EC(aff_is_inf)(const EC(aff) *a) { return FN(is_zero)(a->x) && FN(is_zero)(a->y); }
static inline int EC(jac_is_inf)(const EC(jac) *a) { return FN(is_zero)(a->z); }

static inline void EC(jac_set_inf)(const FN(params) *P, EC(jac) *o) {
    FN(copy)(o->x, P->r);
    FN(copy)(o->y, P->r);
    FN(zero)(o->z);
}

static inline void EC(jac_from_aff)(const FN(params) *P, EC(jac) *o, const EC(aff) *a) {
    if (EC(aff_is_inf)(a)) { EC(jac_set_inf)(P, o); return; }
    FN(copy)(o->x, a->x);
    FN(copy)(o->y, a->y);
    FN(copy)(o->z, P->r);
}

/* dbl-2009-l (a = 0): 2M + 5S */
static inline void EC(jac_dbl)(const FN(params) *P, EC(jac) *o, const EC(jac) *p) {
    if (EC(jac_is_inf)(p)) { *o = *p; return; }
    uint64_t A[NL], B[NL], C[NL], D[NL], E[NL], F[NL], t[NL];
    FN(sqr)(P, A, p->x);
    FN(sqr)(P, B, p->y);
    FN(sqr)(P, C, B);
    FN(add)(P, t, p->x, B);
    FN(sqr)(P, t, t);
    FN(sub)(P, t, t, A);
    FN(sub)(P, t, t, C);
    FN(dbl)(P, D, t);
    FN(dbl)(P, E, A);
    FN(add)(P, E, E, A);
    FN(sqr)(P, F, E);
    uint64_t z3[NL];
    FN(mul)(P, z3, p->y, p->z);
    FN(dbl)(P, z3, z3);
    FN(sub)(P, t, F, D);
    FN(sub)(P, o->x, t, D);
    FN(sub)(P, t, D, o->x);
    FN(mul)(P, t, E, t);
    FN(dbl)(P, C, C);
    FN(dbl)(P, C, C);
    FN(dbl)(P, C, C);
    FN(sub)(P, o->y, t, C);
    FN(copy)(o->z, z3);
}

/* madd-2007-bl: Jacobian += affine, with the doubling / cancellation cases handled */
static inline void EC(jac_add_aff)(const FN(params) *P, EC(jac) *o, const EC(jac) *p, const EC(aff) *q) {
    if (EC(aff_is_inf)(q)) { *o = *p; return; }
    if (EC(jac_is_inf)(p)) { EC(jac_from_aff)(P, o, q); return; }
    uint64_t Z1Z1[NL], U2[NL], S2[NL], H[NL], HH[NL], I[NL], J[NL], r[NL], V[NL], t[NL];
    FN(sqr)(P, Z1Z1, p->z);
    FN(mul)(P, U2, q->x, Z1Z1);
    FN(mul)(P, S2, q->y, p->z);
    FN(mul)(P, S2, S2, Z1Z1);
    if (FN(eq)(U2, p->x)) {
        if (FN(eq)(S2, p->y)) { EC(jac_dbl)(P, o, p); return; }
        EC(jac_set_inf)(P, o);
        return;
    }
    FN(sub)(P, H, U2, p->x);
    FN(sqr)(P, HH, H);
    FN(dbl)(P, I, HH);
    FN(dbl)(P, I, I);
    FN(mul)(P, J, H, I);
    FN(sub)(P, r, S2, p->y);
    FN(dbl)(P, r, r);
    FN(mul)(P, V, p->x, I);
    uint64_t x3[NL], y3[NL], z3[NL];
    FN(sqr)(P, x3, r);
    FN(sub)(P, x3, x3, J);
    FN(sub)(P, x3, x3, V);
    FN(sub)(P, x3, x3, V);
    FN(sub)(P, t, V, x3);
    FN(mul)(P, y3, r, t);
    FN(mul)(P, t, p->y, J);
    FN(dbl)(P, t, t);
    FN(sub)(P, y3, y3, t);
    FN(add)(P, z3, p->z, H);
    FN(sqr)(P, z3, z3);
    FN(sub)(P, z3, z3, Z1Z1);
    FN(sub)(P, z3, z3, HH);
    FN(copy)(o->x, x3);
    FN(copy)(o->y, y3);
    FN(copy)(o->z, z3);
}

static inline void EC(aff_neg)(const FN(params) *P, EC(aff) *o, const EC(aff) *a) {
    FN(copy)(o->x, a->x);
    FN(neg)(P, o->y, a->y);
}

/* add-2007-bl */
static inline void EC(jac_add)(const FN(params) *P, EC(jac) *o, const EC(jac) *p, const EC(jac) *q) {
    if (EC(jac_is_inf)(q)) { *o = *p; return; }
    if (EC(jac_is_inf)(p)) { *o = *q; return; }
    uint64_t Z1Z1[NL], Z2Z2[NL], U1[NL], U2[NL], S1[NL], S2[NL], H[NL], I[NL], J[NL], r[NL], V[NL], t[NL];
    FN(sqr)(P, Z1Z1, p->z);
    FN(sqr)(P, Z2Z2, q->z);
    FN(mul)(P, U1, p->x, Z2Z2);
    FN(mul)(P, U2, q->x, Z1Z1);
    FN(mul)(P, S1, p->y, q->z);
    FN(mul)(P, S1, S1, Z2Z2);
    FN(mul)(P, S2, q->y, p->z);
    FN(mul)(P, S2, S2, Z1Z1);
    if (FN(eq)(U1, U2)) {
        if (FN(eq)(S1, S2)) { EC(jac_dbl)(P, o, p); return; }
        EC(jac_set_inf)(P, o);
        return;
    }
    FN(sub)(P, H, U2, U1);
    FN(dbl)(P, I, H);
    FN(sqr)(P, I, I);
    FN(mul)(P, J, H, I);
    FN(sub)(P, r, S2, S1);
    FN(dbl)(P, r, r);
    FN(mul)(P, V, U1, I);
    uint64_t x3[NL], y3[NL], z3[NL];
    FN(sqr)(P, x3, r);
    FN(sub)(P, x3, x3, J);
    FN(sub)(P, x3, x3, V);
    FN(sub)(P, x3, x3, V);
    FN(sub)(P, t, V, x3);
    FN(mul)(P, y3, r, t);
    FN(mul)(P, t, S1, J);
    FN(dbl)(P, t, t);
    FN(sub)(P, y3, y3, t);
    FN(add)(P, z3, p->z, q->z);
    FN(sqr)(P, z3, z3);
    FN(sub)(P, z3, z3, Z1Z1);
    FN(sub)(P, z3, z3, Z2Z2);
    FN(mul)(P, z3, z3, H);
    FN(copy)(o->x, x3);
    FN(copy)(o->y, y3);
    FN(copy)(o->z, z3);
}

/* `into_affine`: identity -> (0, 0) */
static inline void EC(jac_to_aff)(const FN(params) *P, EC(aff) *o, const EC(jac) *p) {
    if (EC(jac_is_inf)(p)) { FN(zero)(o->x); FN(zero)(o->y); return; }
    uint64_t zi[NL], zi2[NL];
    FN(inv)(P, zi, p->z);
    FN(sqr)(P, zi2, zi);
    FN(mul)(P, o->x, p->x, zi2);
    FN(mul)(P, zi2, zi2, zi);
    FN(mul)(P, o->y, p->y, zi2);
}

/* k * q by left-to-right double-and-add; k: `klimbs` u64 limbs, plain integer */
static inline void EC(scalar_mul)(const FN(params) *P, EC(jac) *o, const EC(aff) *q, const uint64_t *k, int klimbs) {
    EC(jac) acc;
    EC(jac_set_inf)(P, &acc);
    for (int i = klimbs * 64 - 1; i >= 0; i--) {
        EC(jac_dbl)(P, &acc, &acc);
        if ((k[i / 64] >> (i % 64)) & 1) EC(jac_add_aff)(P, &acc, &acc, q);
    }
    *o = acc;
}

/* make_digits + bucket method, as `VariableBaseMSM::msm_bigint` does in ark-ec 0.4.2
 * (signed c-bit digits, one rayon task per window, serial bucket accumulation with mixed
 * additions, running-sum reduction, Horner fold over windows).  `scalar_bits` is the
 * scalar field's MODULUS_BIT_SIZE; scalars are 4 u64 limbs, canonical (not Montgomery). */
static void EC(msm)(const FN(params) *P, EC(jac) *out, const EC(aff) *bases, const uint64_t *scalars,
                    size_t n, int scalar_bits, int threads) {
    int c;
    if (n < 32) c = 3;
    else {
        int lg = 0;
        while (((size_t)1 << lg) < n) lg++;   /* ark_std::log2 = ceil(log2 n) */
        c = lg * 69 / 100 + 2;
    }
    int nd = (scalar_bits + c - 1) / c;
    int32_t *digits = (int32_t *)malloc(sizeof(int32_t) * n * (size_t)nd);
    const uint64_t radix = (uint64_t)1 << c, mask = radix - 1;
#pragma omp parallel for num_threads(threads) schedule(static)
    for (size_t i = 0; i < n; i++) {
        const uint64_t *s = scalars + 4 * i;
        uint64_t carry = 0;
        for (int d = 0; d < nd; d++) {
            int off = d * c, wi = off / 64, bi = off % 64;
            uint64_t buf;
            if (bi < 64 - c || wi == 3) buf = s[wi] >> bi;
            else buf = (s[wi] >> bi) | (s[wi + 1] << (64 - bi));
            uint64_t coef = carry + (buf & mask);
            carry = (coef + radix / 2) >> c;
            digits[i * nd + d] = (int32_t)((int64_t)coef - (int64_t)(carry << c));
        }
        digits[i * nd + nd - 1] += (int32_t)(carry << c);
    }
    EC(jac) *wsum = (EC(jac) *)malloc(sizeof(EC(jac)) * nd);
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
    for (int w = 0; w < nd; w++) {
        size_t nb = (size_t)1 << c;
        EC(jac) *buckets = (EC(jac) *)malloc(sizeof(EC(jac)) * nb);
        for (size_t b = 0; b < nb; b++) EC(jac_set_inf)(P, &buckets[b]);
        for (size_t i = 0; i < n; i++) {
            int32_t d = digits[i * nd + w];
            if (d > 0) EC(jac_add_aff)(P, &buckets[d - 1], &buckets[d - 1], &bases[i]);
            else if (d < 0) {
                EC(aff) neg;
                EC(aff_neg)(P, &neg, &bases[i]);
                EC(jac_add_aff)(P, &buckets[-d - 1], &buckets[-d - 1], &neg);
            }
        }
        EC(jac) run, res;
        EC(jac_set_inf)(P, &run);
        EC(jac_set_inf)(P, &res);
        for (size_t b = nb; b-- > 0;) {
            EC(jac_add)(P, &run, &run, &buckets[b]);
            EC(jac_add)(P, &res, &res, &run);
        }
        wsum[w] = res;
        free(buckets);
    }
    EC(jac) total;
    EC(jac_set_inf)(P, &total);
    for (int w = nd - 1; w >= 1; w--) {
        EC(jac_add)(P, &total, &total, &wsum[w]);
        for (int k = 0; k < c; k++) EC(jac_dbl)(P, &total, &total);
    }
    EC(jac_add)(P, out, &total, &wsum[0]);
    free(wsum);
    free(digits);
}
