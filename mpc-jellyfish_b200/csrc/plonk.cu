// plonk.cu -- device-resident TurboPlonk / UltraPlonk prover rounds around the MSM / NTT kernels
// (the "next" rows either side of the hot path).
//
// UltraPlonk (6 wire types, Plookup; `Plonk<C, 6>`) adds, on top of everything below:
//   PlonkCircuit::compute_{merged_lookup_table,lookup_sorted_vec_polynomials,lookup_prod_polynomial,range_table,key_table,
//   table_dom_sep,q_dom_sep}_polynomial        relation/src/constraint_system.rs:1261-1492
//   Prover::run_plookup_1st/2nd_round, compute_plookup_evaluations, compute_quotient_plookup_contribution,
//   compute_lin_poly_plookup_contribution, plookup_(shifted_)open_polys_ref
//                                          plonk/src/proof_system/prover.rs:98-123,150-190,239-297,427-460,773-888,1037-1113
// The quotient of an UltraPlonk instance has degree 6 n + 8: round 3 runs on seven sub-cosets of n points (six for TurboPlonk).
//
// Replaces, for one TurboPlonk instance with 5 wire types and no Plookup (`Plonk<C, 5>`):
//   PlonkKzgSnark::preprocess              plonk/src/proof_system/snark.rs:529-611
//   PlonkKzgSnark::batch_prove_internal    plonk/src/proof_system/snark.rs:201-469
//   Prover::run_1st..3rd_round, compute_evaluations, compute_(non_)quotient_component_for_lin_poly,
//   compute_opening_proofs, mask_polynomial, split_quotient_polynomial, compute_quotient_polynomial
//                                          plonk/src/proof_system/prover.rs:72-87,125-141,192-235,302-419,463-673,902-1034
//   PlonkCircuit::compute_{wire,pub_input,prod_permutation,selector,extended_permutation}_polynomial(s)
//                                          relation/src/constraint_system.rs:1162-1259
// Circuit construction (gates, variables, the wire permutation) stays with the caller: it hands over
// the selector columns, the extended permutation sigma*(i) = k_w * g^j, the wire -> variable map and,
// per proof, the witness.  Everything between the witness upload and the 13 commitments + 10
// evaluations of the proof stays in HBM; the host only runs the Fiat-Shamir transcript (5 round
// trips of a few hundred bytes).
//
// Polynomial algebra on the device (all exact field arithmetic, so results are bit-identical to the
// reference's serial loops whatever the evaluation order):
//   grand product      z = prefix-product(a) * suffix-product(b) / prod(b): two parallel scans and ONE
//                      inversion instead of n divisions (constraint_system.rs:1197-1223)
//   quotient           one fused pointwise kernel over the 8n coset (prover.rs:605-663,677-759),
//                      1/(n (x - 1)) from a table built once per proving key
//   evaluation         blocked Horner + tree sum (prover.rs:216-235)
//   lin. combination   one kernel over up to 32 (scalar, polynomial) pairs (prover.rs:339-360,490-502,963-1034)
//   p / (X - z)        q_j = z^-(j+1) * sum_{i>j} p_i z^i: power scaling + one additive suffix scan
//                      (prover.rs:504-506; ark-poly's `/` drops the remainder)
#include <string.h>
#include <algorithm>
#include "common.cuh"
#include "ec.cuh"
#include "transcript.hpp"

namespace jf {

static constexpr int NW_TURBO = 5, NW_ULTRA = 6;  // wire types: GATE_WIDTH + 1 (+ the range / lookup wire)
static constexpr int NW_MAX = 6;
static constexpr int PAD = 8;     // room above n for the masking coefficients
// sizes that depend on the Plonk type
template <int NWT> struct Dim {
    static constexpr int NW = NWT;
    static constexpr bool ULTRA = NWT == NW_ULTRA;
    static constexpr int NSEL = 13 + (ULTRA ? 1 : 0);  // q_lc[4], q_mul[2], q_hash[4], q_o, q_c, q_ecc [, q_lookup]
    // field elements drawn from the prng: 2 per wire polynomial, [3 + 3 for h1, h2,] 3 for z, [3 for the lookup product,]
    // NW - 1 split-quotient randomizers (prover.rs:463-486,946-957): 17 / 29
    static constexpr int NBLIND = 2 * NWT + 3 + (NWT - 1) + (ULTRA ? 9 : 0);
    static constexpr int BL_H = 2 * NWT;                 // h1 (3), h2 (3)       [Ultra]
    static constexpr int BL_Z = 2 * NWT + (ULTRA ? 6 : 0);
    static constexpr int BL_PL = BL_Z + 3;               // lookup product (3)   [Ultra]
    static constexpr int BL_SPLIT = BL_Z + 3 + (ULTRA ? 3 : 0);
};

template <class F> __device__ __forceinline__ Fp<F> ldf(const Fp<F> *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 a = q[0], b = q[1];
    Fp<F> r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
template <class F> __device__ __forceinline__ void stf(Fp<F> *p, const Fp<F> &r) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
// a^e for a small exponent (only the set bits of e are walked)
template <class F> __host__ __device__ __forceinline__ Fp<F> pow_small(const Fp<F> &a, uint64_t e) {
    Fp<F> acc = Fp<F>::one();
    int top = 63;
    while (top >= 0 && !((e >> top) & 1)) top--;
    for (int i = top; i >= 0; i--) {
        acc = Fp<F>::sqr(acc);
        if ((e >> i) & 1) acc = Fp<F>::mul(acc, a);
    }
    return acc;
}

// ---- parallel scans over field elements ------------------------------------------------------
struct OpMul {
    template <class T> static __device__ __forceinline__ T apply(const T &a, const T &b) { return T::mul(a, b); }
    template <class T> static __device__ __forceinline__ T identity() { return T::one(); }
};
struct OpAdd {
    template <class T> static __device__ __forceinline__ T apply(const T &a, const T &b) { return T::add(a, b); }
    template <class T> static __device__ __forceinline__ T identity() { return T::zero(); }
};
static constexpr int FS_T = 256, FS_I = 4, FS_TILE = FS_T * FS_I;

template <class F> __device__ __forceinline__ Fp<F> shfl_up_f(const Fp<F> &x, int d) {
    Fp<F> r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = __shfl_up_sync(0xffffffffu, x.v[i], d);
    return r;
}

// Inclusive scan of one tile per CTA (logical order reversed when REV); totals[b] = tile sum.
template <class F, class Op, bool REV>
__global__ void __launch_bounds__(FS_T) fscan_local_kernel(const Fp<F> *in, Fp<F> *out, Fp<F> *totals, size_t n) {
    using E = Fp<F>;
    __shared__ E wt[FS_T / 32];
    const size_t base = (size_t)blockIdx.x * FS_TILE + (size_t)threadIdx.x * FS_I;
    E v[FS_I];
#pragma unroll
    for (int k = 0; k < FS_I; k++) {
        const size_t L = base + k;
        v[k] = L < n ? ldf(in + (REV ? n - 1 - L : L)) : Op::template identity<E>();
        if (k) v[k] = Op::apply(v[k - 1], v[k]);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    E tot = v[FS_I - 1];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        E o = shfl_up_f(tot, d);
        if (lane >= d) tot = Op::apply(o, tot);
    }
    if (lane == 31) wt[warp] = tot;
    __syncthreads();
    if (threadIdx.x == 0) {
        E acc = Op::template identity<E>();
        for (int w = 0; w < FS_T / 32; w++) {
            E t = wt[w];
            wt[w] = acc;
            acc = Op::apply(acc, t);
        }
        if (totals) stf(totals + blockIdx.x, acc);
    }
    __syncthreads();
    E excl = shfl_up_f(tot, 1);
    if (lane == 0) excl = Op::template identity<E>();
    excl = Op::apply(wt[warp], excl);
#pragma unroll
    for (int k = 0; k < FS_I; k++) {
        const size_t L = base + k;
        if (L < n) stf(out + (REV ? n - 1 - L : L), Op::apply(excl, v[k]));
    }
}

// out[L] = op(scanned_totals[tile(L) - 1], out[L]) for every tile but the first
template <class F, class Op, bool REV>
__global__ void fscan_apply_kernel(Fp<F> *out, const Fp<F> *scanned_totals, size_t n) {
    const size_t L = (size_t)blockIdx.x * blockDim.x + threadIdx.x + FS_TILE;
    if (L >= n) return;
    const size_t tile = L / FS_TILE;
    Fp<F> *p = out + (REV ? n - 1 - L : L);
    stf(p, Op::apply(ldf(scanned_totals + tile - 1), ldf(p)));
}

// Inclusive scan (forward, or reverse = suffix scan) of n elements; `tmp` holds >= n/512 + 64 elements.
template <class F, class Op, bool REV>
static int fscan(jf_ctx *ctx, const Fp<F> *in, Fp<F> *out, size_t n, Fp<F> *tmp) {
    if (n == 0) return JF_OK;
    const size_t tiles = (n + FS_TILE - 1) / FS_TILE;
    if (tiles == 1) {
        JF_LAUNCH(ctx, "fscan_local", fscan_local_kernel<F, Op, REV><<<1, FS_T, 0, ctx->stream>>>(in, out, nullptr, n));
        return JF_OK;
    }
    JF_LAUNCH(ctx, "fscan_local", fscan_local_kernel<F, Op, REV><<<(unsigned)tiles, FS_T, 0, ctx->stream>>>(in, out, tmp, n));
    // totals are in logical tile order: always a forward scan
    JF_TRY((fscan<F, Op, false>(ctx, tmp, tmp, tiles, tmp + tiles)));
    const size_t rest = n - FS_TILE;
    JF_LAUNCH(ctx, "fscan_apply", fscan_apply_kernel<F, Op, REV><<<(unsigned)((rest + 255) / 256), 256, 0, ctx->stream>>>(out, tmp, n));
    return JF_OK;
}

// ---- small element-wise kernels ------------------------------------------------------------------
// wire value columns: wv[j*n + i] = witness[wire_vars[j*n + i]]; pi[i] = witness of the io gates' output wire
template <class F>
__global__ void gather_wires_kernel(const Fp<F> *witness, const uint32_t *wire_vars, size_t total, Fp<F> *wv) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) stf(wv + i, ldf(witness + wire_vars[i]));
}
template <class F>
__global__ void pub_input_kernel(const Fp<F> *witness, const uint32_t *wire_vars_out, const uint32_t *gate_ids, uint32_t num,
                                 Fp<F> *pi) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < num) stf(pi + gate_ids[i], ldf(witness + wire_vars_out[gate_ids[i]]));
}
// mask_polynomial (prover.rs:463-486): poly += (b_0 + .. + b_hb X^hb) (X^n - 1); entries n.. are zero on entry
template <class F> __global__ void blind_kernel(Fp<F> *poly, size_t n, const Fp<F> *b, int count) {
    int i = threadIdx.x;
    if (i >= count) return;
    Fp<F> bi = ldf(b + i);
    stf(poly + i, Fp<F>::sub(ldf(poly + i), bi));
    stf(poly + n + i, bi);
}
// copy with row strides (vectors of `len` elements); rows beyond are untouched
template <class F>
__global__ void copy_rows_kernel(const Fp<F> *src, size_t src_stride, Fp<F> *dst, size_t dst_stride, size_t len, size_t zero_to) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t row = blockIdx.y;
    if (i < len) stf(dst + row * dst_stride + i, ldf(src + row * src_stride + i));
    else if (i < zero_to) stf(dst + row * dst_stride + i, Fp<F>::zero());
}

// ---- round 2: numerators / denominators of the grand product ---------------------------------------
template <class F, int NW> struct PermArgs {
    const Fp<F> *wv;       // NW x n wire values
    const Fp<F> *sigma;    // NW x n extended permutation values
    Fp<F> beta_k[NW];      // beta * k_j
    Fp<F> beta, gamma, omega;
    Fp<F> *a, *b;          // n each
    uint32_t n;
};
static constexpr int PERM_I = 8;
template <class F, int NW> __global__ void perm_ab_kernel(const __grid_constant__ PermArgs<F, NW> p) {
    using E = Fp<F>;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t start = (uint64_t)t * PERM_I;
    if (start >= p.n) return;
    E x = pow_small(p.omega, start);
    for (int k = 0; k < PERM_I; k++) {
        const uint64_t j = start + k;
        if (j >= p.n) break;
        E a = E::one(), b = E::one();
        if (j < p.n - 1) {
#pragma unroll
            for (int w = 0; w < NW; w++) {
                E tmp = E::add(ldf(p.wv + (size_t)w * p.n + j), p.gamma);
                a = E::mul(a, E::add(tmp, E::mul(p.beta_k[w], x)));
                b = E::mul(b, E::add(tmp, E::mul(p.beta, ldf(p.sigma + (size_t)w * p.n + j))));
            }
        }
        stf(p.a + j, a);
        stf(p.b + j, b);
        x = E::mul(x, p.omega);
    }
}
template <class F> __global__ void inv_one_kernel(const Fp<F> *in, Fp<F> *out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) stf(out, Fp<F>::inv(ldf(in)));
}
// z[i] = (prod_{j<i} a_j) * (prod_{j>=i} b_j) / prod(b)
template <class F>
__global__ void z_combine_kernel(const Fp<F> *pa, const Fp<F> *sb, const Fp<F> *binv, Fp<F> *z, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fp<F> v = Fp<F>::mul(ldf(sb + i), ldf(binv));
    if (i) v = Fp<F>::mul(v, ldf(pa + i - 1));
    stf(z + i, v);
}
// batch inversion from prefix / suffix products: out[i] = P[i-1] * S[i+1] / total
template <class F>
__global__ void inv_combine_kernel(const Fp<F> *pp, const Fp<F> *sp, const Fp<F> *tinv, Fp<F> *out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fp<F> v = ldf(tinv);
    if (i) v = Fp<F>::mul(v, ldf(pp + i - 1));
    if (i + 1 < n) v = Fp<F>::mul(v, ldf(sp + i + 1));
    stf(out + i, v);
}
// d[i] = scale * (x_i - 1),  x_i = hi[i >> lo_bits] * lo[i & mask]
// sub != 0: position i = r n + j holds the point of index 8 j + r (sub-coset r of the 8n coset, see QuotArgs)
template <class F>
__global__ void xm1_kernel(const Fp<F> *lo, const Fp<F> *hi, int lo_bits, Fp<F> scale, Fp<F> *d, size_t m, int sub, uint32_t log_n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const size_t pos = i;
    if (sub) i = ((i & (((size_t)1 << log_n) - 1)) << 3) | (i >> log_n);
    Fp<F> x = Fp<F>::mul(ldf(lo + (i & (((size_t)1 << lo_bits) - 1))), ldf(hi + (i >> lo_bits)));
    stf(d + pos, Fp<F>::mul(scale, Fp<F>::sub(x, Fp<F>::one())));
}
// table[i] = scale * base^(i * step)
template <class F> __global__ void pow_tab_kernel(Fp<F> *table, Fp<F> base, Fp<F> scale, uint64_t step, uint32_t count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) stf(table + i, Fp<F>::mul(scale, pow_small(base, (uint64_t)i * step)));
}

// ---- round 3: the quotient on the 8n coset -----------------------------------------------------------
// Plookup terms of the quotient (prover.rs:773-888): coset evaluations of the four table polynomials, h1, h2 and the lookup product
template <class F> struct LookupQuot {
    const Fp<F> *range, *key, *tds, *qds, *h1, *h2, *pl;  // m each; q_lookup is selector 13
    Fp<F> tau, bp1, gb, alpha3;                           // 1 + beta, gamma (1 + beta)
};
template <class F, int NW> struct QuotArgs {
    const Fp<F> *sel;    // NSEL x m coset evaluations
    const Fp<F> *sig;    // NW x m
    const Fp<F> *w;      // NW x m
    const Fp<F> *z, *pi; // m each
    LookupQuot<F> lk;    // UltraPlonk only
    const Fp<F> *inv_nx1;  // 1 / (n (x_i - 1))
    const Fp<F> *x_lo, *x_hi;
    int lo_bits;
    Fp<F> beta_k[NW], beta, gamma, alpha, alpha2;
    Fp<F> zh_inv[8];
    Fp<F> omega_inv;    // w_n^-1 (UltraPlonk)
    Fp<F> scale;        // batch_prove: instance i enters the ONE quotient with alpha_base_i (prover.rs:661-669)
    uint32_t acc_mode;  // 0: out = t;  1: out += scale * t
    Fp<F> *out;
    uint32_t m, ratio;  // m: evaluation points per polynomial (8n, or 6n by sub-cosets)
    uint32_t zero_sel;  // bit s set: selector s is the zero polynomial (its term and its loads are skipped)
    // sub != 0: the evaluations are laid out by sub-coset, position r n + j <-> the point g w_8n^(8 j + r), r < sub.
    // Within a row the domain generator w_n is a shift by one, and X^n - 1 is the constant 1 / zh_inv[r].
    uint32_t sub, log_n;
    // row_map[l]: the sub-coset held in local row l.  Identity unless round 3 is dealt out over several GPUs
    // (jf_plonk_pk_shard_commits), where a rank holds only some of the rows; inv_nx1 always covers all of them.
    uint32_t row_map[8];
};
template <class F> __device__ __forceinline__ Fp<F> pow5(const Fp<F> &x) {
    Fp<F> x2 = Fp<F>::sqr(x);
    return Fp<F>::mul(x, Fp<F>::sqr(x2));
}
// eval_merged_table / eval_merged_lookup_witness (structs.rs:926-956): a + q tau (b + tau (c + tau (d + tau e)))
template <class F>
__device__ __forceinline__ Fp<F> merged5(const Fp<F> &tau, const Fp<F> &a, const Fp<F> &q, const Fp<F> &b, const Fp<F> &c, const Fp<F> &d,
                                         const Fp<F> &e) {
    using E = Fp<F>;
    E t = E::add(d, E::mul(tau, e));
    t = E::add(c, E::mul(tau, t));
    t = E::add(b, E::mul(tau, t));
    return E::add(a, E::mul(E::mul(q, tau), t));
}
template <class F, int NW> __global__ void __launch_bounds__(128) quotient_kernel(const __grid_constant__ QuotArgs<F, NW> q) {
    using E = Fp<F>;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= q.m) return;
    const size_t m = q.m;
    E w[NW];
#pragma unroll
    for (int j = 0; j < NW; j++) w[j] = ldf(q.w + j * m + i);
    auto S = [&](int s) { return ldf(q.sel + s * m + i); };
    // circuit part (prover.rs:696-708); selector order q_lc 0-3, q_mul 4-5, q_hash 6-9, q_o 10, q_c 11, q_ecc 12
    auto on = [&](int s) { return !((q.zero_sel >> s) & 1u); };  // uniform across the grid
    E w01 = E::mul(w[0], w[1]), w23 = E::mul(w[2], w[3]);
    E t = q.pi ? ldf(q.pi + i) : E::zero();  // no public inputs: the PI polynomial is zero (flags & 2)
    if (on(11)) t = E::add(t, S(11));
    if (on(0)) t = E::add(t, E::mul(S(0), w[0]));
    if (on(1)) t = E::add(t, E::mul(S(1), w[1]));
    if (on(2)) t = E::add(t, E::mul(S(2), w[2]));
    if (on(3)) t = E::add(t, E::mul(S(3), w[3]));
    if (on(4)) t = E::add(t, E::mul(S(4), w01));
    if (on(5)) t = E::add(t, E::mul(S(5), w23));
    if (on(12)) t = E::add(t, E::mul(S(12), E::mul(E::mul(w01, w23), w[4])));
    if (on(6)) t = E::add(t, E::mul(S(6), pow5(w[0])));
    if (on(7)) t = E::add(t, E::mul(S(7), pow5(w[1])));
    if (on(8)) t = E::add(t, E::mul(S(8), pow5(w[2])));
    if (on(9)) t = E::add(t, E::mul(S(9), pow5(w[3])));
    if (on(10)) t = E::sub(t, E::mul(S(10), w[4]));
    // copy constraints (prover.rs:743-756)
    uint32_t xi = i, inext = i + q.ratio, row = 0, gi = i, ginext;
    if (q.sub) {
        const uint32_t nmask = (1u << q.log_n) - 1, jj = i & nmask;
        row = q.row_map[i >> q.log_n];
        xi = (jj << 3) | row;
        inext = (i & ~nmask) | ((jj + 1) & nmask);
        gi = (row << q.log_n) | jj;  // position in the per-key table, which holds every sub-coset
        ginext = (row << q.log_n) | ((jj + 1) & nmask);
    } else {
        if (inext >= q.m) inext -= q.m;
        row = i % q.ratio;
        ginext = inext;
    }
    const E x = E::mul(ldf(q.x_lo + (xi & ((1u << q.lo_bits) - 1))), ldf(q.x_hi + (xi >> q.lo_bits)));
    const E zx = ldf(q.z + i);
    E r1 = zx, r2 = ldf(q.z + inext);
#pragma unroll
    for (int j = 0; j < NW; j++) {
        E wg = E::add(w[j], q.gamma);
        r1 = E::mul(r1, E::add(wg, E::mul(q.beta_k[j], x)));
        r2 = E::mul(r2, E::add(wg, E::mul(q.beta, ldf(q.sig + j * m + i))));
    }
    t = E::add(t, E::mul(q.alpha, E::sub(r1, r2)));
    E t2 = E::mul(E::mul(q.alpha2, E::sub(zx, E::one())), ldf(q.inv_nx1 + gi));
    if constexpr (NW == NW_ULTRA) {
        // Plookup (prover.rs:773-888).  L_1 / Z_H = 1 / (n (x - 1)) = inv_nx1[i]; L_n / Z_H = w^-1 / (n (x - w^-1)) = 1 / (n (w x - 1))
        // = inv_nx1[inext], because w x is the point one domain step further on the coset.
        const LookupQuot<F> &k = q.lk;
        const E lag1 = ldf(q.inv_nx1 + gi), lagn = ldf(q.inv_nx1 + ginext);
        const E ql = S(13), ql_n = ldf(q.sel + 13 * m + inext);
        const E h1x = ldf(k.h1 + i), h1n = ldf(k.h1 + inext), h2x = ldf(k.h2 + i), h2n = ldf(k.h2 + inext);
        const E px = ldf(k.pl + i), pn = ldf(k.pl + inext);
        const E mt_x = merged5(k.tau, ldf(k.range + i), ql, ldf(k.tds + i), ldf(k.key + i), w[3], w[4]);
        const E mt_n = merged5(k.tau, ldf(k.range + inext), ql_n, ldf(k.tds + inext), ldf(k.key + inext), ldf(q.w + 3 * m + inext),
                               ldf(q.w + 4 * m + inext));
        const E ml_x = merged5(k.tau, w[5], ql, ldf(k.qds + i), w[0], w[1], w[2]);
        E ap = k.alpha3;
        t2 = E::add(t2, E::mul(ap, E::mul(E::sub(h1x, h2n), lagn)));
        ap = E::mul(ap, q.alpha);
        const E pm1 = E::sub(px, E::one());
        t2 = E::add(t2, E::mul(ap, E::mul(pm1, lag1)));
        ap = E::mul(ap, q.alpha);
        t2 = E::add(t2, E::mul(ap, E::mul(pm1, lagn)));
        ap = E::mul(ap, q.alpha);
        // (x - w^-1) = (w x - 1) / w; with lagn = 1 / (n (w x - 1)) it is cheaper to rebuild x - w^-1 from x and the constant
        const E a = E::mul(E::mul(E::mul(px, k.bp1), E::add(q.gamma, ml_x)), E::add(E::add(k.gb, mt_x), E::mul(q.beta, mt_n)));
        const E b = E::mul(E::mul(pn, E::add(E::add(k.gb, h1x), E::mul(q.beta, h1n))), E::add(E::add(k.gb, h2x), E::mul(q.beta, h2n)));
        t = E::add(t, E::mul(ap, E::mul(E::sub(x, q.omega_inv), E::sub(a, b))));
    }
    E res = E::add(E::mul(t, q.zh_inv[row]), t2);
    if (q.acc_mode) res = E::add(ldf(q.out + i), E::mul(q.scale, res));
    stf(q.out + i, res);
}
// WrongQuotientPolyDegree (prover.rs:916-919): coefficient `deg` must be non-zero, all above zero
template <class F> __global__ void degree_check_kernel(const Fp<F> *c, size_t deg, size_t m, int *err) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x + deg;
    if (i >= m) return;
    const bool z = ldf(c + i).is_zero();
    if ((i == deg) == z) *err = JF_ERR_QUOTIENT_DEGREE;
}
// Quotient coefficients from its interpolants on the sub-cosets: with t = sum_k X^(k n) t_k (deg t_k < n) and
// X^n = c_r on coset r, the interpolant there is T_r = sum_k c_r^k t_k, so t_k[j] = sum_r Vinv[k][r] T_r[j] where
// V[r][k] = c_r^k is a SUB x SUB Vandermonde matrix (inverted once per proving key on the host).
// SUB = 6: 6 n points determine the TurboPlonk quotient (degree 5 n + 7) for every n >= 8; SUB = 7: 7 n points the UltraPlonk
// quotient (degree 6 n + 8) for every n >= 9 (both used from n = 16, which leaves coefficients above the degree to check)
static constexpr int SUB_MAX = 7;
template <class F, int SUB> struct SolveArgs {
    const Fp<F> *T;  // SUB rows of n
    Fp<F> *t;        // SUB * n coefficients
    uint32_t n;
    Fp<F> vinv[SUB][SUB];
};
template <class F, int SUB> __global__ void __launch_bounds__(128) subcoset_solve_kernel(const __grid_constant__ SolveArgs<F, SUB> a) {
    using E = Fp<F>;
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.n) return;
    E T[SUB];
#pragma unroll
    for (int r = 0; r < SUB; r++) T[r] = ldf(a.T + (size_t)r * a.n + j);
#pragma unroll
    for (int k = 0; k < SUB; k++) {
        E acc = E::mul(a.vinv[k][0], T[0]);
#pragma unroll
        for (int r = 1; r < SUB; r++) acc = E::add(acc, E::mul(a.vinv[k][r], T[r]));
        stf(a.t + (size_t)k * a.n + j, acc);
    }
}
// split_quotient_polynomial (prover.rs:902-960): part i = t[i (n+2) .. ] with the masking edits
template <class F, int NW>
__global__ void split_kernel(const Fp<F> *t, size_t n, size_t total_len, const Fp<F> *b, Fp<F> *parts, size_t stride) {
    const int part = blockIdx.y;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t start = (size_t)part * (n + 2);
    const size_t len = part < NW - 1 ? n + 2 : total_len - start;
    Fp<F> v = Fp<F>::zero();
    if (i < len) v = ldf(t + start + i);
    if (i == 0 && part > 0) v = Fp<F>::sub(v, ldf(b + part - 1));
    if (i == n + 2 && part < NW - 1) v = ldf(b + part);
    if (i < stride) stf(parts + part * stride + i, v);
}

// ---- Plookup (UltraPlonk) -----------------------------------------------------------------------------------
// range table column: i for i < range_size, else 0 (compute_range_table, constraint_system.rs:1423-1439), Montgomery form
template <class F> __global__ void range_table_kernel(Fp<F> *out, uint32_t n, uint32_t range_size) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) stf(out + i, i < range_size ? Fp<F>::from_u32(i) : Fp<F>::zero());
}
// merged table value and merged lookup-wire value of every gate (merged_table_value / merged_lookup_wire_value,
// constraint_system.rs:1441-1491).  lk: range | key | table dom sep | q dom sep | q_lookup evaluation columns (n each);
// wv: 6 x n wire values (0 key, 1-2 lookup values, 3-4 table values, 5 range wire)
template <class F>
__global__ void merged_values_kernel(const Fp<F> *lk, const Fp<F> *wv, Fp<F> tau, uint32_t n, Fp<F> *mt, Fp<F> *ml) {
    using E = Fp<F>;
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t N = n;
    const E ql = ldf(lk + 4 * N + i);
    stf(mt + i, merged5(tau, ldf(lk + i), ql, ldf(lk + 2 * N + i), ldf(lk + N + i), ldf(wv + 3 * N + i), ldf(wv + 4 * N + i)));
    stf(ml + i, merged5(tau, ldf(wv + 5 * N + i), ql, ldf(lk + 3 * N + i), ldf(wv + i), ldf(wv + N + i), ldf(wv + 2 * N + i)));
}
// numerators / denominators of the lookup product (compute_lookup_prod_polynomial, constraint_system.rs:1311-1368):
// a_j = (1+beta)(gamma + ml_j)(gamma(1+beta) + mt_j + beta mt_(j+1)),  b_j = (gb + s_j + beta s_(j+1))(gb + s_(n-1+j) + beta s_(n+j)), j < n - 2
template <class F>
__global__ void lookup_ab_kernel(const Fp<F> *mt, const Fp<F> *ml, const Fp<F> *sorted, Fp<F> beta, Fp<F> gamma, Fp<F> bp1, Fp<F> gb,
                                 uint32_t n, Fp<F> *a, Fp<F> *b) {
    using E = Fp<F>;
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    E av = E::one(), bv = E::one();
    if (j + 2 < n) {
        av = E::mul(E::mul(bp1, E::add(gamma, ldf(ml + j))), E::add(E::add(gb, ldf(mt + j)), E::mul(beta, ldf(mt + j + 1))));
        bv = E::mul(E::add(E::add(gb, ldf(sorted + j)), E::mul(beta, ldf(sorted + j + 1))),
                    E::add(E::add(gb, ldf(sorted + (size_t)n - 1 + j)), E::mul(beta, ldf(sorted + (size_t)n + j))));
    }
    stf(a + j, av);
    stf(b + j, bv);
}
// ---- sorted lookup vector (compute_lookup_sorted_vec_polynomials, constraint_system.rs:1370-1418) ---------------
// The reference walks the merged table in order and emits every entry once plus, at the FIRST row that holds a value, once per
// lookup of that value; a lookup value absent from the table leaves the vector short (its error).  On the device:
//   sv_insert : open-addressing set over the table VALUES; a slot holds the smallest row index with that value
//   sv_count  : every lookup finds its value's slot and bumps that row's count (a miss raises `flag`)
//   sv_block / sv_carry / sv_offsets : exclusive prefix sum of (1 + count) -> the row's first position in the vector
//   sv_expand : every output position finds its row by bisection of the offsets and copies that row's value
static constexpr uint32_t SV_EMPTY = 0xFFFFFFFFu;
static constexpr int SV_T = 256, SV_I = 4;  // rows per block of the prefix sum: SV_T * SV_I
template <class F> __device__ __forceinline__ uint32_t sv_hash(const Fp<F> &k) {
    uint32_t h = 0x811C9DC5u;
#pragma unroll
    for (int i = 0; i < F::N; i++) h = (h ^ k.v[i]) * 0x01000193u + (h >> 15);
    h ^= h >> 16;
    h *= 0x7FEB352Du;
    return h ^ (h >> 15);
}
template <class F> __device__ __forceinline__ bool sv_eq(const Fp<F> &a, const Fp<F> &b) {
    uint32_t d = 0;
#pragma unroll
    for (int i = 0; i < F::N; i++) d |= a.v[i] ^ b.v[i];
    return d == 0;
}
template <class F> __global__ void sv_insert_kernel(const Fp<F> *mt, uint32_t n, uint32_t *slots, uint32_t mask) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Fp<F> key = ldf(mt + i);
    uint32_t h = sv_hash(key) & mask;
    for (;;) {
        uint32_t cur = *(volatile uint32_t *)(slots + h);
        if (cur == SV_EMPTY) {
            cur = atomicCAS(slots + h, SV_EMPTY, i);
            if (cur == SV_EMPTY) return;
        }
        // an occupied slot keeps its value for good (only the row index may still fall)
        if (sv_eq(ldf(mt + cur), key)) {
            if (cur > i) atomicMin(slots + h, i);
            return;
        }
        h = (h + 1) & mask;
    }
}
template <class F>
__global__ void sv_count_kernel(const Fp<F> *mt, const Fp<F> *ml, uint32_t n_lookups, const uint32_t *slots, uint32_t mask, uint32_t *cnt,
                                uint32_t *flag) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t row = SV_EMPTY;
    if (i < n_lookups) {
        const Fp<F> key = ldf(ml + i);
        uint32_t h = sv_hash(key) & mask;
        for (;;) {
            const uint32_t cur = slots[h];
            if (cur == SV_EMPTY) break;
            if (sv_eq(ldf(mt + cur), key)) {
                row = cur;
                break;
            }
            h = (h + 1) & mask;
        }
        if (row == SV_EMPTY) *flag = 1;
    }
    // circuits look the same few values up over and over: one atomic per distinct row and warp
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, row);
    if (row != SV_EMPTY && (threadIdx.x & 31) == (uint32_t)(__ffs(peers) - 1)) atomicAdd(cnt + row, (uint32_t)__popc(peers));
}
__global__ void __launch_bounds__(SV_T) sv_block_kernel(const uint32_t *cnt, uint32_t n, uint32_t *bsum) {
    __shared__ uint32_t sh[SV_T / 32];
    const uint32_t base = (blockIdx.x * SV_T + threadIdx.x) * SV_I;
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < SV_I; k++)
        if (base + k < n) v += 1 + cnt[base + k];
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < SV_T / 32; w++) t += sh[w];
        bsum[blockIdx.x] = t;
    }
}
// exclusive prefix sum of the block totals, in place, by one block; bsum[blocks] = the length of the vector
__global__ void __launch_bounds__(1024) sv_carry_kernel(uint32_t *bsum, uint32_t blocks) {
    __shared__ uint32_t sh[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < blocks; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < blocks ? bsum[i] : 0;
        uint32_t x = v;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if ((threadIdx.x & 31) >= (uint32_t)o) x += y;
        }
        if ((threadIdx.x & 31) == 31) sh[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = sh[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, w, o);
                if (threadIdx.x >= (uint32_t)o) w += y;
            }
            sh[threadIdx.x] = w;
        }
        __syncthreads();
        const uint32_t before = carry + (threadIdx.x >= 32 ? sh[(threadIdx.x >> 5) - 1] : 0) + x - v;
        if (i < blocks) bsum[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) bsum[blocks] = carry;
}
__global__ void __launch_bounds__(SV_T) sv_offsets_kernel(const uint32_t *cnt, uint32_t n, const uint32_t *bsum, uint32_t *off) {
    __shared__ uint32_t sh[SV_T / 32];
    const uint32_t base = (blockIdx.x * SV_T + threadIdx.x) * SV_I;
    uint32_t r[SV_I], v = 0;
#pragma unroll
    for (int k = 0; k < SV_I; k++) {
        r[k] = base + k < n ? 1 + cnt[base + k] : 0;
        v += r[k];
    }
    uint32_t x = v;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if ((threadIdx.x & 31) >= (uint32_t)o) x += y;
    }
    if ((threadIdx.x & 31) == 31) sh[threadIdx.x >> 5] = x;
    __syncthreads();
    uint32_t before = bsum[blockIdx.x] + x - v;
    for (uint32_t w = 0; w < (threadIdx.x >> 5); w++) before += sh[w];
#pragma unroll
    for (int k = 0; k < SV_I; k++) {
        if (base + k < n) off[base + k] = before;
        before += r[k];
    }
}
template <class F> __global__ void sv_expand_kernel(const Fp<F> *mt, const uint32_t *off, uint32_t n, uint32_t len, Fp<F> *out) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= len) return;
    uint32_t lo = 0, hi = n;  // the last row with off[row] <= p
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (off[mid] <= p) lo = mid;
        else hi = mid;
    }
    stf(out + p, ldf(mt + lo));
}
template <class F> __global__ void set_one_kernel(Fp<F> *p) {
    if (threadIdx.x == 0 && blockIdx.x == 0) stf(p, Fp<F>::one());
}

// ---- round 4: evaluation -------------------------------------------------------------------------------
static constexpr int EV_T = 256, EV_I = 8;
template <class F> __global__ void __launch_bounds__(EV_T) eval_partial_kernel(const Fp<F> *c, size_t len, Fp<F> x, Fp<F> *partials) {
    using E = Fp<F>;
    __shared__ E sh[EV_T];
    const size_t start = ((size_t)blockIdx.x * EV_T + threadIdx.x) * EV_I;
    E acc = E::zero();
    if (start < len) {
        const int cnt = (int)(len - start < (size_t)EV_I ? len - start : EV_I);
        for (int k = cnt - 1; k >= 0; k--) acc = E::add(E::mul(acc, x), ldf(c + start + k));
        acc = E::mul(acc, pow_small(x, start));
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int d = EV_T / 2; d >= 1; d >>= 1) {
        if ((int)threadIdx.x < d) sh[threadIdx.x] = E::add(sh[threadIdx.x], sh[threadIdx.x + d]);
        __syncthreads();
    }
    if (threadIdx.x == 0) stf(partials + blockIdx.x, sh[0]);
}
template <class F> __global__ void __launch_bounds__(EV_T) sum_kernel(const Fp<F> *in, size_t count, Fp<F> *out) {
    using E = Fp<F>;
    __shared__ E sh[EV_T];
    E acc = E::zero();
    for (size_t i = threadIdx.x; i < count; i += EV_T) acc = E::add(acc, ldf(in + i));
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int d = EV_T / 2; d >= 1; d >>= 1) {
        if ((int)threadIdx.x < d) sh[threadIdx.x] = E::add(sh[threadIdx.x], sh[threadIdx.x + d]);
        __syncthreads();
    }
    if (threadIdx.x == 0) stf(out, sh[0]);
}

// ---- round 5: linear combination, division by a linear factor ------------------------------------------
static constexpr int LC_MAX = 48;
template <class F> struct LinArgs {
    const Fp<F> *p[LC_MAX];
    uint32_t len[LC_MAX];
    Fp<F> s[LC_MAX];
    int count;
    int accumulate;  // != 0: out += (a batch proof has more (scalar, polynomial) pairs than one launch carries)
    Fp<F> *out;
    uint32_t out_len;
};
template <class F> __global__ void lincomb_kernel(const __grid_constant__ LinArgs<F> a) {
    using E = Fp<F>;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.out_len) return;
    E acc = a.accumulate ? ldf(a.out + i) : E::zero();
    for (int k = 0; k < a.count; k++)
        if (i < a.len[k]) acc = E::add(acc, E::mul(a.s[k], ldf(a.p[k] + i)));
    stf(a.out + i, acc);
}
// out[j] = in[j + shift] * base^(j + e0), j < len
static constexpr int MP_I = 8;
template <class F>
__global__ void mulpow_kernel(const Fp<F> *in, Fp<F> *out, size_t len, uint32_t shift, uint32_t e0, Fp<F> base) {
    using E = Fp<F>;
    const size_t start = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * MP_I;
    if (start >= len) return;
    E pw = pow_small(base, start + e0);
    for (int k = 0; k < MP_I && start + k < len; k++) {
        stf(out + start + k, E::mul(ldf(in + start + k + shift), pw));
        pw = E::mul(pw, base);
    }
}

// ---- proof linking (plonk/src/proof_system/proof_linking.rs) ---------------------------------------------
// out[j] = sum_k in[j + k N] c^k, j < N: `in` (len coefficients) reduced mod X^N - c
template <class F> __global__ void fold_kernel(const Fp<F> *in, size_t len, size_t N, Fp<F> c, Fp<F> *out) {
    using E = Fp<F>;
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    E acc = E::zero();
    for (size_t k = (len + N - 1) / N; k-- > 0;) {
        const size_t idx = j + k * N;
        acc = E::mul(acc, c);
        if (idx < len) acc = E::add(acc, ldf(in + idx));
    }
    stf(out + j, acc);
}
// vals: a polynomial's values on the N-th roots of unity (natural order).  *flag = 1 unless it vanishes on the linking domain
// {g^(offset + i), i < size}, g the 2^alignment-th root of unity = w_N^stride.
template <class F>
__global__ void link_roots_check_kernel(const Fp<F> *vals, uint32_t stride, uint32_t mask, uint32_t offset, uint32_t size, int *flag) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= size) return;
    if (!ldf(vals + (size_t)((offset + i) & mask) * stride).is_zero()) *flag = 1;
}
// rem = p mod z for a MONIC z of degree s and a p of m coefficients held in shared memory: schoolbook, one CTA, m - s dependent
// steps of s independent products each.  (Proof linking: p = (a1 - a2) mod (X^(2^a) - 1), z = Z_D, which divides X^(2^a) - 1.)
template <class F> __global__ void __launch_bounds__(1024) poly_mod_monic_kernel(const Fp<F> *p, uint32_t m, const Fp<F> *z, uint32_t s, Fp<F> *rem) {
    extern __shared__ uint4 poly_mod_smem[];
    Fp<F> *r = reinterpret_cast<Fp<F> *>(poly_mod_smem);
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) r[i] = ldf(p + i);
    __syncthreads();
    for (uint32_t k = m; k-- > s;) {
        const Fp<F> c = r[k];  // not written in this step
        if (!c.is_zero())
            for (uint32_t j = threadIdx.x; j < s; j += blockDim.x) r[k - s + j] = Fp<F>::sub(r[k - s + j], Fp<F>::mul(c, ldf(z + j)));
        __syncthreads();
    }
    for (uint32_t i = threadIdx.x; i < s; i += blockDim.x) stf(rem + i, r[i]);
}
// a[j] -= b[j], j < n
template <class F> __global__ void vsub_kernel(Fp<F> *a, const Fp<F> *b, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) stf(a + i, Fp<F>::sub(ldf(a + i), ldf(b + i)));
}
template <class F> __global__ void vmul_kernel(const Fp<F> *a, const Fp<F> *b, Fp<F> *out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) stf(out + i, Fp<F>::mul(ldf(a + i), ldf(b + i)));
}

}  // namespace jf

// ================================================================================================
// host side
// ================================================================================================
using namespace jf;

struct jf_plonk_pk {
    int curve = JF_BN254;
    int nw = jf::NW_TURBO;  // 5: TurboPlonk, 6: UltraPlonk (jf_ultraplonk_preprocess)
    // UltraPlonk only: the four Plookup table polynomials (range, key, table dom sep, q dom sep; 4 x n coefficients), their
    // evaluation columns (+ q_lookup's: 5 x n values), and the per-proof h1 / h2 / lookup-product polynomials (3 x np)
    unsigned range_bit_len = 0;
    void *d_lk = nullptr, *d_lk_evals = nullptr, *d_hp = nullptr;
    void *d_mt = nullptr, *d_ml = nullptr, *d_sorted = nullptr;  // merged table (n), merged lookup values (n), sorted vector (2n)
    uint32_t *d_sv = nullptr;  // scratch of the sorted vector: flag | counts (n) | offsets (n) | block sums | slots (sv_slots)
    uint32_t sv_slots = 0;     // power of two >= 4 n
    const jf_srs *srs = nullptr;
    unsigned log_n = 0, log_m = 0;
    size_t n = 0, m = 0, np = 0;  // np = n + PAD: stride of the n-sized polynomial buffers
    int sub = 0;                  // 6 | 7: the quotient is evaluated on that many of the 8 sub-cosets (n points each); 0: all 8n points
    size_t mq = 0;                // evaluation points per polynomial: sub ? sub * n : m
    uint64_t sub_off[jf::SUB_MAX * 4];               // offsets g w_8n^r of the sub-cosets (Montgomery limbs)
    uint32_t vinv[jf::SUB_MAX * jf::SUB_MAX][8];     // inverse Vandermonde matrix of subcoset_solve_kernel, row-major sub x sub
    size_t num_vars = 0;
    uint32_t num_inputs = 0;
    int cache_coset = 0;
    // one proof on several GPUs (jf_plonk_pk_shard_commits): this rank commits coefficients [shard_start, shard_start + len(srs_slice))
    // against its slice of the key; the XYZZ partials of all ranks are exchanged before a commitment is normalised
    jf_comm *comm = nullptr;
    const jf_srs *srs_slice = nullptr;
    size_t shard_start = 0;
    void *d_parts = nullptr;  // 32 slots x nranks partials
    // ... and round 3 dealt out by sub-coset: this rank transforms / evaluates only rows_local of the `sub` rows (row r lives on
    // rank r mod nranks); the rows' interpolants are broadcast before the Vandermonde solve
    int shard_rows = 0, rows_local = 0;
    int row_map[jf::SUB_MAX] = {0};
    uint64_t sub_off_local[jf::SUB_MAX * 4];
    // flags & 8: the key in the Lagrange basis of this domain (+ the two masking points), and the scalar vectors of a round's wire
    // commitments: NW x (n + 2) = the wire's values on the domain, then its two masking scalars
    jf_srs *lag = nullptr;
    void *d_lagsc = nullptr;
    int proofs_done = 0;    // batch_prove calls that left their wire polynomials in d_w (jf_plonk_link_hint, proof linking)
    uint32_t zero_sel = 0;  // selectors that are identically zero (flags & 2)
    int skip_zero = 0;      // flags & 2: zero polynomials (such selectors; PI without public inputs) are not transformed
    std::vector<uint32_t> pub_vars;  // variable index of every public input, in io-gate order
    // device, persistent
    void *d_sel = nullptr, *d_sig = nullptr, *d_sig_evals = nullptr;  // NSEL x n, NW x n coefficients; NW x n values
    uint32_t *d_wire_vars = nullptr, *d_gate_ids = nullptr;
    void *d_xlo = nullptr, *d_xhi = nullptr, *d_inv_nx1 = nullptr;
    void *d_cached = nullptr;  // (NSEL + NW) x m coset evaluations when cache_coset
    int lo_bits = 0;
    // device, per-proof workspace
    void *d_wit = nullptr, *d_bl = nullptr, *d_wv = nullptr, *d_w = nullptr /* (NW+1) x np: wires + PI */, *d_z = nullptr;
    void *d_a = nullptr, *d_b = nullptr, *d_tmp = nullptr, *d_e = nullptr, *d_q = nullptr, *d_split = nullptr;
    void *d_bp = nullptr, *d_t = nullptr, *d_s = nullptr, *d_wz = nullptr, *d_small = nullptr, *d_res = nullptr;
    void *d_side = nullptr;  // scratch of the side stream's division (3 np + scan temporaries)
    // host constants (Montgomery)
    uint32_t k[jf::NW_MAX][8], omega_n[8], omega_n_inv[8], omega_m[8], gen[8], zh_inv[8][8];
    uint64_t gen_limbs[4];
    // verifying-key commitments (affine Montgomery x || y) and their transcript serialisation
    std::vector<uint64_t> vk_xy;
    std::vector<int> vk_inf;
    std::vector<uint8_t> vk_transcript_prefix;  // everything append_vk_and_pub_input adds before the public inputs
    std::vector<void *> allocs;
    // secondary stream: coset NTTs that do not depend on the next challenge run beside the commitments
    cudaStream_t side = nullptr;
    cudaEvent_t ev_main = nullptr, ev_side = nullptr;
};

namespace jf {

template <class C> struct HostCurve {
    using Fq = typename C::Fq;
    using Fr = typename C::Fr;
    using Er = Fp<Fr>;
    using Eq = Fp<Fq>;
    static constexpr int L = Fq::N / 2;  // u64 limbs per base-field element

    static Er fr_from_limbs(const uint64_t *l) {
        Er r;
        for (int i = 0; i < 4; i++) {
            r.v[2 * i] = (uint32_t)l[i];
            r.v[2 * i + 1] = (uint32_t)(l[i] >> 32);
        }
        return r;
    }
    static void fr_to_limbs(const Er &a, uint64_t *l) {
        for (int i = 0; i < 4; i++) l[i] = (uint64_t)a.v[2 * i] | ((uint64_t)a.v[2 * i + 1] << 32);
    }
    // canonical little-endian bytes of a Montgomery-form element (`serialize_compressed` of a field element)
    static void fr_bytes(const Er &a, uint8_t out[32]) {
        Er c = Er::from_mont(a);
        memcpy(out, c.v, 32);
    }
    // `from_le_bytes_mod_order` of up to 64 bytes -> Montgomery form
    static Er fr_from_le_bytes_mod_order(const uint8_t *b, size_t n) {
        uint8_t buf[64] = {0};
        memcpy(buf, b, n);
        Er chunk[2];
        for (int c = 0; c < 2; c++) {
            memcpy(chunk[c].v, buf + 32 * c, 32);
            for (int it = 0; it < 8; it++) {  // 2^256 < 8 p for every field here
                uint32_t t[8], borrow;
                chain_sub_p<Fr>(t, chunk[c].v, borrow);
                if (borrow) break;
                memcpy(chunk[c].v, t, 32);
            }
            chunk[c] = Er::to_mont(chunk[c]);
        }
        // value = lo + hi * 2^256; the element 2^256 has Montgomery form R^2 mod p
        return Er::add(chunk[0], Er::mul(chunk[1], Er::r_squared()));
    }
    // `serialize_compressed` of a G1 point (`to_bytes!`, utilities/src/macros.rs:13-18).
    // BN254: ark-ec's generic short-Weierstrass form -- x little-endian, bit 7 of the LAST byte = y > -y, bit 6 = infinity.
    // BLS12-381: ark-bls12-381 0.4.0 overrides `serialize_with_mode` for G1 with the ZCash / IETF form -- x BIG-endian, top
    // bits of the FIRST byte: 7 = compressed, 6 = infinity, 5 = y > -y (published vectors: tests/golden/constants.json).
    static constexpr bool ZCASH_G1 = Fq::N == 12;
    static void g1_bytes(const uint64_t *xy, int inf, uint8_t *out) {
        const int nb = 8 * L;
        memset(out, 0, nb);
        if (inf) {
            if (ZCASH_G1) out[0] = 0xC0;
            else out[nb - 1] |= 0x40;
            return;
        }
        Eq x, y;
        memcpy(x.v, xy, nb);
        memcpy(y.v, xy + L, nb);
        Eq xc = Eq::from_mont(x), yc = Eq::from_mont(y), ny = Eq::from_mont(Eq::neg(y));
        bool greater = false;
        for (int i = Fq::N - 1; i >= 0; i--) {
            if (yc.v[i] != ny.v[i]) {
                greater = yc.v[i] > ny.v[i];
                break;
            }
        }
        if (ZCASH_G1) {
            const uint8_t *le = reinterpret_cast<const uint8_t *>(xc.v);
            for (int i = 0; i < nb; i++) out[i] = le[nb - 1 - i];
            out[0] |= greater ? 0xA0 : 0x80;
        } else {
            memcpy(out, xc.v, nb);
            if (greater) out[nb - 1] |= 0x80;
        }
    }
};

// frees everything a (possibly half-built) proving key owns; the caller has synchronised the streams that used it
static void release_pk(jf_plonk_pk *pk) {
    for (void *p : pk->allocs) cudaFree(p);
    if (pk->lag) {
        cudaFree(pk->lag->d_points);
        delete pk->lag;
    }
    if (pk->ev_main) cudaEventDestroy(pk->ev_main);
    if (pk->ev_side) cudaEventDestroy(pk->ev_side);
    if (pk->side) cudaStreamDestroy(pk->side);
    delete pk;
}

static int dalloc(jf_ctx *ctx, jf_plonk_pk *pk, size_t bytes, void **out) {
    JF_CUDA(ctx, cudaMalloc(out, bytes ? bytes : 64));
    pk->allocs.push_back(*out);
    return JF_OK;
}

template <class C, int NWT = NW_TURBO> struct Plonk {
    using Fr = typename C::Fr;
    using Fq = typename C::Fq;
    using E = Fp<Fr>;
    using H = HostCurve<C>;
    using D = Dim<NWT>;
    static constexpr int L = H::L;
    static constexpr size_t PT = sizeof(XYZZ<Fq>);
    static constexpr int NW = D::NW, NSEL = D::NSEL, NBLIND = D::NBLIND;
    static constexpr bool ULTRA = D::ULTRA;
    static constexpr int NVK = NSEL + NW + (ULTRA ? 4 : 0);  // verifying-key commitments
    static constexpr int NROWS = NSEL + 2 * NW + 2 + (ULTRA ? 7 : 0);  // coset-evaluation rows of round 3
    static constexpr int SUB = ULTRA ? 7 : 6;  // sub-cosets of n points that determine the quotient (degree NW (n + 1) + 2)

    static E kf(const jf_plonk_pk *pk, int j) {
        E r;
        memcpy(r.v, pk->k[j], 32);
        return r;
    }
    static E lf(const uint32_t *w) {
        E r;
        memcpy(r.v, w, 32);
        return r;
    }

    // d_out[n.. ] untouched; in-place inverse NTT over the size-n domain of `batch` vectors `stride` apart
    static int intt_n(jf_ctx *ctx, const jf_plonk_pk *pk, void *d, size_t batch, size_t stride) {
        return ntt_run(ctx, C::FR_ID, d, d, pk->n, pk->log_n, 1, nullptr, batch, stride);
    }

    // commitments of `count` device polynomials (Montgomery coefficients) -> d_res[slot..]
    static int commit_dev(jf_ctx *ctx, const jf_plonk_pk *pk, const void *d_poly, size_t len, int slot) {
        if (pk->comm) {  // this rank's slice of the sum; fetch_commits exchanges the partials
            const size_t a = pk->shard_start;
            return msm_run(ctx, pk->srs_slice, 0, (const char *)d_poly + sizeof(E) * std::min(a, len), len > a ? len - a : 0, 1,
                           (char *)pk->d_res + PT * slot);
        }
        return msm_run(ctx, pk->srs, 0, d_poly, len, 1, (char *)pk->d_res + PT * slot);
    }
    // `count` commitments on the main stream: every MSM runs its bulk phases (digits, sort, bucket accumulation) into
    // its own bucket array, then ONE bucket reduction handles all of them (its ~20 dependent levels cost the same
    // latency for 1 or 18 bucket sets).
    struct CommitJob { const void *poly; size_t len; int slot; };
    static int commit_many(jf_ctx *ctx, jf_plonk_pk *pk, const CommitJob *jobs, int count) {
        MsmJob mj[NVK];
        if (pk->comm) {
            const size_t a = pk->shard_start;
            for (int i = 0; i < count; i++)
                mj[i] = MsmJob{0, (const char *)jobs[i].poly + sizeof(E) * std::min(a, jobs[i].len), jobs[i].len > a ? jobs[i].len - a : 0, 1,
                               (char *)pk->d_res + PT * jobs[i].slot};
            return msm_run_many(ctx, pk->srs_slice, mj, count);
        }
        for (int i = 0; i < count; i++) mj[i] = MsmJob{0, jobs[i].poly, jobs[i].len, 1, (char *)pk->d_res + PT * jobs[i].slot};
        return msm_run_many(ctx, pk->srs, mj, count);
    }
    // bring `count` XYZZ results back and normalise (into_affine)
    static int fetch_commits(jf_ctx *ctx, const jf_plonk_pk *pk, int slot, int count, uint64_t *xy, int *inf) {
        void *h;
        const int parts = pk->comm ? comm_size(pk->comm) : 1;
        JF_TRY(pinned(ctx, PT * parts * count + 64, &h));
        const char *src = (const char *)pk->d_res + PT * slot;
        if (pk->comm) {
            // always on the main stream, in program order: the mailbox protocol of the exchange needs one ordered stream
            for (int i = 0; i < count; i++)
                JF_TRY(comm_exchange_from(ctx, pk->comm, (const char *)pk->d_res + PT * (slot + i), PT,
                                          (char *)pk->d_parts + PT * parts * (slot + i)));
            src = (const char *)pk->d_parts + PT * parts * slot;
        }
        JF_CUDA(ctx, cudaMemcpyAsync(h, src, PT * parts * count, cudaMemcpyDeviceToHost, ctx->stream));
        int herr[2] = {0, 0};
        JF_CUDA(ctx, cudaMemcpyAsync(herr, ctx->d_err, sizeof herr, cudaMemcpyDeviceToHost, ctx->stream));
        JF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (herr[0] || herr[1]) {
            cudaMemsetAsync(ctx->d_err, 0, sizeof herr, ctx->stream);
            if (herr[1]) return fail(ctx, herr[1], "prove: WrongQuotientPolyDegree (the witness does not satisfy the circuit)");
            if (herr[0] == JF_ERR_COMM) return fail(ctx, herr[0], "prove: a peer did not deliver its partial commitment in time");
            return fail(ctx, herr[0], "prove: a scalar is not below the group order");
        }
        for (int i = 0; i < count; i++)
            JF_TRY(msm_finish_host(ctx, pk->curve, (const uint64_t *)((const char *)h + PT * parts * i), parts, xy + (size_t)2 * L * i, inf + i));
        return JF_OK;
    }

    static int eval_dev(jf_ctx *ctx, const jf_plonk_pk *pk, const E *poly, size_t len, const E &x, E *d_out) {
        return eval_dev(ctx, (E *)pk->d_tmp, poly, len, x, d_out);
    }
    // d_out = poly(x); `partials` holds >= len / 2048 + 1 elements
    static int eval_dev(jf_ctx *ctx, E *partials, const E *poly, size_t len, const E &x, E *d_out) {
        const size_t per = (size_t)EV_T * EV_I;
        const unsigned blocks = (unsigned)((len + per - 1) / per);
        JF_LAUNCH(ctx, "eval_partial", eval_partial_kernel<Fr><<<blocks ? blocks : 1, EV_T, 0, ctx->stream>>>(poly, len, x, partials));
        JF_LAUNCH(ctx, "eval_sum", sum_kernel<Fr><<<1, EV_T, 0, ctx->stream>>>(partials, blocks ? blocks : 1, d_out));
        return JF_OK;
    }

    // out (len-1 coefficients) = p / (X - z), remainder dropped
    static int div_linear_dev(jf_ctx *ctx, const jf_plonk_pk *pk, const E *p, size_t len, const E &z, E *out) {
        return div_linear_dev(ctx, (E *)pk->d_t, (E *)pk->d_s, (E *)pk->d_tmp, p, len, z, out);
    }
    // t, s: len elements each; tmp: len / 512 + 64 elements
    static int div_linear_dev(jf_ctx *ctx, E *t, E *s, E *tmp, const E *p, size_t len, const E &z, E *out) {
        if (len < 2) return fail(ctx, JF_ERR_INVALID_ARG, "degenerate polynomial: nothing to divide");
        if (z.is_zero()) {  // p / X: drop the constant term
            JF_CUDA(ctx, cudaMemcpyAsync(out, p + 1, sizeof(E) * (len - 1), cudaMemcpyDeviceToDevice, ctx->stream));
            return JF_OK;
        }
        const unsigned b1 = (unsigned)((len + 256 * MP_I - 1) / (256 * MP_I));
        JF_LAUNCH(ctx, "mulpow", mulpow_kernel<Fr><<<b1, 256, 0, ctx->stream>>>(p, t, len, 0, 0, z));
        JF_TRY((fscan<Fr, OpAdd, true>(ctx, t, s, len, tmp)));
        const E zinv = E::inv(z);
        JF_LAUNCH(ctx, "mulpow", mulpow_kernel<Fr><<<b1, 256, 0, ctx->stream>>>(s, out, len - 1, 1, 1, zinv));
        return JF_OK;
    }

    static void tr_fr(Transcript &tr, const char *label, const E &a) {
        uint8_t b[32];
        H::fr_bytes(a, b);
        tr.append_message(label, b, 32);
    }
    static void tr_g1(Transcript &tr, const char *label, const uint64_t *xy, int inf) {
        uint8_t b[8 * L];
        H::g1_bytes(xy, inf, b);
        tr.append_message(label, b, 8 * L);
    }
    static E challenge(Transcript &tr, const char *label) {
        uint8_t buf[64];
        size_t n = tr.challenge_bytes(label, buf);
        E c = H::fr_from_le_bytes_mod_order(buf, n);
        uint8_t cb[32];
        H::fr_bytes(c, cb);
        tr.challenge_done(label, cb, 32);
        return c;
    }

    // ------------------------------------------------------------------------------------------
    // the three per-gate Plookup columns of an UltraPlonk circuit (null for TurboPlonk)
    struct LookupCols { unsigned range_bit_len; const uint64_t *table_key, *table_dom_sep, *q_dom_sep; };
    static int preprocess(jf_ctx *ctx, const jf_srs *srs, unsigned log_n, const uint64_t *selector_evals,
                          const uint64_t *sigma_evals, const uint64_t *k, const uint32_t *wire_variables, size_t num_vars,
                          const uint32_t *pub_gate_ids, size_t num_inputs, int flags, jf_plonk_pk **out,
                          const LookupCols *lc = nullptr) {
        // n = 2: the TurboPlonk quotient (degree 17) does not fit 8n; UltraPlonk: 6 n + 9 coefficients need n >= 8 to fit 8n
        if (log_n < (ULTRA ? 3u : 2u) || log_n + 3 > (unsigned)Fr::TWO_ADICITY || log_n > 26)
            return fail(ctx, JF_ERR_DOMAIN_TOO_LARGE, "preprocess: unsupported domain size");
        if (ULTRA && (!lc || lc->range_bit_len > log_n))
            return fail(ctx, JF_ERR_INVALID_ARG, "preprocess: Domain size < range size (constraint_system.rs:1428-1434)");
        const size_t n = (size_t)1 << log_n, m = n * 8, np = n + PAD;
        if (srs->n < n + 3) return fail(ctx, JF_ERR_INVALID_ARG, "preprocess: the commit key needs n + 3 points (srs_size = n + 2)");
        jf_plonk_pk *pk = new jf_plonk_pk();
        pk->curve = srs->curve;
        pk->nw = NW;
        pk->range_bit_len = lc ? lc->range_bit_len : 0;
        pk->srs = srs;
        pk->log_n = log_n;
        pk->log_m = log_n + 3;
        pk->n = n;
        pk->m = m;
        pk->np = np;
        pk->num_vars = num_vars;
        pk->num_inputs = (uint32_t)num_inputs;
        pk->cache_coset = flags & 1;
        pk->skip_zero = (flags & 2) ? 1 : 0;
        // 6 n >= 5 n + 8 coefficients needs n >= 8; n >= 16 also leaves n - 8 >= 8 coefficients above the quotient's degree
        // for the WrongQuotientPolyDegree check (at n = 8 the six-row interpolant has no coefficient above degree 47 at all)
        // (UltraPlonk: 7 n >= 6 n + 9 coefficients needs n >= 9; n >= 16 leaves n - 9 >= 7 coefficients above the degree)
        pk->sub = (log_n >= 4 && !(flags & 4)) ? SUB : 0;
        pk->mq = pk->sub ? (size_t)pk->sub * n : m;
        if (flags & 2) {
            for (int sel = 0; sel < NSEL; sel++) {
                const uint64_t *col = selector_evals + (size_t)sel * 4 * n;
                uint64_t acc = 0;
                for (size_t i = 0; i < 4 * n; i++) acc |= col[i];
                if (!acc) pk->zero_sel |= 1u << sel;
            }
            pk->zero_sel &= 0x1FFFu;  // q_lookup (selector 13) is read unconditionally by the Plookup terms
        }
        int prio_least = 0, prio_greatest = 0;
        cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
        if (cudaStreamCreateWithPriority(&pk->side, cudaStreamNonBlocking, prio_least) != cudaSuccess ||
            cudaEventCreateWithFlags(&pk->ev_main, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&pk->ev_side, cudaEventDisableTiming) != cudaSuccess) {
            release_pk(pk);
            return fail(ctx, JF_ERR_CUDA, "preprocess: cannot create the side stream");
        }
        int rc = preprocess_inner(ctx, pk, selector_evals, sigma_evals, k, wire_variables, pub_gate_ids, lc);
        if (rc == JF_OK && (flags & 8)) {  // Lagrange-basis key for the wire commitments (lagrange.cu), once per proving key
            rc = srs_lagrange(ctx, srs, log_n, 1, &pk->lag);
            if (rc == JF_OK) rc = dalloc(ctx, pk, sizeof(E) * NW * (n + 2), &pk->d_lagsc);
        }
        if (rc != JF_OK) {
            cudaStreamSynchronize(ctx->stream);
            cudaStreamSynchronize(pk->side);
            release_pk(pk);
            return rc;
        }
        *out = pk;
        return JF_OK;
    }

    static int preprocess_inner(jf_ctx *ctx, jf_plonk_pk *pk, const uint64_t *selector_evals, const uint64_t *sigma_evals,
                                const uint64_t *k, const uint32_t *wire_variables, const uint32_t *pub_gate_ids, const LookupCols *lc) {
        const size_t n = pk->n, m = pk->m, np = pk->np, mq = pk->mq;
        cudaStream_t st = ctx->stream;
        for (size_t i = 0; i < (size_t)NW * n; i++)
            if (wire_variables[i] >= pk->num_vars) return fail(ctx, JF_ERR_INVALID_ARG, "preprocess: wire variable out of range");
        for (size_t i = 0; i < pk->num_inputs; i++) {
            if (pub_gate_ids[i] >= n) return fail(ctx, JF_ERR_INVALID_ARG, "preprocess: io gate id out of range");
            pk->pub_vars.push_back(wire_variables[(size_t)4 * n + pub_gate_ids[i]]);  // output wire (GATE_WIDTH)
        }
        // ---- host constants ----
        for (int j = 0; j < NW; j++) {
            E kj = H::fr_from_limbs(k + 4 * j);
            memcpy(pk->k[j], kj.v, 32);
        }
        E g = E::from_u32(Fr::GENERATOR);
        memcpy(pk->gen, g.v, 32);
        H::fr_to_limbs(g, pk->gen_limbs);
        // two-adic root: GENERATOR^((p-1)/2^s), then squared down to the n- and m-th roots
        uint32_t e[8];
        Limbs<Fr>::p(e);
        e[0] -= 1;
        for (int s = 0; s < Fr::TWO_ADICITY; s++)
            for (int i = 0; i < 8; i++) e[i] = (e[i] >> 1) | (i < 7 ? e[i + 1] << 31 : 0);
        E root = E::pow(g, e, 8);
        E wm = root;
        for (unsigned s = 0; s < Fr::TWO_ADICITY - pk->log_m; s++) wm = E::sqr(wm);
        E wn = wm;
        for (int s = 0; s < 3; s++) wn = E::sqr(wn);
        memcpy(pk->omega_m, wm.v, 32);
        memcpy(pk->omega_n, wn.v, 32);
        {
            const E wni = E::inv(wn);
            memcpy(pk->omega_n_inv, wni.v, 32);
        }
        for (int r = 0; r < 8; r++) {  // 1 / ((g w_m^r)^n - 1)   (prover.rs:530-537)
            E x = E::mul(g, pow_small(wm, r));
            E zh = E::sub(pow_small(x, n), E::one());
            zh = E::inv(zh);
            memcpy(pk->zh_inv[r], zh.v, 32);
        }
        if (pk->sub) {
            // sub-coset offsets and the inverse of V[r][k] = c_r^k, c_r = (g w_m^r)^n, by Gauss-Jordan elimination
            E V[SUB][2 * SUB];
            for (int r = 0; r < SUB; r++) {
                E off = E::mul(g, pow_small(wm, r));
                H::fr_to_limbs(off, pk->sub_off + 4 * r);
                const E c = pow_small(off, n);
                E pw = E::one();
                for (int kk = 0; kk < SUB; kk++) {
                    V[r][kk] = pw;
                    pw = E::mul(pw, c);
                    V[r][SUB + kk] = kk == r ? E::one() : E::zero();
                }
            }
            for (int col = 0; col < SUB; col++) {
                int piv = col;
                while (piv < SUB && V[piv][col].is_zero()) piv++;
                if (piv == SUB) return fail(ctx, JF_ERR_INVALID_ARG, "preprocess: singular sub-coset matrix");
                for (int kk = 0; kk < 2 * SUB; kk++) std::swap(V[piv][kk], V[col][kk]);
                const E inv = E::inv(V[col][col]);
                for (int kk = 0; kk < 2 * SUB; kk++) V[col][kk] = E::mul(V[col][kk], inv);
                for (int r = 0; r < SUB; r++) {
                    if (r == col || V[r][col].is_zero()) continue;
                    const E f = V[r][col];
                    for (int kk = 0; kk < 2 * SUB; kk++) V[r][kk] = E::sub(V[r][kk], E::mul(f, V[col][kk]));
                }
            }
            for (int r = 0; r < SUB; r++)
                for (int kk = 0; kk < SUB; kk++) memcpy(pk->vinv[r * SUB + kk], V[r][SUB + kk].v, 32);
        }
        // ---- device buffers ----
        const size_t fe = sizeof(E);
        JF_TRY(dalloc(ctx, pk, fe * NSEL * n, &pk->d_sel));
        JF_TRY(dalloc(ctx, pk, fe * NW * n, &pk->d_sig));
        JF_TRY(dalloc(ctx, pk, fe * NW * n, &pk->d_sig_evals));
        JF_TRY(dalloc(ctx, pk, sizeof(uint32_t) * NW * n, (void **)&pk->d_wire_vars));
        JF_TRY(dalloc(ctx, pk, sizeof(uint32_t) * (pk->num_inputs + 1), (void **)&pk->d_gate_ids));
        pk->lo_bits = (pk->log_m + 1) / 2;
        const uint32_t lo_n = 1u << pk->lo_bits, hi_n = 1u << (pk->log_m - pk->lo_bits);
        JF_TRY(dalloc(ctx, pk, fe * lo_n, &pk->d_xlo));
        JF_TRY(dalloc(ctx, pk, fe * hi_n, &pk->d_xhi));
        JF_TRY(dalloc(ctx, pk, fe * mq, &pk->d_inv_nx1));
        JF_TRY(dalloc(ctx, pk, fe * (pk->num_vars + 1), &pk->d_wit));
        JF_TRY(dalloc(ctx, pk, fe * NBLIND, &pk->d_bl));
        JF_TRY(dalloc(ctx, pk, fe * NW * n, &pk->d_wv));
        JF_TRY(dalloc(ctx, pk, fe * (NW + 1) * np, &pk->d_w));
        JF_TRY(dalloc(ctx, pk, fe * np, &pk->d_z));
        JF_TRY(dalloc(ctx, pk, fe * m, &pk->d_a));   // scan inputs / outputs (m-sized for the inversion table)
        JF_TRY(dalloc(ctx, pk, fe * m, &pk->d_b));
        JF_TRY(dalloc(ctx, pk, fe * (m / 256 + 4096), &pk->d_tmp));
        const int ne = pk->cache_coset ? NW + 2 + (ULTRA ? 3 : 0) : NROWS;  // resident: selectors, sigmas [, the four table polynomials]
        JF_TRY(dalloc(ctx, pk, fe * (size_t)ne * mq, &pk->d_e));
        JF_TRY(dalloc(ctx, pk, fe * m, &pk->d_q));
        JF_TRY(dalloc(ctx, pk, fe * NW * np, &pk->d_split));
        JF_TRY(dalloc(ctx, pk, fe * np, &pk->d_bp));
        JF_TRY(dalloc(ctx, pk, fe * np, &pk->d_t));
        JF_TRY(dalloc(ctx, pk, fe * np, &pk->d_s));
        JF_TRY(dalloc(ctx, pk, fe * np, &pk->d_wz));
        JF_TRY(dalloc(ctx, pk, fe * (3 * np + np / 256 + 4096), &pk->d_side));
        JF_TRY(dalloc(ctx, pk, fe * 64, &pk->d_small));
        JF_TRY(dalloc(ctx, pk, PT * 32, &pk->d_res));
        if (ULTRA) {
            JF_TRY(dalloc(ctx, pk, fe * 4 * n, &pk->d_lk));
            JF_TRY(dalloc(ctx, pk, fe * 5 * n, &pk->d_lk_evals));
            JF_TRY(dalloc(ctx, pk, fe * 3 * np, &pk->d_hp));
            JF_TRY(dalloc(ctx, pk, fe * n, &pk->d_mt));
            JF_TRY(dalloc(ctx, pk, fe * n, &pk->d_ml));
            JF_TRY(dalloc(ctx, pk, fe * 2 * n, &pk->d_sorted));
            pk->sv_slots = (uint32_t)(4 * n);
            JF_TRY(dalloc(ctx, pk, sizeof(uint32_t) * (16 + 2 * n + (n / (SV_T * SV_I) + 2) + pk->sv_slots), (void **)&pk->d_sv));
        }
        if (pk->cache_coset) JF_TRY(dalloc(ctx, pk, fe * (size_t)(NSEL + NW + (ULTRA ? 4 : 0)) * mq, &pk->d_cached));
        JF_CUDA(ctx, cudaMemcpyAsync(pk->d_sel, selector_evals, fe * NSEL * n, cudaMemcpyHostToDevice, st));
        JF_CUDA(ctx, cudaMemcpyAsync(pk->d_sig, sigma_evals, fe * NW * n, cudaMemcpyHostToDevice, st));
        JF_CUDA(ctx, cudaMemcpyAsync(pk->d_sig_evals, sigma_evals, fe * NW * n, cudaMemcpyHostToDevice, st));
        JF_CUDA(ctx, cudaMemcpyAsync(pk->d_wire_vars, wire_variables, sizeof(uint32_t) * NW * n, cudaMemcpyHostToDevice, st));
        if (pk->num_inputs)
            JF_CUDA(ctx, cudaMemcpyAsync(pk->d_gate_ids, pub_gate_ids, sizeof(uint32_t) * pk->num_inputs, cudaMemcpyHostToDevice, st));
        // x_i = g w_m^i tables and 1 / (n (x_i - 1)) by one batch inversion
        JF_LAUNCH(ctx, "pow_tab", pow_tab_kernel<Fr><<<(lo_n + 127) / 128, 128, 0, st>>>((E *)pk->d_xlo, wm, E::one(), 1, lo_n));
        JF_LAUNCH(ctx, "pow_tab", pow_tab_kernel<Fr><<<(hi_n + 127) / 128, 128, 0, st>>>((E *)pk->d_xhi, wm, g, lo_n, hi_n));
        {
            E nf = E::from_u32((uint32_t)n);
            E *d = (E *)pk->d_q, *pp = (E *)pk->d_a, *sp = (E *)pk->d_b, *small = (E *)pk->d_small;
            JF_LAUNCH(ctx, "xm1", xm1_kernel<Fr><<<(unsigned)((mq + 255) / 256), 256, 0, st>>>(
                (const E *)pk->d_xlo, (const E *)pk->d_xhi, pk->lo_bits, nf, d, mq, pk->sub, (uint32_t)pk->log_n));
            JF_TRY((fscan<Fr, OpMul, false>(ctx, d, pp, mq, (E *)pk->d_tmp)));
            JF_TRY((fscan<Fr, OpMul, true>(ctx, d, sp, mq, (E *)pk->d_tmp)));
            JF_LAUNCH(ctx, "inv_one", inv_one_kernel<Fr><<<1, 32, 0, st>>>(sp, small));
            JF_LAUNCH(ctx, "inv_combine", inv_combine_kernel<Fr><<<(unsigned)((mq + 255) / 256), 256, 0, st>>>(pp, sp, small, (E *)pk->d_inv_nx1, mq));
        }
        if (ULTRA) {
            // Plookup columns: evaluation copies (range | key | table dom sep | q dom sep | q_lookup) for round 1.5, and the four
            // table polynomials (snark.rs:541-556; constraint_system.rs:1263-1288)
            E *lke = (E *)pk->d_lk_evals;
            JF_LAUNCH(ctx, "range_table", range_table_kernel<Fr><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(lke, (uint32_t)n, 1u << lc->range_bit_len));
            JF_CUDA(ctx, cudaMemcpyAsync(lke + n, lc->table_key, fe * n, cudaMemcpyHostToDevice, st));
            JF_CUDA(ctx, cudaMemcpyAsync(lke + 2 * n, lc->table_dom_sep, fe * n, cudaMemcpyHostToDevice, st));
            JF_CUDA(ctx, cudaMemcpyAsync(lke + 3 * n, lc->q_dom_sep, fe * n, cudaMemcpyHostToDevice, st));
            JF_CUDA(ctx, cudaMemcpyAsync(lke + 4 * n, (const E *)pk->d_sel + (size_t)(NSEL - 1) * n, fe * n, cudaMemcpyDeviceToDevice, st));
            JF_CUDA(ctx, cudaMemcpyAsync(pk->d_lk, lke, fe * 4 * n, cudaMemcpyDeviceToDevice, st));
            JF_TRY(intt_n(ctx, pk, pk->d_lk, 4, n));
        }
        // selector / sigma polynomials (ifft) and the verifying-key commitments (18; UltraPlonk: 14 + 6 + 4)
        JF_TRY(intt_n(ctx, pk, pk->d_sel, NSEL, n));
        JF_TRY(intt_n(ctx, pk, pk->d_sig, NW, n));
        {
            CommitJob jobs[NVK];
            for (int i = 0; i < NSEL; i++) jobs[i] = {(E *)pk->d_sel + (size_t)i * n, n, i};
            for (int i = 0; i < NW; i++) jobs[NSEL + i] = {(E *)pk->d_sig + (size_t)i * n, n, NSEL + i};
            for (int i = NSEL + NW; i < NVK; i++) jobs[i] = {(E *)pk->d_lk + (size_t)(i - NSEL - NW) * n, n, i};
            JF_TRY(commit_many(ctx, pk, jobs, NVK));
        }
        pk->vk_xy.resize((size_t)NVK * 2 * L);
        pk->vk_inf.resize(NVK);
        JF_TRY(fetch_commits(ctx, pk, 0, NVK, pk->vk_xy.data(), pk->vk_inf.data()));
        if (pk->cache_coset) {
            JF_TRY(coset_fft_rows(ctx, pk, (const E *)pk->d_sel, n, n, NSEL, (E *)pk->d_cached, pk->zero_sel));
            JF_TRY(coset_fft_rows(ctx, pk, (const E *)pk->d_sig, n, n, NW, (E *)pk->d_cached + (size_t)NSEL * mq));
            if (ULTRA) JF_TRY(coset_fft_rows(ctx, pk, (const E *)pk->d_lk, n, n, 4, (E *)pk->d_cached + (size_t)(NSEL + NW) * mq));
        }
        // transcript prefix (transcript/mod.rs:45-88): sizes, k, selector and sigma commitments
        JF_CUDA(ctx, cudaStreamSynchronize(st));
        return JF_OK;
    }

    // sub-coset rows this GPU holds (all of them unless round 3 is dealt out over several GPUs)
    static int rows_eff(const jf_plonk_pk *pk) { return pk->shard_rows ? pk->rows_local : pk->sub; }
    static size_t mq_eff(const jf_plonk_pk *pk) { return pk->shard_rows ? (size_t)pk->rows_local * pk->n : pk->mq; }
    static const uint64_t *offs_eff(const jf_plonk_pk *pk) { return pk->shard_rows ? pk->sub_off_local : pk->sub_off; }

    // rows of coefficients (`len` valid, `stride` apart) -> rows of m coset evaluations in `dst`
    // Rows whose bit is set in `skip` are zero polynomials: their (all-zero) evaluations are never read.
    static int coset_fft_rows(jf_ctx *ctx, const jf_plonk_pk *pk, const E *src, size_t stride, size_t len, int rows, E *dst,
                              uint32_t skip = 0) {
        const size_t m = pk->m, in_len = pk->n + 3;
        if (pk->sub) {
            // sub-coset form: row r of the result holds the polynomial on the cosets (g w_m^s) <w_n>, s < sub, n points each;
            // the transform reads the coefficients where they are (reduced mod X^n - c_s on the way in)
            if (rows_eff(pk) == 0) return JF_OK;  // round 3 dealt out over more GPUs than there are sub-cosets: nothing here
            int r = 0;
            while (r < rows) {
                if ((skip >> r) & 1u) {
                    r++;
                    continue;
                }
                int b = 1;
                while (b < 5 && r + b < rows && !((skip >> (r + b)) & 1u)) b++;
                JF_TRY(ntt_run_cosets(ctx, C::FR_ID, src + (size_t)r * stride, stride, len, dst + (size_t)r * mq_eff(pk), pk->log_n, 0,
                                      offs_eff(pk), rows_eff(pk), b));
                r += b;
            }
            return JF_OK;
        }
        dim3 grid((unsigned)((in_len + 255) / 256), rows);
        JF_LAUNCH(ctx, "copy_rows", copy_rows_kernel<Fr><<<grid, 256, 0, ctx->stream>>>(src, stride, dst, m, len, in_len));
        int r = 0;
        while (r < rows) {
            if ((skip >> r) & 1u) {
                r++;
                continue;
            }
            int b = 1;
            while (b < 5 && r + b < rows && !((skip >> (r + b)) & 1u)) b++;
            JF_TRY(ntt_run(ctx, C::FR_ID, dst + (size_t)r * m, dst + (size_t)r * m, in_len, pk->log_m, 0, pk->gen_limbs, b, m));
            r += b;
        }
        return JF_OK;
    }

    // Issue `body` on the side stream, ordered after everything issued on the main stream so far.
    template <class Fn> static int on_side(jf_ctx *ctx, jf_plonk_pk *pk, Fn body) {
        cudaStream_t main_stream = ctx->stream;
        JF_CUDA(ctx, cudaEventRecord(pk->ev_main, main_stream));
        JF_CUDA(ctx, cudaStreamWaitEvent(pk->side, pk->ev_main, 0));
        ctx->stream = pk->side;
        ctx->lane = 1;
        int rc = body();
        ctx->stream = main_stream;
        ctx->lane = 0;
        return rc;
    }
    // the main stream waits for everything issued on the side stream so far
    static int join_side(jf_ctx *ctx, jf_plonk_pk *pk) {
        JF_CUDA(ctx, cudaEventRecord(pk->ev_side, pk->side));
        JF_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, pk->ev_side, 0));
        return JF_OK;
    }

    static void append_vk_and_pub_input(Transcript &tr, const jf_plonk_pk *pk, const E *pub, size_t npub) {
        // transcript/mod.rs:45-102: the PlookupVerifyingKey commitments are NOT part of it
        const uint32_t bits = Fr::BITS;
        const uint64_t dom = pk->n, nin = pk->num_inputs;
        tr.append_message("field size in bits", (const uint8_t *)&bits, 4);
        tr.append_message("domain size", (const uint8_t *)&dom, 8);
        tr.append_message("input size", (const uint8_t *)&nin, 8);
        for (int j = 0; j < NW; j++) tr_fr(tr, "wire subsets separators", kf(pk, j));
        for (int i = 0; i < NSEL; i++) tr_g1(tr, "selector commitments", pk->vk_xy.data() + (size_t)2 * L * i, pk->vk_inf[i]);
        for (int i = 0; i < NW; i++)
            tr_g1(tr, "sigma commitments", pk->vk_xy.data() + (size_t)2 * L * (NSEL + i), pk->vk_inf[NSEL + i]);
        for (size_t i = 0; i < npub; i++) tr_fr(tr, "public input", pub[i]);
    }

    // compute_lookup_sorted_vec_polynomials (constraint_system.rs:1370-1418): the lookup values merged into the table, in
    // table order.  The merge is keyed by field VALUE (a hash map in the reference); it runs on the host over the two
    // n-element vectors the device just computed -- 2 x 32 n bytes down, 64 n bytes up -- and is the only part of an
    // UltraPlonk proof that is not a device kernel.
    static int sorted_vector(jf_ctx *ctx, jf_plonk_pk *pk) {
        const size_t n = pk->n;
        cudaStream_t st = ctx->stream;
        const uint32_t blocks = (uint32_t)((n + SV_T * SV_I - 1) / (SV_T * SV_I)), mask = pk->sv_slots - 1, cap = (uint32_t)(2 * n - 1);
        uint32_t *flag = pk->d_sv, *cnt = flag + 16, *off = cnt + n, *bsum = off + n, *slots = bsum + blocks + 1;
        const E *mt = (const E *)pk->d_mt, *ml = (const E *)pk->d_ml;
        JF_CUDA(ctx, cudaMemsetAsync(flag, 0, sizeof(uint32_t) * (16 + n), st));
        JF_CUDA(ctx, cudaMemsetAsync(slots, 0xFF, sizeof(uint32_t) * pk->sv_slots, st));
        const unsigned g = (unsigned)((n + 255) / 256);
        JF_LAUNCH(ctx, "sv_insert", sv_insert_kernel<Fr><<<g, 256, 0, st>>>(mt, (uint32_t)n, slots, mask));
        // only the first n - 1 gates look up (the last slot never holds a lookup gate)
        JF_LAUNCH(ctx, "sv_count", sv_count_kernel<Fr><<<g, 256, 0, st>>>(mt, ml, (uint32_t)(n - 1), slots, mask, cnt, flag));
        JF_LAUNCH(ctx, "sv_block", sv_block_kernel<<<blocks, SV_T, 0, st>>>(cnt, (uint32_t)n, bsum));
        JF_LAUNCH(ctx, "sv_carry", sv_carry_kernel<<<1, 1024, 0, st>>>(bsum, blocks));
        JF_LAUNCH(ctx, "sv_offsets", sv_offsets_kernel<<<blocks, SV_T, 0, st>>>(cnt, (uint32_t)n, bsum, off));
        JF_LAUNCH(ctx, "sv_expand", sv_expand_kernel<Fr><<<(cap + 255) / 256, 256, 0, st>>>(mt, off, (uint32_t)n, cap, (E *)pk->d_sorted));
        void *h;
        JF_TRY(pinned(ctx, 64, &h));
        JF_CUDA(ctx, cudaMemcpyAsync(h, flag, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        JF_CUDA(ctx, cudaStreamSynchronize(st));
        if (*(const uint32_t *)h)
            return fail(ctx, JF_ERR_INVALID_ARG, "prove: The sorted vector has wrong length, some lookup variables might be outside the table");
        return JF_OK;
    }

    // ------------------------------------------------------------------------------------------
    // one (scalar, polynomial) pair of a linear combination; `lincomb_many` issues as many launches as the list needs
    struct Term { const E *p; size_t len; E s; };
    static int lincomb_many(jf_ctx *ctx, const std::vector<Term> &terms, E *out, size_t out_len) {
        size_t done = 0;
        do {
            LinArgs<Fr> la;
            const size_t cnt = std::min<size_t>(LC_MAX, terms.size() - done);
            for (size_t c = 0; c < cnt; c++) {
                la.p[c] = terms[done + c].p;
                la.len[c] = (uint32_t)terms[done + c].len;
                la.s[c] = terms[done + c].s;
            }
            la.count = (int)cnt;
            la.accumulate = done ? 1 : 0;
            la.out = out;
            la.out_len = (uint32_t)out_len;
            JF_LAUNCH(ctx, "lincomb", lincomb_kernel<Fr><<<(unsigned)((out_len + 127) / 128), 128, 0, ctx->stream>>>(la));
            done += cnt;
        } while (done < terms.size());
        return JF_OK;
    }

    template <class ProofT>
    static int prove(jf_ctx *ctx, jf_plonk_pk *pk, const uint64_t *witness, const uint64_t *blinders, int kind,
                     const uint8_t *extra, size_t extra_len, ProofT *out) {
        return batch_prove(ctx, &pk, 1, &witness, blinders, kind, extra, extra_len, out);
    }

    // `PlonkKzgSnark::batch_prove` -> batch_prove_internal (snark.rs:201-469): `count` instances over one domain and one commit
    // key share the transcript, ONE quotient polynomial (instance i enters with alpha_base_i = (alpha^3 | alpha^7)^i) and ONE pair
    // of opening proofs.  Every instance has its own proving key (and so its own device workspace).  outs[i] receives
    // instance i's commitments and evaluations; the shared parts (split quotient, opening proofs) are written to every entry.
    // blinders, in the order the reference's single prng is consumed: the wire masks of every instance, [the h1 / h2 masks,] the
    // z masks, [the lookup-product masks,] then the NW - 1 split-quotient randomizers.
    template <class ProofT>
    static int batch_prove(jf_ctx *ctx, jf_plonk_pk *const *pks, size_t count, const uint64_t *const *witnesses,
                           const uint64_t *blinders, int kind, const uint8_t *extra, size_t extra_len, ProofT *outs) {
        // The side-stream work of EVERY instance goes to the first key's side stream: the context keeps one set of scratch
        // buffers per lane (main / side), so two side streams working at once would share them.
        jf_plonk_pk *pk0 = pks[0];
        const size_t n = pk0->n, m = mq_eff(pk0), np = pk0->np;  // m: evaluation points per polynomial in round 3 (on this GPU)
        const size_t fe = sizeof(E);
        cudaStream_t st = ctx->stream;
        for (size_t i = 0; i < count; i++) {  // snark.rs:226-260
            const jf_plonk_pk *pk = pks[i];
            if (pk->nw != NW || pk->curve != pk0->curve) return fail(ctx, JF_ERR_INVALID_ARG, "batch_prove: inconsistent plonk circuit types");
            if (pk->n != n || pk->sub != pk0->sub) return fail(ctx, JF_ERR_INVALID_ARG, "batch_prove: proving key domain size differs from the expected domain size");
            if (pk->comm != pk0->comm || pk->shard_rows != pk0->shard_rows || pk->shard_start != pk0->shard_start || pk->srs_slice != pk0->srs_slice)
                return fail(ctx, JF_ERR_INVALID_ARG, "batch_prove: the instances must be sharded alike (jf_plonk_pk_shard_commits)");
            if (pk->srs != pk0->srs) return fail(ctx, JF_ERR_INVALID_ARG, "batch_prove: the instances must share one commit key");
            for (size_t j = 0; j < i; j++)
                if (pks[j] == pk) return fail(ctx, JF_ERR_INVALID_ARG, "batch_prove: every instance needs its own proving key (workspace)");
            memset(&outs[i], 0, sizeof outs[i]);
        }
        // blinders: one device copy for the whole batch
        constexpr int PER = NBLIND - (NW - 1);  // per-instance masks
        const size_t nbl = count * PER + (NW - 1);
        E *bl;
        {
            void *p;
            JF_TRY(scratch(ctx, "batch_blinders", fe * nbl, &p));
            bl = (E *)p;
        }
        JF_CUDA(ctx, cudaMemcpyAsync(bl, blinders, fe * nbl, cudaMemcpyHostToDevice, st));
        const E *bl_w = bl, *bl_h = bl + count * 2 * NW, *bl_z = bl_h + (ULTRA ? count * 6 : 0), *bl_pl = bl_z + count * 3,
                *bl_split = bl_pl + (ULTRA ? count * 3 : 0);
        // per-instance views
        struct Inst {
            jf_plonk_pk *pk;
            E *W, *PI, *Z, *WV, *H1, *H2, *PL, *small;
            const E *LK, *QL, *sel_c, *sig_c;
            E *w_c, *z_c, *pi_c, *lk_c;   // lk_c: h1, h2, lookup product (UltraPlonk)
            const E *lkt_c;               // range, key, table dom sep, q dom sep: per proof, or the resident copies
            E ev[2 * NW + 15];
        };
        std::vector<Inst> I(count);
        Transcript tr(kind, "PlonkProof");
        if (extra) tr.append_message("extra info", extra, extra_len);
        for (size_t i = 0; i < count; i++) {
            jf_plonk_pk *pk = pks[i];
            Inst &s = I[i];
            s.pk = pk;
            s.W = (E *)pk->d_w;
            s.PI = s.W + (size_t)NW * np;
            s.Z = (E *)pk->d_z;
            s.WV = (E *)pk->d_wv;
            s.small = (E *)pk->d_small;
            s.H1 = (E *)pk->d_hp;                                            // UltraPlonk
            s.H2 = s.H1 + np;
            s.PL = s.H1 + 2 * np;
            s.LK = (const E *)pk->d_lk;                                      // range | key | table dom sep | q dom sep
            s.QL = (const E *)pk->d_sel + (size_t)(NSEL - 1) * n;            // q_lookup polynomial (UltraPlonk)
            JF_CUDA(ctx, cudaMemcpyAsync(pk->d_wit, witnesses[i], fe * pk->num_vars, cudaMemcpyHostToDevice, st));
            // coset-evaluation slots: selectors, sigmas (or the resident copies), wires, z, PI [, range, key, tds, qds, h1, h2, pl]
            E *Ev = (E *)pk->d_e;
            if (pk->cache_coset) {
                s.sel_c = (const E *)pk->d_cached;
                s.sig_c = s.sel_c + (size_t)NSEL * m;
                s.w_c = Ev;
            } else {
                s.sel_c = Ev;
                s.sig_c = Ev + (size_t)NSEL * m;
                s.w_c = Ev + (size_t)(NSEL + NW) * m;
                // independent of this proof's challenges: runs on the side stream beside rounds 1 and 2
                JF_TRY(on_side(ctx, pk0, [&]() -> int {
                    JF_TRY(coset_fft_rows(ctx, pk, (const E *)pk->d_sel, n, n, NSEL, Ev, pk->zero_sel));
                    return coset_fft_rows(ctx, pk, (const E *)pk->d_sig, n, n, NW, Ev + (size_t)NSEL * m);
                }));
            }
            s.z_c = s.w_c + (size_t)NW * m;
            s.pi_c = s.z_c + m;
            if (ULTRA && pk->cache_coset) {
                s.lkt_c = s.sig_c + (size_t)NW * m;
                s.lk_c = s.pi_c + m;  // 3 more rows
            } else {
                E *t = s.pi_c + m;    // UltraPlonk: 7 more rows
                s.lkt_c = t;
                s.lk_c = t + 4 * m;
                if (ULTRA) JF_TRY(on_side(ctx, pk0, [&]() -> int { return coset_fft_rows(ctx, pk, s.LK, n, n, 4, t); }));
            }
            std::vector<E> pub(pk->num_inputs);
            for (size_t k = 0; k < pk->num_inputs; k++) pub[k] = H::fr_from_limbs(witnesses[i] + 4 * (size_t)pk->pub_vars[k]);
            append_vk_and_pub_input(tr, pk, pub.data(), pub.size());
        }
        // ---- round 1 (prover.rs:72-87): wire polynomials, masking, NW commitments, PI polynomial ----
        for (size_t i = 0; i < count; i++) {
            Inst &s = I[i];
            jf_plonk_pk *pk = s.pk;
            ProofT *out = &outs[i];
            JF_LAUNCH(ctx, "gather_wires", gather_wires_kernel<Fr><<<(unsigned)((NW * n + 255) / 256), 256, 0, st>>>(
                (const E *)pk->d_wit, pk->d_wire_vars, (size_t)NW * n, s.WV));
            JF_CUDA(ctx, cudaMemsetAsync(s.W, 0, fe * (NW + 1) * np, st));
            JF_CUDA(ctx, cudaMemcpy2DAsync(s.W, fe * np, s.WV, fe * n, fe * n, NW, cudaMemcpyDeviceToDevice, st));
            if (pk->num_inputs)
                JF_LAUNCH(ctx, "pub_input", pub_input_kernel<Fr><<<(pk->num_inputs + 127) / 128, 128, 0, st>>>(
                    (const E *)pk->d_wit, pk->d_wire_vars + (size_t)4 * n, pk->d_gate_ids, pk->num_inputs, s.PI));
            JF_TRY(intt_n(ctx, pk, s.W, NW + 1, np));
            for (int j = 0; j < NW; j++)
                JF_LAUNCH(ctx, "blind", blind_kernel<Fr><<<1, 32, 0, st>>>(s.W + (size_t)j * np, n, bl_w + i * 2 * NW + 2 * j, 2));
            // the wire / PI polynomials are final: their coset NTTs run beside the commitments
            JF_TRY(on_side(ctx, pk0, [&]() -> int {
                JF_TRY(coset_fft_rows(ctx, pk, s.W, np, n + 2, NW, s.w_c));
                if (pk->num_inputs == 0 && pk->skip_zero) return JF_OK;  // PI(X) = 0: nothing to transform
                return coset_fft_rows(ctx, pk, s.PI, np, n, 1, s.pi_c);
            }));
            if (pk->lag && !pk->comm) {
                // Lagrange basis: commit(wire j) = sum_i value_i [L_i(beta)] G + b_0 (P_n - P_0) + b_1 (P_(n+1) - P_1): an MSM over the
                // witness values (zeros are skipped, small values have few digits) and the two masking scalars; same point
                E *sc = (E *)pk->d_lagsc;
                JF_CUDA(ctx, cudaMemcpy2DAsync(sc, fe * (n + 2), s.WV, fe * n, fe * n, NW, cudaMemcpyDeviceToDevice, st));
                JF_CUDA(ctx, cudaMemcpy2DAsync(sc + n, fe * (n + 2), bl_w + i * 2 * NW, fe * 2, fe * 2, NW, cudaMemcpyDeviceToDevice, st));
                MsmJob mj[NW];
                for (int j = 0; j < NW; j++) mj[j] = MsmJob{0, sc + (size_t)j * (n + 2), n + 2, 1, (char *)pk->d_res + PT * j};
                JF_TRY(msm_run_many(ctx, pk->lag, mj, NW));
            } else {
                CommitJob jobs[NW];
                for (int j = 0; j < NW; j++) jobs[j] = {s.W + (size_t)j * np, n + 2, j};
                JF_TRY(commit_many(ctx, pk, jobs, NW));
            }
            JF_TRY(fetch_commits(ctx, pk, 0, NW, out->wires_poly_comms, out->wires_inf));
            for (int j = 0; j < NW; j++) tr_g1(tr, "witness_poly_comms", out->wires_poly_comms + 2 * L * j, out->wires_inf[j]);
        }
        const E tau = challenge(tr, "tau");  // squeezed even without Plookup (snark.rs:293)
        // ---- round 1.5 (Plookup; prover.rs:98-123, constraint_system.rs:1290-1309,1370-1418) ----
        if constexpr (ULTRA) {
            for (size_t i = 0; i < count; i++) {
                Inst &s = I[i];
                jf_plonk_pk *pk = s.pk;
                ProofT *out = &outs[i];
                JF_LAUNCH(ctx, "merged_values", merged_values_kernel<Fr><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
                    (const E *)pk->d_lk_evals, s.WV, tau, (uint32_t)n, (E *)pk->d_mt, (E *)pk->d_ml));
                JF_TRY(sorted_vector(ctx, pk));
                JF_CUDA(ctx, cudaMemsetAsync(s.H1, 0, fe * 3 * np, st));
                JF_CUDA(ctx, cudaMemcpyAsync(s.H1, pk->d_sorted, fe * n, cudaMemcpyDeviceToDevice, st));
                JF_CUDA(ctx, cudaMemcpyAsync(s.H2, (const E *)pk->d_sorted + (n - 1), fe * n, cudaMemcpyDeviceToDevice, st));
                JF_TRY(intt_n(ctx, pk, s.H1, 2, np));
                JF_LAUNCH(ctx, "blind", blind_kernel<Fr><<<1, 32, 0, st>>>(s.H1, n, bl_h + i * 6, 3));
                JF_LAUNCH(ctx, "blind", blind_kernel<Fr><<<1, 32, 0, st>>>(s.H2, n, bl_h + i * 6 + 3, 3));
                JF_TRY(on_side(ctx, pk0, [&]() -> int { return coset_fft_rows(ctx, pk, s.H1, np, n + 3, 2, s.lk_c); }));
                CommitJob jobs[2] = {{s.H1, n + 3, 0}, {s.H2, n + 3, 1}};
                JF_TRY(commit_many(ctx, pk, jobs, 2));
                JF_TRY(fetch_commits(ctx, pk, 0, 2, out->h_poly_comms, out->h_inf));
                for (int j = 0; j < 2; j++) tr_g1(tr, "h_poly_comms", out->h_poly_comms + 2 * L * j, out->h_inf[j]);
            }
        }
        // ---- round 2 (prover.rs:125-141; constraint_system.rs:1197-1223) ----
        const E beta = challenge(tr, "beta"), gamma = challenge(tr, "gamma");
        for (size_t i = 0; i < count; i++) {
            Inst &s = I[i];
            jf_plonk_pk *pk = s.pk;
            ProofT *out = &outs[i];
            PermArgs<Fr, NW> pa;
            pa.wv = s.WV;
            pa.sigma = (const E *)pk->d_sig_evals;
            for (int j = 0; j < NW; j++) pa.beta_k[j] = E::mul(beta, kf(pk, j));
            pa.beta = beta;
            pa.gamma = gamma;
            pa.omega = lf(pk->omega_n);
            pa.a = (E *)pk->d_a;
            pa.b = (E *)pk->d_b;
            pa.n = (uint32_t)n;
            const unsigned threads = (unsigned)((n + PERM_I - 1) / PERM_I);
            JF_LAUNCH(ctx, "perm_ab", perm_ab_kernel<Fr, NW><<<(threads + 127) / 128, 128, 0, st>>>(pa));
            E *PA = (E *)pk->d_t, *SB = (E *)pk->d_s;
            JF_TRY((fscan<Fr, OpMul, false>(ctx, pa.a, PA, n, (E *)pk->d_tmp)));
            JF_TRY((fscan<Fr, OpMul, true>(ctx, pa.b, SB, n, (E *)pk->d_tmp)));
            JF_LAUNCH(ctx, "inv_one", inv_one_kernel<Fr><<<1, 32, 0, st>>>(SB, s.small));
            JF_CUDA(ctx, cudaMemsetAsync(s.Z, 0, fe * np, st));
            JF_LAUNCH(ctx, "z_combine", z_combine_kernel<Fr><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(PA, SB, s.small, s.Z, (uint32_t)n));
            JF_TRY(intt_n(ctx, pk, s.Z, 1, np));
            JF_LAUNCH(ctx, "blind", blind_kernel<Fr><<<1, 32, 0, st>>>(s.Z, n, bl_z + i * 3, 3));
            JF_TRY(on_side(ctx, pk0, [&]() -> int { return coset_fft_rows(ctx, pk, s.Z, np, n + 3, 1, s.z_c); }));
            JF_TRY(commit_dev(ctx, pk, s.Z, n + 3, 0));
            JF_TRY(fetch_commits(ctx, pk, 0, 1, out->prod_perm_poly_comm, &out->prod_perm_inf));
            tr_g1(tr, "perm_poly_comms", out->prod_perm_poly_comm, out->prod_perm_inf);
        }
        const E bp1 = E::add(E::one(), beta), gb = E::mul(gamma, bp1);
        // ---- round 2.5 (Plookup product; prover.rs:150-190, constraint_system.rs:1311-1368) ----
        if constexpr (ULTRA) {
            for (size_t i = 0; i < count; i++) {
                Inst &s = I[i];
                jf_plonk_pk *pk = s.pk;
                ProofT *out = &outs[i];
                E *A = (E *)pk->d_a, *B = (E *)pk->d_b, *PA = (E *)pk->d_t, *SB = (E *)pk->d_s;
                JF_LAUNCH(ctx, "lookup_ab", lookup_ab_kernel<Fr><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
                    (const E *)pk->d_mt, (const E *)pk->d_ml, (const E *)pk->d_sorted, beta, gamma, bp1, gb, (uint32_t)n, A, B));
                JF_TRY((fscan<Fr, OpMul, false>(ctx, A, PA, n, (E *)pk->d_tmp)));
                JF_TRY((fscan<Fr, OpMul, true>(ctx, B, SB, n, (E *)pk->d_tmp)));
                JF_LAUNCH(ctx, "inv_one", inv_one_kernel<Fr><<<1, 32, 0, st>>>(SB, s.small));
                JF_LAUNCH(ctx, "z_combine", z_combine_kernel<Fr><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(PA, SB, s.small, s.PL, (uint32_t)n));
                JF_LAUNCH(ctx, "set_one", set_one_kernel<Fr><<<1, 32, 0, st>>>(s.PL + (n - 1)));  // the reference pushes 1 as the last value
                JF_TRY(intt_n(ctx, pk, s.PL, 1, np));
                JF_LAUNCH(ctx, "blind", blind_kernel<Fr><<<1, 32, 0, st>>>(s.PL, n, bl_pl + i * 3, 3));
                JF_TRY(on_side(ctx, pk0, [&]() -> int { return coset_fft_rows(ctx, pk, s.PL, np, n + 3, 1, s.lk_c + 2 * m); }));
                JF_TRY(commit_dev(ctx, pk, s.PL, n + 3, 0));
                JF_TRY(fetch_commits(ctx, pk, 0, 1, out->prod_lookup_poly_comm, &out->prod_lookup_inf));
                tr_g1(tr, "plookup_poly_comms", out->prod_lookup_poly_comm, out->prod_lookup_inf);
            }
        }
        // ---- round 3 (prover.rs:192-209, 512-673, 773-888, 902-960): ONE quotient for the whole batch ----
        const E alpha = challenge(tr, "alpha");
        const E alpha2 = E::sqr(alpha), alpha3 = E::mul(alpha2, alpha);
        const E alpha_step = ULTRA ? E::mul(E::sqr(alpha3), alpha) : alpha3;  // alpha^7 | alpha^3 (prover.rs:665-669)
        const size_t deg = NW * (n + 1) + 2;  // quotient_polynomial_degree (prover.rs:1126-1128)
        const size_t last_len = deg + 1 - (size_t)(NW - 1) * (n + 2);
        const E *split = (const E *)pk0->d_split;
        {
            E alpha_base = E::one();
            for (size_t i = 0; i < count; i++) {
                Inst &s = I[i];
                jf_plonk_pk *pk = s.pk;
                JF_TRY(join_side(ctx, pk0));  // this instance's coset evaluation vectors are in place
                QuotArgs<Fr, NW> q;
                q.sel = s.sel_c;
                q.sig = s.sig_c;
                q.w = s.w_c;
                q.z = s.z_c;
                q.pi = (pk->num_inputs == 0 && pk->skip_zero) ? nullptr : s.pi_c;
                q.inv_nx1 = (const E *)pk->d_inv_nx1;
                q.x_lo = (const E *)pk->d_xlo;
                q.x_hi = (const E *)pk->d_xhi;
                q.lo_bits = pk->lo_bits;
                for (int j = 0; j < NW; j++) q.beta_k[j] = E::mul(beta, kf(pk, j));
                q.beta = beta;
                q.gamma = gamma;
                q.alpha = alpha;
                q.alpha2 = alpha2;
                for (int r = 0; r < 8; r++) q.zh_inv[r] = lf(pk->zh_inv[r]);
                q.omega_inv = lf(pk->omega_n_inv);
                if (ULTRA) {
                    q.lk.range = s.lkt_c;
                    q.lk.key = s.lkt_c + m;
                    q.lk.tds = s.lkt_c + 2 * m;
                    q.lk.qds = s.lkt_c + 3 * m;
                    q.lk.h1 = s.lk_c;
                    q.lk.h2 = s.lk_c + m;
                    q.lk.pl = s.lk_c + 2 * m;
                    q.lk.tau = tau;
                    q.lk.bp1 = bp1;
                    q.lk.gb = gb;
                    q.lk.alpha3 = alpha3;
                }
                q.scale = alpha_base;
                q.acc_mode = i ? 1u : 0u;
                q.out = (E *)pk0->d_q;
                q.m = (uint32_t)m;
                q.ratio = 8;
                q.zero_sel = pk->zero_sel;
                q.sub = (uint32_t)pk->sub;
                q.log_n = pk->log_n;
                for (int r = 0; r < 8; r++) q.row_map[r] = pk->shard_rows ? (uint32_t)pk->row_map[r < SUB_MAX ? r : 0] : (uint32_t)r;
                if (m) JF_LAUNCH(ctx, "quotient", quotient_kernel<Fr, NW><<<(unsigned)((m + 127) / 128), 128, 0, st>>>(q));
                alpha_base = E::mul(alpha_base, alpha_step);
            }
            jf_plonk_pk *pk = pk0;
            const E *T = (const E *)pk->d_q;  // quotient coefficients
            const size_t m_full = pk->mq;
            if (pk->sub) {
                if (rows_eff(pk)) JF_TRY(ntt_run_cosets(ctx, C::FR_ID, pk->d_q, n, n, pk->d_q, pk->log_n, 1, offs_eff(pk), rows_eff(pk), 1));
                SolveArgs<Fr, SUB> sa;
                sa.T = (const E *)pk->d_q;
                if (pk->shard_rows) {
                    // every rank needs every row's interpolant: row r travels from rank r mod nranks (NVLink, one grouped broadcast)
                    JF_TRY(comm_bcast_rows(ctx, pk->comm, pk->d_q, pk->d_b, fe * n, pk->sub));
                    sa.T = (const E *)pk->d_b;  // free since round 2.5
                }
                sa.t = (E *)pk->d_a;  // free since round 2
                sa.n = (uint32_t)n;
                memcpy(sa.vinv, pk->vinv, sizeof sa.vinv);  // row-major SUB x SUB on both sides
                JF_LAUNCH(ctx, "subcoset_solve", subcoset_solve_kernel<Fr, SUB><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(sa));
                T = (const E *)pk->d_a;
            } else {
                JF_TRY(ntt_run(ctx, C::FR_ID, pk->d_q, pk->d_q, m, pk->log_m, 1, pk->gen_limbs, 1, m));
            }
            JF_LAUNCH(ctx, "degree_check", degree_check_kernel<Fr><<<(unsigned)((m_full - deg + 255) / 256), 256, 0, st>>>(T, deg, m_full, ctx->d_err + 1));
            dim3 grid((unsigned)((np + 255) / 256), NW);
            JF_LAUNCH(ctx, "split", split_kernel<Fr, NW><<<grid, 256, 0, st>>>(T, n, deg + 1, bl_split, (E *)pk->d_split, np));
            CommitJob jobs[NW];
            for (int i = 0; i < NW; i++) jobs[i] = {(E *)pk->d_split + (size_t)i * np, i < NW - 1 ? n + 3 : last_len, i};
            JF_TRY(commit_many(ctx, pk, jobs, NW));
            ProofT *out = &outs[0];
            JF_TRY(fetch_commits(ctx, pk, 0, NW, out->split_quot_poly_comms, out->split_inf));
            for (int i = 0; i < NW; i++) tr_g1(tr, "quot_poly_comms", out->split_quot_poly_comms + 2 * L * i, out->split_inf[i]);
        }
        // ---- round 4 (prover.rs:216-235) and 4.5 (Plookup evaluations, :239-297) ----
        const E zeta = challenge(tr, "zeta");
        const E omega = lf(pk0->omega_n);
        const E zeta_w = E::mul(zeta, omega);
        constexpr int NEV = 2 * NW;  // NW wires, NW - 1 sigmas, z(zeta w)
        // Plookup evaluations in `PlookupEvaluations` declaration order (structs.rs:496-541)
        enum { P_RANGE, P_KEY, P_TDS, P_QDS, P_H1, P_QL, P_PLN, P_RANGEN, P_KEYN, P_TDSN, P_H1N, P_H2N, P_QLN, P_W3N, P_W4N };
        for (size_t i = 0; i < count; i++) {
            Inst &s = I[i];
            jf_plonk_pk *pk = s.pk;
            E *small = s.small;
            for (int j = 0; j < NW; j++) JF_TRY(eval_dev(ctx, pk, s.W + (size_t)j * np, n + 2, zeta, small + j));
            for (int j = 0; j < NW - 1; j++) JF_TRY(eval_dev(ctx, pk, (const E *)pk->d_sig + (size_t)j * n, n, zeta, small + NW + j));
            JF_TRY(eval_dev(ctx, pk, s.Z, n + 3, zeta_w, small + NEV - 1));
            if (ULTRA) {
                E *pe = small + NEV;
                const E *LK = s.LK;
                JF_TRY(eval_dev(ctx, pk, LK, n, zeta, pe + P_RANGE));
                JF_TRY(eval_dev(ctx, pk, LK + n, n, zeta, pe + P_KEY));
                JF_TRY(eval_dev(ctx, pk, LK + 2 * n, n, zeta, pe + P_TDS));
                JF_TRY(eval_dev(ctx, pk, LK + 3 * n, n, zeta, pe + P_QDS));
                JF_TRY(eval_dev(ctx, pk, s.H1, n + 3, zeta, pe + P_H1));
                JF_TRY(eval_dev(ctx, pk, s.QL, n, zeta, pe + P_QL));
                JF_TRY(eval_dev(ctx, pk, s.PL, n + 3, zeta_w, pe + P_PLN));
                JF_TRY(eval_dev(ctx, pk, LK, n, zeta_w, pe + P_RANGEN));
                JF_TRY(eval_dev(ctx, pk, LK + n, n, zeta_w, pe + P_KEYN));
                JF_TRY(eval_dev(ctx, pk, LK + 2 * n, n, zeta_w, pe + P_TDSN));
                JF_TRY(eval_dev(ctx, pk, s.H1, n + 3, zeta_w, pe + P_H1N));
                JF_TRY(eval_dev(ctx, pk, s.H2, n + 3, zeta_w, pe + P_H2N));
                JF_TRY(eval_dev(ctx, pk, s.QL, n, zeta_w, pe + P_QLN));
                JF_TRY(eval_dev(ctx, pk, s.W + 3 * np, n + 2, zeta_w, pe + P_W3N));
                JF_TRY(eval_dev(ctx, pk, s.W + 4 * np, n + 2, zeta_w, pe + P_W4N));
            }
            JF_CUDA(ctx, cudaMemcpyAsync(s.ev, small, fe * (NEV + (ULTRA ? 15 : 0)), cudaMemcpyDeviceToHost, st));
        }
        JF_CUDA(ctx, cudaStreamSynchronize(st));
        for (size_t i = 0; i < count; i++) {  // append_proof_evaluations for every instance ...
            const E *ev = I[i].ev;
            for (int j = 0; j < NW; j++) tr_fr(tr, "wire_evals", ev[j]);
            for (int j = 0; j < NW - 1; j++) tr_fr(tr, "wire_sigma_evals", ev[NW + j]);
            tr_fr(tr, "perm_next_eval", ev[NEV - 1]);
        }
        if (ULTRA) {  // ... then append_plookup_evaluations for every instance (transcript/mod.rs:168-201): six of the fifteen
            for (size_t i = 0; i < count; i++) {
                const E *pe = I[i].ev + NEV;
                tr_fr(tr, "lookup_table_eval", pe[P_RANGE]);
                tr_fr(tr, "h_1_eval", pe[P_H1]);
                tr_fr(tr, "prod_next_eval", pe[P_PLN]);
                tr_fr(tr, "lookup_table_next_eval", pe[P_RANGEN]);
                tr_fr(tr, "h_1_next_eval", pe[P_H1N]);
                tr_fr(tr, "h_2_next_eval", pe[P_H2N]);
            }
        }
        // ---- round 5 (snark.rs:403-449; prover.rs:302-360, 362-460, 490-509, 963-1113) ----
        const E v = challenge(tr, "v");
        {
            const E one = E::one();
            const E vanish = E::sub(pow_small(zeta, n), one);
            const E zeta_n2 = E::mul(E::mul(E::add(vanish, one), zeta), zeta);
            const E nf = E::from_u32((uint32_t)n);
            const E lagrange_1 = E::mul(vanish, E::inv(E::mul(nf, E::sub(zeta, one))));
            std::vector<Term> terms, shifted_terms;
            // quotient part: -Z_H(zeta) * sum_i zeta^(i (n+2)) t_i
            E coeff = E::neg(vanish);
            for (int i = 0; i < NW; i++) {
                terms.push_back({split + (size_t)i * np, i < NW - 1 ? n + 3 : last_len, coeff});
                coeff = E::mul(coeff, zeta_n2);
            }
            // the non-quotient parts enter with alpha_base_i; the opening batch continues the powers of v across the instances:
            // lin + v w_0 + .. + v^NW w_(NW-1) + v^(NW+1) sigma_0 + .. [+ range, key, h1, q_lookup, tds, qds] + (next instance) ..
            E alpha_base = one, vp = v, sp = one;
            for (size_t i = 0; i < count; i++) {
                Inst &s = I[i];
                jf_plonk_pk *pk = s.pk;
                const E *we = s.ev, *se = s.ev + NW, *pe = s.ev + NEV;
                const E pne = s.ev[NEV - 1];
                // circuit part: the 13 arithmetic selectors (q_lookup is not part of it)
                const E *sel = (const E *)pk->d_sel;
                const E w01 = E::mul(we[0], we[1]), w23 = E::mul(we[2], we[3]);
                const E qs[13] = {we[0], we[1], we[2], we[3], w01, w23, pow_small(we[0], 5), pow_small(we[1], 5), pow_small(we[2], 5),
                                  pow_small(we[3], 5), E::neg(we[4]), one, E::mul(E::mul(w01, w23), we[4])};
                for (int k = 0; k < 13; k++)
                    if (!((pk->zero_sel >> k) & 1u)) terms.push_back({sel + (size_t)k * n, n, E::mul(alpha_base, qs[k])});
                // permutation part
                E c1 = alpha;
                for (int j = 0; j < NW; j++) c1 = E::mul(c1, E::add(E::add(we[j], E::mul(E::mul(beta, kf(pk, j)), zeta)), gamma));
                c1 = E::add(c1, E::mul(alpha2, lagrange_1));
                E c2 = E::mul(E::mul(alpha, beta), pne);
                for (int j = 0; j < NW - 1; j++) c2 = E::mul(c2, E::add(E::add(we[j], E::mul(beta, se[j])), gamma));
                terms.push_back({s.Z, n + 3, E::mul(alpha_base, c1)});
                const E *sig = (const E *)pk->d_sig;
                terms.push_back({sig + (size_t)(NW - 1) * n, n, E::mul(alpha_base, E::neg(c2))});
                if (ULTRA) {  // compute_lin_poly_plookup_contribution (prover.rs:1037-1113)
                    const E alpha4 = E::sqr(alpha2), alpha5 = E::mul(alpha4, alpha), alpha6 = E::mul(alpha4, alpha2);
                    const E g_inv = lf(pk->omega_n_inv), zmg = E::sub(zeta, g_inv);
                    const E lagrange_n = E::mul(E::mul(vanish, g_inv), E::inv(E::mul(nf, zmg)));
                    auto merged = [&](const E &a, const E &qv, const E &b, const E &cc, const E &d, const E &e) {
                        E t = E::add(d, E::mul(tau, e));
                        t = E::add(cc, E::mul(tau, t));
                        t = E::add(b, E::mul(tau, t));
                        return E::add(a, E::mul(E::mul(qv, tau), t));
                    };
                    const E mt = merged(pe[P_RANGE], pe[P_QL], pe[P_TDS], pe[P_KEY], we[3], we[4]);
                    const E mtn = merged(pe[P_RANGEN], pe[P_QLN], pe[P_TDSN], pe[P_KEYN], pe[P_W3N], pe[P_W4N]);
                    const E ml = merged(we[5], pe[P_QL], pe[P_QDS], we[0], we[1], we[2]);
                    E cpl = E::mul(E::mul(E::mul(E::mul(alpha6, zmg), bp1), E::add(gamma, ml)), E::add(E::add(gb, mt), E::mul(beta, mtn)));
                    cpl = E::add(cpl, E::add(E::mul(alpha4, lagrange_1), E::mul(alpha5, lagrange_n)));
                    terms.push_back({s.PL, n + 3, E::mul(alpha_base, cpl)});
                    const E ch2 = E::neg(E::mul(E::mul(E::mul(alpha6, zmg), pe[P_PLN]), E::add(E::add(gb, pe[P_H1]), E::mul(beta, pe[P_H1N]))));
                    terms.push_back({s.H2, n + 3, E::mul(alpha_base, ch2)});
                }
                for (int j = 0; j < NW; j++) {
                    terms.push_back({s.W + (size_t)j * np, n + 2, vp});
                    vp = E::mul(vp, v);
                }
                for (int j = 0; j < NW - 1; j++) {
                    terms.push_back({sig + (size_t)j * n, n, vp});
                    vp = E::mul(vp, v);
                }
                shifted_terms.push_back({s.Z, n + 3, sp});
                sp = E::mul(sp, v);
                if (ULTRA) {  // plookup_open_polys_ref / plookup_shifted_open_polys_ref (prover.rs:427-460)
                    const E *LK = s.LK;
                    const E *ps[6] = {LK, LK + n, s.H1, s.QL, LK + 2 * n, LK + 3 * n};
                    const size_t ls[6] = {n, n, n + 3, n, n, n};
                    for (int j = 0; j < 6; j++) {
                        terms.push_back({ps[j], ls[j], vp});
                        vp = E::mul(vp, v);
                    }
                    const E *qs2[9] = {s.PL, LK, LK + n, s.H1, s.H2, s.QL, s.W + 3 * np, s.W + 4 * np, LK + 2 * n};
                    const size_t ls2[9] = {n + 3, n, n, n + 3, n + 3, n, n + 2, n + 2, n};
                    for (int j = 0; j < 9; j++) {
                        shifted_terms.push_back({qs2[j], ls2[j], sp});
                        sp = E::mul(sp, v);
                    }
                }
                alpha_base = E::mul(alpha_base, alpha_step);
            }
            jf_plonk_pk *pk = pk0;
            JF_TRY(lincomb_many(ctx, terms, (E *)pk->d_bp, n + 3));
            // the shifted opening is independent of the batch polynomial: side stream, own scratch
            E *side_buf = (E *)pk->d_side;  // t, s, quotient (np each), scan temporaries
            JF_TRY(on_side(ctx, pk0, [&]() -> int {
                const E *shifted = I[0].Z;
                if (shifted_terms.size() > 1) {
                    // the PI slot of W is free after round 3 (its coset evaluations were taken in round 1)
                    JF_TRY(lincomb_many(ctx, shifted_terms, I[0].PI, n + 3));
                    shifted = I[0].PI;
                }
                JF_TRY(div_linear_dev(ctx, side_buf, side_buf + np, side_buf + 3 * np, shifted, n + 3, zeta_w, side_buf + 2 * np));
                return commit_dev(ctx, pk, side_buf + 2 * np, n + 2, 1);
            }));
            E *WZ = (E *)pk->d_wz;
            JF_TRY(div_linear_dev(ctx, pk, (const E *)pk->d_bp, n + 3, zeta, WZ));
            JF_TRY(commit_dev(ctx, pk, WZ, n + 2, 0));
            JF_TRY(join_side(ctx, pk0));
            uint64_t xy[2 * 2 * 6];
            int inf[2];
            JF_TRY(fetch_commits(ctx, pk, 0, 2, xy, inf));
            ProofT *out = &outs[0];
            memcpy(out->opening_proof, xy, sizeof(uint64_t) * 2 * L);
            memcpy(out->shifted_opening_proof, xy + 2 * L, sizeof(uint64_t) * 2 * L);
            out->opening_inf = inf[0];
            out->shifted_opening_inf = inf[1];
        }
        for (size_t i = 0; i < count; i++) {
            ProofT *out = &outs[i];
            const E *ev = I[i].ev;
            if (i) {  // the shared parts of the batch proof, replicated
                memcpy(out->split_quot_poly_comms, outs[0].split_quot_poly_comms, sizeof out->split_quot_poly_comms);
                memcpy(out->split_inf, outs[0].split_inf, sizeof out->split_inf);
                memcpy(out->opening_proof, outs[0].opening_proof, sizeof out->opening_proof);
                memcpy(out->shifted_opening_proof, outs[0].shifted_opening_proof, sizeof out->shifted_opening_proof);
                out->opening_inf = outs[0].opening_inf;
                out->shifted_opening_inf = outs[0].shifted_opening_inf;
            }
            for (int j = 0; j < NW; j++) H::fr_to_limbs(ev[j], out->wires_evals + 4 * j);
            for (int j = 0; j < NW - 1; j++) H::fr_to_limbs(ev[NW + j], out->wire_sigma_evals + 4 * j);
            H::fr_to_limbs(ev[NEV - 1], out->perm_next_eval);
            if constexpr (ULTRA) {
                for (int j = 0; j < 15; j++) H::fr_to_limbs(ev[NEV + j], out->plookup_evals + 4 * j);
                H::fr_to_limbs(tau, out->challenges + 0);
                H::fr_to_limbs(beta, out->challenges + 4);
                H::fr_to_limbs(gamma, out->challenges + 8);
                H::fr_to_limbs(alpha, out->challenges + 12);
                H::fr_to_limbs(zeta, out->challenges + 16);
                H::fr_to_limbs(v, out->challenges + 20);
            } else {
                H::fr_to_limbs(beta, out->challenges + 0);
                H::fr_to_limbs(gamma, out->challenges + 4);
                H::fr_to_limbs(alpha, out->challenges + 8);
                H::fr_to_limbs(zeta, out->challenges + 12);
                H::fr_to_limbs(v, out->challenges + 16);
            }
            out->curve = pk0->curve;
            I[i].pk->proofs_done++;
        }
        return JF_OK;
    }

    // `UnivariateKzgPCS::open` (primitives/src/pcs/univariate_kzg/mod.rs:135-161) for `batch` polynomials:
    // witness polynomial p / (X - z), its commitment, and p(z); everything on the device.
    static int kzg_open(jf_ctx *ctx, const jf_srs *srs, const uint64_t *const *polys, const size_t *lens, size_t batch,
                        const uint64_t *points, uint64_t *out_xy, int *out_inf, uint64_t *out_evals) {
        size_t max_len = 1;
        for (size_t i = 0; i < batch; i++) max_len = std::max(max_len, lens[i]);
        void *dp, *dt, *ds, *dw, *dtmp, *dres;
        const size_t fe = sizeof(E);
        JF_TRY(scratch(ctx, "open_p", fe * max_len, &dp));
        JF_TRY(scratch(ctx, "open_t", fe * max_len, &dt));
        JF_TRY(scratch(ctx, "open_s", fe * max_len, &ds));
        JF_TRY(scratch(ctx, "open_w", fe * max_len, &dw));
        JF_TRY(scratch(ctx, "open_tmp", fe * (max_len / 256 + 4096), &dtmp));
        JF_TRY(scratch(ctx, "open_res", (PT + fe) * (batch ? batch : 1), &dres));
        E *d_ev = (E *)((char *)dres + PT * batch);
        // The witness-polynomial commitments form MSM groups (one bucket reduction per group); the upload, evaluation
        // and division of polynomial i are enqueued right before its MSM, so one set of buffers serves the whole batch.
        struct Open {
            jf_ctx *ctx;
            const uint64_t *const *polys;
            const uint64_t *points;
            const size_t *len;  // trimmed lengths
            E *dp, *dt, *ds, *dw, *dtmp, *d_ev;
            size_t first;
        };
        std::vector<size_t> tl(batch);
        for (size_t i = 0; i < batch; i++) {
            // DensePolynomial semantics: trailing (high-degree) zero coefficients do not count
            size_t len = lens[i];
            while (len && !(polys[i][4 * (len - 1)] | polys[i][4 * (len - 1) + 1] | polys[i][4 * (len - 1) + 2] | polys[i][4 * (len - 1) + 3])) len--;
            if (len > srs->n + 1) return fail(ctx, JF_ERR_INVALID_ARG, "open: polynomial degree exceeds the commit key");
            tl[i] = len;
        }
        auto prepare = [](void *user, int k) -> int {
            Open *o = (Open *)user;
            jf_ctx *ctx = o->ctx;
            const size_t i = o->first + k, len = o->len[i];
            const size_t fe = sizeof(E);
            const E z = H::fr_from_limbs(o->points + 4 * i);
            if (len) JF_CUDA(ctx, cudaMemcpyAsync(o->dp, o->polys[i], fe * len, cudaMemcpyHostToDevice, ctx->stream));
            if (len == 0) JF_CUDA(ctx, cudaMemsetAsync(o->d_ev + i, 0, fe, ctx->stream));
            else JF_TRY(eval_dev(ctx, o->dtmp, (const E *)o->dp, len, z, o->d_ev + i));
            if (len >= 2) JF_TRY(div_linear_dev(ctx, o->dt, o->ds, o->dtmp, (const E *)o->dp, len, z, o->dw));
            return JF_OK;
        };
        constexpr size_t GROUP = 16;
        for (size_t g0 = 0; g0 < batch; g0 += GROUP) {
            const int cnt = (int)std::min(GROUP, batch - g0);
            MsmJob jobs[GROUP];
            for (int k = 0; k < cnt; k++)  // constant / zero polynomial: empty MSM, identity proof
                jobs[k] = MsmJob{0, dw, tl[g0 + k] >= 2 ? tl[g0 + k] - 1 : 0, 1, (char *)dres + PT * (g0 + k)};
            Open o{ctx, polys, points, tl.data(), (E *)dp, (E *)dt, (E *)ds, (E *)dw, (E *)dtmp, d_ev, g0};
            JF_TRY(msm_run_many(ctx, srs, jobs, cnt, prepare, &o));
        }
        void *h;
        JF_TRY(pinned(ctx, (PT + fe) * batch + 64, &h));
        JF_CUDA(ctx, cudaMemcpyAsync(h, dres, (PT + fe) * batch, cudaMemcpyDeviceToHost, ctx->stream));
        int herr = 0;
        JF_CUDA(ctx, cudaMemcpyAsync(&herr, ctx->d_err, sizeof herr, cudaMemcpyDeviceToHost, ctx->stream));
        JF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (herr) {
            cudaMemsetAsync(ctx->d_err, 0, sizeof herr, ctx->stream);
            return fail(ctx, herr, "open: a coefficient is not a reduced field element");
        }
        for (size_t i = 0; i < batch; i++)
            JF_TRY(msm_finish_host(ctx, srs->curve, (const uint64_t *)((const char *)h + PT * i), 1, out_xy + (size_t)2 * L * i, out_inf + i));
        memcpy(out_evals, (const char *)h + PT * batch, fe * batch);
        return JF_OK;
    }

    // ------------------------------------------------------------------------------------------
    // `PlonkKzgSnark::link_proofs` (plonk/src/proof_system/proof_linking.rs:79-216): quotient (a1 - a2) / Z_D over the linking
    // domain D = {g^(offset + i), i < size} (g the 2^alignment-th root of unity), its commitment, the challenge eta from a fresh
    // transcript, and the KZG opening at eta of a1 - a2 - q Z_D(eta).  d_a1 / d_a2: the two first-wire polynomials in HBM.
    //
    // The reference builds Z_D by `size` polynomial multiplications and divides by schoolbook long division.  Here:
    //  * a1 - a2 is evaluated on the 2^alignment-th roots of unity (one fold mod X^N - 1, one size-N transform).  When it
    //    vanishes on D (every honestly linked pair of proofs) the division is exact, and the quotient is the pointwise ratio on the
    //    coset GENERATOR <w_N> -- three size-N transforms and one batch inversion, whatever `size` is; Z_D's coefficients follow from
    //    the q-binomial theorem in O(size) host multiplications (its roots are consecutive powers of g).
    //  * otherwise (the pair is NOT linked: the proof will be rejected, but the reference still produces bytes) ark-poly's `/`
    //    drops a non-zero remainder; dividing by the `size` linear factors one after the other (scan division each) yields that
    //    same floor quotient.  flags & 1 forces this form; the tests compare the two.
    static E root_of_unity(unsigned log) {
        uint32_t e[8];
        Limbs<Fr>::p(e);
        e[0] -= 1;
        for (int s = 0; s < Fr::TWO_ADICITY; s++)
            for (int i = 0; i < 8; i++) e[i] = (e[i] >> 1) | (i < 7 ? e[i + 1] << 31 : 0);
        E w = E::pow(E::from_u32(Fr::GENERATOR), e, 8);
        for (unsigned s = 0; s < Fr::TWO_ADICITY - log; s++) w = E::sqr(w);
        return w;
    }
    // coefficients (low degree first, size + 1 of them) of Z_D = prod_{i < size} (X - g^(offset + i)):
    // prod_{i<s} (y - g^i) = sum_k (-1)^k g^(k(k-1)/2) [s; k]_g y^(s-k)  with  [s; k]_g = [s; k-1]_g (1 - g^(s-k+1)) / (1 - g^k),
    // and Z_D(X) = g^(offset s) P(X / g^offset).  1 - g^k != 0 because size < 2^alignment.
    static std::vector<E> vanishing_coeffs(const E &g, size_t offset, size_t size) {
        const size_t s = size;
        std::vector<E> gp(s + 1), den(s + 1), pre(s + 1);
        gp[0] = E::one();
        for (size_t k = 1; k <= s; k++) gp[k] = E::mul(gp[k - 1], g);
        E run = E::one();
        for (size_t k = 1; k <= s; k++) {  // batch inversion of 1 - g^k
            den[k] = E::sub(E::one(), gp[k]);
            pre[k] = run;
            run = E::mul(run, den[k]);
        }
        E inv = E::inv(run);
        std::vector<E> dinv(s + 1);
        for (size_t k = s; k >= 1; k--) {
            dinv[k] = E::mul(inv, pre[k]);
            inv = E::mul(inv, den[k]);
        }
        std::vector<E> z(s + 1);
        E binom = E::one(), e = E::one();            // [s; k]_g and g^(k(k-1)/2 + offset k)
        E step = pow_small(g, (uint64_t)offset);     // g^(offset + k)
        z[s] = E::one();
        for (size_t k = 1; k <= s; k++) {
            binom = E::mul(E::mul(binom, den[s - k + 1]), dinv[k]);
            e = E::mul(e, step);
            step = E::mul(step, g);
            const E c = E::mul(binom, e);
            z[s - k] = (k & 1) ? E::neg(c) : c;
        }
        return z;
    }

    // scratch of the linking quotient: diff = the dividend (max_len coefficients), A / B ping-pong or transform buffers, T / S scan
    // buffers, each max(max_len, N) + 8 elements
    struct LinkBufs {
        E *diff, *A, *B, *T, *S, *tmp, *small;
        int *flag;
    };
    static unsigned link_log_N(size_t max_len, unsigned alignment, size_t size) {
        const size_t qlen = max_len > size ? max_len - size : 0;
        unsigned log_N = alignment;
        while (((size_t)1 << log_N) < std::max(qlen, size + 1)) log_N++;
        return log_N;
    }
    static int link_bufs(jf_ctx *ctx, size_t max_len, unsigned alignment, size_t size, int flags, LinkBufs *lb) {
        const size_t fe = sizeof(E);
        const unsigned log_N = link_log_N(max_len, alignment, size);
        const bool may_be_fast = !(flags & 1) && max_len > size && log_N <= (unsigned)Fr::TWO_ADICITY && log_N <= 27;
        const size_t cap = std::max(max_len, may_be_fast ? (size_t)1 << log_N : (size_t)0) + 8;
        void *p_diff, *p_a, *p_b, *p_t, *p_s, *p_tmp, *p_small;
        JF_TRY(scratch(ctx, "link_diff", fe * cap, &p_diff));
        JF_TRY(scratch(ctx, "link_a", fe * cap, &p_a));
        JF_TRY(scratch(ctx, "link_b", fe * cap, &p_b));
        JF_TRY(scratch(ctx, "link_t", fe * cap, &p_t));
        JF_TRY(scratch(ctx, "link_s", fe * cap, &p_s));
        JF_TRY(scratch(ctx, "link_tmp", fe * (cap / 256 + 4096), &p_tmp));
        JF_TRY(scratch(ctx, "link_small", fe + 64, &p_small));
        *lb = LinkBufs{(E *)p_diff, (E *)p_a, (E *)p_b, (E *)p_t, (E *)p_s, (E *)p_tmp, (E *)p_small, (int *)((char *)p_small + fe)};
        return JF_OK;
    }
    // *Q (one of lb.A / lb.B, *q_len coefficients) = floor(lb.diff / Z_D); *path: 0 exact division on a coset, 1 linear divisions
    static int link_quotient(jf_ctx *ctx, const LinkBufs &lb, size_t max_len, unsigned alignment, size_t offset, size_t size, int flags,
                             E **Q_out, size_t *q_len_out, int *path) {
        const size_t fe = sizeof(E);
        cudaStream_t st = ctx->stream;
        E *diff = lb.diff, *A = lb.A, *B = lb.B, *T = lb.T, *S = lb.S, *tmp = lb.tmp, *small = lb.small;
        int *d_flag = lb.flag;
        const E g = root_of_unity(alignment);
        const E r0 = pow_small(g, (uint64_t)offset);
        const size_t qlen = max_len > size ? max_len - size : 0;
        const unsigned log_N = link_log_N(max_len, alignment, size);
        const size_t N = (size_t)1 << log_N;
        bool fast = !(flags & 1) && qlen > 0 && log_N <= (unsigned)Fr::TWO_ADICITY && log_N <= 27, vanishes = false;
        const E gen = E::from_u32(Fr::GENERATOR);
        uint64_t gen_limbs[4];
        H::fr_to_limbs(gen, gen_limbs);
        if (fast) {  // does a1 - a2 vanish on the linking domain?
            JF_CUDA(ctx, cudaMemsetAsync(d_flag, 0, sizeof(int), st));
            JF_LAUNCH(ctx, "fold", fold_kernel<Fr><<<(unsigned)((N + 255) / 256), 256, 0, st>>>(diff, max_len, N, E::one(), A));
            JF_TRY(ntt_run(ctx, C::FR_ID, A, A, N, log_N, 0, nullptr, 1, N));
            JF_LAUNCH(ctx, "link_roots_check", link_roots_check_kernel<Fr><<<(unsigned)((size + 255) / 256), 256, 0, st>>>(
                A, (uint32_t)(N >> alignment), (uint32_t)(((size_t)1 << alignment) - 1), (uint32_t)offset, (uint32_t)size, d_flag));
            int flag = 0;
            JF_CUDA(ctx, cudaMemcpyAsync(&flag, d_flag, sizeof flag, cudaMemcpyDeviceToHost, st));
            JF_CUDA(ctx, cudaStreamSynchronize(st));
            // A non-zero remainder (the pair is not linked; or the dividend is one party's SHARE of a1 - a2): the floor quotient is
            // (p - p mod Z_D) / Z_D, and because Z_D divides X^(2^a) - 1, p mod Z_D = (p mod (X^(2^a) - 1)) mod Z_D: one fold to 2^a
            // coefficients and a schoolbook reduction of those in shared memory.  With the remainder taken off the exact division
            // applies.  (2^a coefficients must fit one CTA's shared memory: alignments up to 12.)
            vanishes = flag == 0;
            fast = vanishes || alignment <= 12;
        }
        E *Q = A;  // quotient coefficients
        size_t q_len = qlen;
        if (fast) {
            *path = vanishes ? 0 : 2;
            const std::vector<E> z = vanishing_coeffs(g, offset, size);
            void *hz;
            JF_TRY(pinned(ctx, fe * (size + 1), &hz));
            memcpy(hz, z.data(), fe * (size + 1));
            JF_CUDA(ctx, cudaMemcpyAsync(B, hz, fe * (size + 1), cudaMemcpyHostToDevice, st));
            JF_LAUNCH(ctx, "fold", fold_kernel<Fr><<<(unsigned)((N + 255) / 256), 256, 0, st>>>(diff, max_len, N, pow_small(gen, (uint64_t)N), A));
            if (!vanishes) {
                const uint32_t m = 1u << alignment;
                const size_t smem = fe * m;
                JF_CUDA(ctx, cudaFuncSetAttribute(poly_mod_monic_kernel<Fr>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                JF_LAUNCH(ctx, "fold", fold_kernel<Fr><<<(m + 255) / 256, 256, 0, st>>>(diff, max_len, (size_t)m, E::one(), T));
                JF_LAUNCH(ctx, "poly_mod_monic", poly_mod_monic_kernel<Fr><<<1, 1024, smem, st>>>(T, m, B, (uint32_t)size, S));
                JF_LAUNCH(ctx, "vsub", vsub_kernel<Fr><<<(unsigned)((size + 255) / 256), 256, 0, st>>>(A, S, size));
            }
            JF_TRY(ntt_run(ctx, C::FR_ID, A, A, N, log_N, 0, gen_limbs, 1, N));
            JF_TRY(ntt_run(ctx, C::FR_ID, B, B, size + 1, log_N, 0, gen_limbs, 1, N));
            JF_TRY((fscan<Fr, OpMul, false>(ctx, B, T, N, tmp)));
            JF_TRY((fscan<Fr, OpMul, true>(ctx, B, S, N, tmp)));
            JF_LAUNCH(ctx, "inv_one", inv_one_kernel<Fr><<<1, 32, 0, st>>>(S, small));
            JF_LAUNCH(ctx, "inv_combine", inv_combine_kernel<Fr><<<(unsigned)((N + 255) / 256), 256, 0, st>>>(T, S, small, B, N));
            JF_LAUNCH(ctx, "vmul", vmul_kernel<Fr><<<(unsigned)((N + 255) / 256), 256, 0, st>>>(A, B, A, N));
            JF_TRY(ntt_run(ctx, C::FR_ID, A, A, N, log_N, 1, gen_limbs, 1, N));
        } else {
            *path = 1;
            // floor((a1 - a2) / Z_D) as `size` divisions by a linear factor
            const E *cur = diff;
            size_t len = max_len;
            E r = r0, rinv = E::inv(r0);
            const E ginv = E::inv(g);
            E *bufs[2] = {A, B};
            int which = 0;
            for (size_t i = 0; i < size && len; i++) {
                if (len < 2) {
                    len = 0;
                    break;
                }
                E *nxt = bufs[which];
                const unsigned b1 = (unsigned)((len + 256 * MP_I - 1) / (256 * MP_I));
                JF_LAUNCH(ctx, "mulpow", mulpow_kernel<Fr><<<b1, 256, 0, st>>>(cur, T, len, 0, 0, r));
                JF_TRY((fscan<Fr, OpAdd, true>(ctx, T, S, len, tmp)));
                JF_LAUNCH(ctx, "mulpow", mulpow_kernel<Fr><<<b1, 256, 0, st>>>(S, nxt, len - 1, 1, 1, rinv));
                cur = nxt;
                which ^= 1;
                len--;
                r = E::mul(r, g);
                rinv = E::mul(rinv, ginv);
            }
            q_len = len;
            Q = const_cast<E *>(cur);
            if (q_len == 0) Q = A;
        }
        *Q_out = Q;
        *q_len_out = q_len;
        return JF_OK;
    }

    // floor(p / Z_D) for `batch` polynomials in host memory (division by the PUBLIC vanishing polynomial of a link group is linear,
    // so the collaborative prover applies it to every component of its shares: multiprover/proof_system/proof_linking.rs:127-138).
    // outs[i] receives max(lens[i] - size, 0) coefficients.
    static int div_link_domain(jf_ctx *ctx, const uint64_t *const *polys, const size_t *lens, size_t batch, unsigned alignment,
                               size_t offset, size_t size, int flags, uint64_t *const *outs) {
        if (alignment > (unsigned)Fr::TWO_ADICITY || alignment > 30)
            return fail(ctx, JF_ERR_DOMAIN_TOO_LARGE, "div_link_domain: the group alignment exceeds the field's two-adicity");
        if (size == 0 || offset + size >= ((size_t)1 << alignment))
            return fail(ctx, JF_ERR_INVALID_ARG, "div_link_domain: the link group is empty or exceeds its alignment");
        const size_t fe = sizeof(E);
        size_t max_len = 0;
        for (size_t i = 0; i < batch; i++) max_len = std::max(max_len, lens[i]);
        if (max_len >> 27) return fail(ctx, JF_ERR_INVALID_ARG, "div_link_domain: polynomial too long");
        LinkBufs lb;
        JF_TRY(link_bufs(ctx, max_len, alignment, size, flags, &lb));
        for (size_t i = 0; i < batch; i++) {
            const size_t want = lens[i] > size ? lens[i] - size : 0;
            if (!want) continue;
            JF_CUDA(ctx, cudaMemcpyAsync(lb.diff, polys[i], fe * lens[i], cudaMemcpyHostToDevice, ctx->stream));
            E *Q = lb.A;
            size_t q_len = 0;
            int path = 0;
            JF_TRY(link_quotient(ctx, lb, lens[i], alignment, offset, size, flags, &Q, &q_len, &path));
            if (q_len < want) memset(outs[i] + 4 * q_len, 0, fe * (want - q_len));
            if (q_len) JF_CUDA(ctx, cudaMemcpyAsync(outs[i], Q, fe * std::min(q_len, want), cudaMemcpyDeviceToHost, ctx->stream));
            JF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // lb is reused by the next polynomial
        }
        return JF_OK;
    }

    struct LinkOut {
        uint64_t q_xy[12], open_xy[12], eta[4];
        int q_inf, open_inf, path;
    };
    static int link_proofs(jf_ctx *ctx, const jf_srs *srs, const E *d_a1, size_t len1, const uint64_t *a1_comm, int a1_inf,
                           const E *d_a2, size_t len2, const uint64_t *a2_comm, int a2_inf, unsigned alignment, size_t offset,
                           size_t size, int kind, int flags, LinkOut *out) {
        if (alignment > (unsigned)Fr::TWO_ADICITY || alignment > 30)
            return fail(ctx, JF_ERR_DOMAIN_TOO_LARGE, "link_proofs: the group alignment exceeds the field's two-adicity");
        if (size == 0 || offset + size >= ((size_t)1 << alignment))  // validate_layout (linkable_circuit.rs:352-370)
            return fail(ctx, JF_ERR_INVALID_ARG, "link_proofs: the link group is empty or exceeds its alignment");
        const size_t max_len = std::max(len1, len2);
        if (max_len > srs->n + 1) return fail(ctx, JF_ERR_INVALID_ARG, "link_proofs: polynomial degree exceeds the commit key");
        const size_t fe = sizeof(E);
        cudaStream_t st = ctx->stream;
        const E g = root_of_unity(alignment);
        const E r0 = pow_small(g, (uint64_t)offset);
        memset(out, 0, sizeof *out);
        LinkBufs lb;
        JF_TRY(link_bufs(ctx, max_len, alignment, size, flags, &lb));
        void *p_res, *h;
        JF_TRY(scratch(ctx, "link_res", 2 * PT, &p_res));
        // one pinned staging area for the whole call (Z_D's coefficients go up through it, the two results come down): reserved
        // here at its largest so that no later request re-allocates it under a copy in flight
        JF_TRY(pinned(ctx, std::max(fe * (size + 1), 2 * PT + 64), &h));
        E *diff = lb.diff, *A = lb.A, *B = lb.B, *T = lb.T, *S = lb.S, *tmp = lb.tmp;
        // a1 - a2
        if (max_len) {
            std::vector<Term> terms = {{d_a1, len1, E::one()}, {d_a2, len2, E::neg(E::one())}};
            JF_TRY(lincomb_many(ctx, terms, diff, max_len));
        }
        E *Q = A;
        size_t q_len = 0;
        JF_TRY(link_quotient(ctx, lb, max_len, alignment, offset, size, flags, &Q, &q_len, &out->path));
        // quotient commitment (`UnivariateKzgPCS::commit`, mod.rs:90-116)
        JF_TRY(msm_run(ctx, srs, 0, Q, q_len, 1, p_res));
        JF_CUDA(ctx, cudaMemcpyAsync(h, p_res, PT, cudaMemcpyDeviceToHost, st));
        int herr = 0;
        JF_CUDA(ctx, cudaMemcpyAsync(&herr, ctx->d_err, sizeof herr, cudaMemcpyDeviceToHost, st));
        JF_CUDA(ctx, cudaStreamSynchronize(st));
        if (herr) {
            cudaMemsetAsync(ctx->d_err, 0, sizeof herr, st);
            return fail(ctx, herr, "link_proofs: a coefficient is not a reduced field element");
        }
        JF_TRY(msm_finish_host(ctx, srs->curve, (const uint64_t *)h, 1, out->q_xy, &out->q_inf));
        // compute_quotient_challenge (:171-191)
        Transcript tr(kind, "PlonkLinkingProof");
        tr_g1(tr, "linking_wire_comms", a1_comm, a1_inf);
        tr_g1(tr, "linking_wire_comms", a2_comm, a2_inf);
        tr_g1(tr, "quotient_comm", out->q_xy, out->q_inf);
        const E eta = challenge(tr, "eta");
        H::fr_to_limbs(eta, out->eta);
        E zd = E::one(), r = r0;  // compute_vanishing_poly_eval (:162-177)
        for (size_t i = 0; i < size; i++) {
            zd = E::mul(zd, E::sub(eta, r));
            r = E::mul(r, g);
        }
        // compute_identity_opening (:193-216): (a1 - a2 - q Z_D(eta)) / (X - eta), committed
        E *ident = (Q == A) ? B : A;
        size_t wlen = 0;
        if (max_len >= 2) {
            std::vector<Term> terms = {{diff, max_len, E::one()}, {Q, q_len, E::neg(zd)}};
            JF_TRY(lincomb_many(ctx, terms, ident, max_len));
            E *wit = diff;  // a1 - a2 is not needed any more
            JF_TRY(div_linear_dev(ctx, T, S, tmp, ident, max_len, eta, wit));
            wlen = max_len - 1;
            JF_TRY(msm_run(ctx, srs, 0, wit, wlen, 1, (char *)p_res + PT));
        } else {
            JF_TRY(msm_run(ctx, srs, 0, diff, 0, 1, (char *)p_res + PT));
        }
        JF_CUDA(ctx, cudaMemcpyAsync(h, (char *)p_res + PT, PT, cudaMemcpyDeviceToHost, st));
        JF_CUDA(ctx, cudaMemcpyAsync(&herr, ctx->d_err, sizeof herr, cudaMemcpyDeviceToHost, st));
        JF_CUDA(ctx, cudaStreamSynchronize(st));
        if (herr) {
            cudaMemsetAsync(ctx->d_err, 0, sizeof herr, st);
            return fail(ctx, herr, "link_proofs: a coefficient is not a reduced field element");
        }
        JF_TRY(msm_finish_host(ctx, srs->curve, (const uint64_t *)h, 1, out->open_xy, &out->open_inf));
        return JF_OK;
    }

    // `Proof<E>` CanonicalSerialize, compressed (structs.rs:62-84; PlookupProof :255-265, PlookupEvaluations :496-541)
    template <class ProofT> static size_t serialize(const ProofT *p, uint8_t *out) {
        uint8_t *o = out;
        auto u64 = [&](uint64_t v) { memcpy(o, &v, 8); o += 8; };
        auto g1 = [&](const uint64_t *xy, int inf) { H::g1_bytes(xy, inf, o); o += 8 * L; };
        auto fr = [&](const uint64_t *l) { H::fr_bytes(H::fr_from_limbs(l), o); o += 32; };
        u64(NW);
        for (int j = 0; j < NW; j++) g1(p->wires_poly_comms + 2 * L * j, p->wires_inf[j]);
        g1(p->prod_perm_poly_comm, p->prod_perm_inf);
        u64(NW);
        for (int j = 0; j < NW; j++) g1(p->split_quot_poly_comms + 2 * L * j, p->split_inf[j]);
        g1(p->opening_proof, p->opening_inf);
        g1(p->shifted_opening_proof, p->shifted_opening_inf);
        u64(NW);
        for (int j = 0; j < NW; j++) fr(p->wires_evals + 4 * j);
        u64(NW - 1);
        for (int j = 0; j < NW - 1; j++) fr(p->wire_sigma_evals + 4 * j);
        fr(p->perm_next_eval);
        if constexpr (ULTRA) {
            *o++ = 1;  // plookup_proof: Some
            u64(2);
            for (int j = 0; j < 2; j++) g1(p->h_poly_comms + 2 * L * j, p->h_inf[j]);
            g1(p->prod_lookup_poly_comm, p->prod_lookup_inf);
            for (int j = 0; j < 15; j++) fr(p->plookup_evals + 4 * j);
        } else {
            *o++ = 0;  // plookup_proof: None
        }
        return (size_t)(o - out);
    }

    // `BatchProof<E>` CanonicalSerialize, compressed (structs.rs:271-292): per-instance vectors first, then the shared parts
    template <class ProofT> static size_t serialize_batch(const ProofT *ps, size_t count, uint8_t *out) {
        uint8_t *o = out;
        auto u64 = [&](uint64_t v) { memcpy(o, &v, 8); o += 8; };
        auto g1 = [&](const uint64_t *xy, int inf) { H::g1_bytes(xy, inf, o); o += 8 * L; };
        auto fr = [&](const uint64_t *l) { H::fr_bytes(H::fr_from_limbs(l), o); o += 32; };
        u64(count);  // wires_poly_comms_vec
        for (size_t i = 0; i < count; i++) {
            u64(NW);
            for (int j = 0; j < NW; j++) g1(ps[i].wires_poly_comms + 2 * L * j, ps[i].wires_inf[j]);
        }
        u64(count);  // prod_perm_poly_comms_vec
        for (size_t i = 0; i < count; i++) g1(ps[i].prod_perm_poly_comm, ps[i].prod_perm_inf);
        u64(count);  // poly_evals_vec
        for (size_t i = 0; i < count; i++) {
            u64(NW);
            for (int j = 0; j < NW; j++) fr(ps[i].wires_evals + 4 * j);
            u64(NW - 1);
            for (int j = 0; j < NW - 1; j++) fr(ps[i].wire_sigma_evals + 4 * j);
            fr(ps[i].perm_next_eval);
        }
        u64(count);  // plookup_proofs_vec
        for (size_t i = 0; i < count; i++) {
            if constexpr (ULTRA) {
                *o++ = 1;
                u64(2);
                for (int j = 0; j < 2; j++) g1(ps[i].h_poly_comms + 2 * L * j, ps[i].h_inf[j]);
                g1(ps[i].prod_lookup_poly_comm, ps[i].prod_lookup_inf);
                for (int j = 0; j < 15; j++) fr(ps[i].plookup_evals + 4 * j);
            } else {
                *o++ = 0;
            }
        }
        u64(NW);  // split_quot_poly_comms
        for (int j = 0; j < NW; j++) g1(ps[0].split_quot_poly_comms + 2 * L * j, ps[0].split_inf[j]);
        g1(ps[0].opening_proof, ps[0].opening_inf);
        g1(ps[0].shifted_opening_proof, ps[0].shifted_opening_inf);
        return (size_t)(o - out);
    }
    static size_t batch_size(size_t count) {
        const size_t pt = 8 * L;
        size_t per = 8 + NW * pt + pt + 8 + NW * 32 + 8 + (NW - 1) * 32 + 32 + 1;
        if (ULTRA) per += 8 + 2 * pt + pt + 15 * 32;
        return 4 * 8 + count * per + 8 + NW * pt + 2 * pt;
    }
};

struct Bn254Plonk : Bn254G1 {
    static constexpr int FR_ID = JF_BN254_FR;
};
struct Bls12381Plonk : Bls12381G1 {
    static constexpr int FR_ID = JF_BLS12_381_FR;
};

}  // namespace jf

#define JF_GUARD(ctx)                             \
    if (!(ctx)) return JF_ERR_INVALID_ARG;        \
    std::lock_guard<std::mutex> lock_((ctx)->mu); \
    cudaSetDevice((ctx)->device)

// entry of the two proof-linking forms (host hints / resident polynomials)
template <class P>
static int link_entry(jf_ctx *ctx, const jf_srs *srs, const void *d_a1, size_t len1, const uint64_t *a1_comm, int a1_inf, const void *d_a2,
                      size_t len2, const uint64_t *a2_comm, int a2_inf, unsigned alignment, size_t offset, size_t size, int kind, int flags,
                      jf_link_proof *out) {
    using PL = Plonk<P>;
    typename PL::LinkOut lo;
    int rc = PL::link_proofs(ctx, srs, (const typename PL::E *)d_a1, len1, a1_comm, a1_inf, (const typename PL::E *)d_a2, len2, a2_comm,
                             a2_inf, alignment, offset, size, kind, flags, &lo);
    if (rc != JF_OK) {
        cudaStreamSynchronize(ctx->stream);  // nothing of this call may still be in flight when the caller reuses its buffers
        return rc;
    }
    memset(out, 0, sizeof *out);
    out->curve = srs->curve;
    memcpy(out->quotient_commitment, lo.q_xy, sizeof lo.q_xy);
    out->quotient_inf = lo.q_inf;
    memcpy(out->opening_proof, lo.open_xy, sizeof lo.open_xy);
    out->opening_inf = lo.open_inf;
    memcpy(out->eta, lo.eta, sizeof lo.eta);
    out->path = lo.path;
    return JF_OK;
}

// shared argument checks and the error drain of the two batch entry points
template <int NWT, class ProofT>
static int batch_prove_entry(jf_ctx *ctx, jf_plonk_pk *const *pks, size_t count, const uint64_t *const *witnesses, const uint64_t *blinders,
                             int transcript_kind, const uint8_t *extra_msg, size_t extra_len, ProofT *out) {
    if (!pks || !witnesses || !blinders || !out || count == 0)
        return fail(ctx, JF_ERR_INVALID_ARG, "batch_prove: zero number of circuits/proving keys or a null argument");
    if (count > 64) return fail(ctx, JF_ERR_INVALID_ARG, "batch_prove: at most 64 instances");
    for (size_t i = 0; i < count; i++)
        if (!pks[i] || !witnesses[i]) return fail(ctx, JF_ERR_INVALID_ARG, "batch_prove: null proving key or witness");
    if (transcript_kind != 0 && transcript_kind != 1) return fail(ctx, JF_ERR_INVALID_ARG, "batch_prove: unknown transcript");
    const int rc = pks[0]->curve == JF_BN254
                       ? Plonk<Bn254Plonk, NWT>::batch_prove(ctx, pks, count, witnesses, blinders, transcript_kind, extra_msg, extra_len, out)
                       : Plonk<Bls12381Plonk, NWT>::batch_prove(ctx, pks, count, witnesses, blinders, transcript_kind, extra_msg, extra_len, out);
    if (rc != JF_OK) {
        cudaStreamSynchronize(ctx->stream);
        for (size_t i = 0; i < count; i++)
            if (pks[i]->side) cudaStreamSynchronize(pks[i]->side);
    }
    return rc;
}

extern "C" {

int jf_plonk_preprocess(jf_ctx *ctx, const jf_srs *srs, unsigned log_n, const uint64_t *selector_evals, const uint64_t *sigma_evals,
                        const uint64_t *k, const uint32_t *wire_variables, size_t num_vars, const uint32_t *pub_input_gate_ids,
                        size_t num_inputs, int flags, jf_plonk_pk **out) {
    JF_GUARD(ctx);
    if (!srs || !selector_evals || !sigma_evals || !k || !wire_variables || !out || (num_inputs && !pub_input_gate_ids))
        return fail(ctx, JF_ERR_INVALID_ARG, "plonk_preprocess: null argument");
    *out = nullptr;
    if (srs->curve == JF_BN254)
        return Plonk<Bn254Plonk>::preprocess(ctx, srs, log_n, selector_evals, sigma_evals, k, wire_variables, num_vars,
                                             pub_input_gate_ids, num_inputs, flags, out);
    return Plonk<Bls12381Plonk>::preprocess(ctx, srs, log_n, selector_evals, sigma_evals, k, wire_variables, num_vars,
                                            pub_input_gate_ids, num_inputs, flags, out);
}

int jf_ultraplonk_preprocess(jf_ctx *ctx, const jf_srs *srs, unsigned log_n, const uint64_t *selector_evals, const uint64_t *sigma_evals,
                             const uint64_t *k, const uint32_t *wire_variables, size_t num_vars, const uint32_t *pub_input_gate_ids,
                             size_t num_inputs, unsigned range_bit_len, const uint64_t *table_key_evals,
                             const uint64_t *table_dom_sep_evals, const uint64_t *q_dom_sep_evals, int flags, jf_plonk_pk **out) {
    JF_GUARD(ctx);
    if (!srs || !selector_evals || !sigma_evals || !k || !wire_variables || !out || (num_inputs && !pub_input_gate_ids) ||
        !table_key_evals || !table_dom_sep_evals || !q_dom_sep_evals)
        return fail(ctx, JF_ERR_INVALID_ARG, "ultraplonk_preprocess: null argument");
    *out = nullptr;
    if (srs->curve == JF_BN254) {
        Plonk<Bn254Plonk, NW_ULTRA>::LookupCols lc{range_bit_len, table_key_evals, table_dom_sep_evals, q_dom_sep_evals};
        return Plonk<Bn254Plonk, NW_ULTRA>::preprocess(ctx, srs, log_n, selector_evals, sigma_evals, k, wire_variables, num_vars,
                                                       pub_input_gate_ids, num_inputs, flags & 6, out, &lc);
    }
    Plonk<Bls12381Plonk, NW_ULTRA>::LookupCols lc{range_bit_len, table_key_evals, table_dom_sep_evals, q_dom_sep_evals};
    return Plonk<Bls12381Plonk, NW_ULTRA>::preprocess(ctx, srs, log_n, selector_evals, sigma_evals, k, wire_variables, num_vars,
                                                      pub_input_gate_ids, num_inputs, flags & 6, out, &lc);
}

int jf_ultraplonk_prove(jf_ctx *ctx, jf_plonk_pk *pk, const uint64_t *witness, const uint64_t *blinders, int transcript_kind,
                        const uint8_t *extra_msg, size_t extra_len, jf_ultraplonk_proof *out) {
    JF_GUARD(ctx);
    if (!pk || !witness || !blinders || !out) return fail(ctx, JF_ERR_INVALID_ARG, "ultraplonk_prove: null argument");
    if (pk->nw != NW_ULTRA) return fail(ctx, JF_ERR_INVALID_ARG, "ultraplonk_prove: the proving key is a TurboPlonk key (use jf_plonk_prove)");
    if (transcript_kind != 0 && transcript_kind != 1) return fail(ctx, JF_ERR_INVALID_ARG, "ultraplonk_prove: unknown transcript");
    const int rc = pk->curve == JF_BN254
                       ? Plonk<Bn254Plonk, NW_ULTRA>::prove(ctx, pk, witness, blinders, transcript_kind, extra_msg, extra_len, out)
                       : Plonk<Bls12381Plonk, NW_ULTRA>::prove(ctx, pk, witness, blinders, transcript_kind, extra_msg, extra_len, out);
    if (rc != JF_OK) {
        cudaStreamSynchronize(ctx->stream);
        if (pk->side) cudaStreamSynchronize(pk->side);
    }
    return rc;
}

int jf_plonk_batch_prove(jf_ctx *ctx, jf_plonk_pk *const *pks, size_t count, const uint64_t *const *witnesses, const uint64_t *blinders,
                         int transcript_kind, const uint8_t *extra_msg, size_t extra_len, jf_plonk_proof *out) {
    JF_GUARD(ctx);
    return batch_prove_entry<NW_TURBO>(ctx, pks, count, witnesses, blinders, transcript_kind, extra_msg, extra_len, out);
}

int jf_ultraplonk_batch_prove(jf_ctx *ctx, jf_plonk_pk *const *pks, size_t count, const uint64_t *const *witnesses,
                              const uint64_t *blinders, int transcript_kind, const uint8_t *extra_msg, size_t extra_len,
                              jf_ultraplonk_proof *out) {
    JF_GUARD(ctx);
    return batch_prove_entry<NW_ULTRA>(ctx, pks, count, witnesses, blinders, transcript_kind, extra_msg, extra_len, out);
}

long jf_plonk_batch_proof_serialize(const jf_plonk_proof *proofs, size_t count, uint8_t *out, size_t cap) {
    if (!proofs || !out || count == 0) return JF_ERR_INVALID_ARG;
    const bool bn = proofs[0].curve == JF_BN254;
    if (!bn && proofs[0].curve != JF_BLS12_381) return JF_ERR_INVALID_ARG;
    if (cap < (bn ? Plonk<Bn254Plonk>::batch_size(count) : Plonk<Bls12381Plonk>::batch_size(count))) return JF_ERR_INVALID_ARG;
    return (long)(bn ? Plonk<Bn254Plonk>::serialize_batch(proofs, count, out) : Plonk<Bls12381Plonk>::serialize_batch(proofs, count, out));
}

long jf_ultraplonk_batch_proof_serialize(const jf_ultraplonk_proof *proofs, size_t count, uint8_t *out, size_t cap) {
    if (!proofs || !out || count == 0) return JF_ERR_INVALID_ARG;
    const bool bn = proofs[0].curve == JF_BN254;
    if (!bn && proofs[0].curve != JF_BLS12_381) return JF_ERR_INVALID_ARG;
    if (cap < (bn ? Plonk<Bn254Plonk, NW_ULTRA>::batch_size(count) : Plonk<Bls12381Plonk, NW_ULTRA>::batch_size(count))) return JF_ERR_INVALID_ARG;
    return (long)(bn ? Plonk<Bn254Plonk, NW_ULTRA>::serialize_batch(proofs, count, out)
                     : Plonk<Bls12381Plonk, NW_ULTRA>::serialize_batch(proofs, count, out));
}

long jf_ultraplonk_proof_serialize(const jf_ultraplonk_proof *proof, uint8_t *out, size_t cap) {
    if (!proof || !out) return JF_ERR_INVALID_ARG;
    const bool bn = proof->curve == JF_BN254;
    if (!bn && proof->curve != JF_BLS12_381) return JF_ERR_INVALID_ARG;
    const size_t L = bn ? 4 : 6;
    const size_t need = 8 + 6 * 8 * L + 8 * L + 8 + 6 * 8 * L + 2 * 8 * L + 8 + 6 * 32 + 8 + 5 * 32 + 32 + 1 + 8 + 2 * 8 * L + 8 * L + 15 * 32;
    if (cap < need) return JF_ERR_INVALID_ARG;
    return (long)(bn ? Plonk<Bn254Plonk, NW_ULTRA>::serialize(proof, out) : Plonk<Bls12381Plonk, NW_ULTRA>::serialize(proof, out));
}

int jf_plonk_vk_commitments(jf_ctx *ctx, const jf_plonk_pk *pk, uint64_t *out_xy, int *out_inf) {
    JF_GUARD(ctx);
    if (!pk || !out_xy || !out_inf) return fail(ctx, JF_ERR_INVALID_ARG, "plonk_vk_commitments: null argument");
    memcpy(out_xy, pk->vk_xy.data(), pk->vk_xy.size() * sizeof(uint64_t));
    memcpy(out_inf, pk->vk_inf.data(), pk->vk_inf.size() * sizeof(int));
    return JF_OK;
}

void jf_plonk_pk_free(jf_ctx *ctx, jf_plonk_pk *pk) {
    if (!pk) return;
    if (ctx) {
        std::lock_guard<std::mutex> lock(ctx->mu);
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        if (pk->side) cudaStreamSynchronize(pk->side);
    }
    release_pk(pk);
}

int jf_plonk_prove(jf_ctx *ctx, jf_plonk_pk *pk, const uint64_t *witness, const uint64_t *blinders, int transcript_kind,
                   const uint8_t *extra_msg, size_t extra_len, jf_plonk_proof *out) {
    JF_GUARD(ctx);
    if (!pk || !witness || !blinders || !out) return fail(ctx, JF_ERR_INVALID_ARG, "plonk_prove: null argument");
    if (pk->nw != NW_TURBO) return fail(ctx, JF_ERR_INVALID_ARG, "plonk_prove: the proving key is an UltraPlonk key (use jf_ultraplonk_prove)");
    if (transcript_kind != 0 && transcript_kind != 1) return fail(ctx, JF_ERR_INVALID_ARG, "plonk_prove: unknown transcript");
    const int rc = pk->curve == JF_BN254
                       ? Plonk<Bn254Plonk>::prove(ctx, pk, witness, blinders, transcript_kind, extra_msg, extra_len, out)
                       : Plonk<Bls12381Plonk>::prove(ctx, pk, witness, blinders, transcript_kind, extra_msg, extra_len, out);
    if (rc != JF_OK) {
        // an early return may leave kernels and copies queued against the key's scratch and the caller's buffers:
        // drain both streams so the next call starts clean (the context keeps its stream and lane)
        cudaStreamSynchronize(ctx->stream);
        if (pk->side) cudaStreamSynchronize(pk->side);
    }
    return rc;
}

int jf_kzg_open(jf_ctx *ctx, const jf_srs *srs, const uint64_t *const *polys, const size_t *lens, size_t batch,
                const uint64_t *points, uint64_t *out_proof_xy, int *out_infinity, uint64_t *out_evals) {
    JF_GUARD(ctx);
    if (!srs || !out_proof_xy || !out_infinity || !out_evals || (batch && (!polys || !lens || !points)))
        return fail(ctx, JF_ERR_INVALID_ARG, "kzg_open: null argument");
    for (size_t i = 0; i < batch; i++)
        if (lens[i] && !polys[i]) return fail(ctx, JF_ERR_INVALID_ARG, "kzg_open: null polynomial");
    if (batch == 0) return JF_OK;
    if (srs->curve == JF_BN254) return Plonk<Bn254Plonk>::kzg_open(ctx, srs, polys, lens, batch, points, out_proof_xy, out_infinity, out_evals);
    return Plonk<Bls12381Plonk>::kzg_open(ctx, srs, polys, lens, batch, points, out_proof_xy, out_infinity, out_evals);
}

// ---- proof linking ----------------------------------------------------------------------------------
int jf_plonk_pk_shard_commits(jf_ctx *ctx, jf_plonk_pk *pk, jf_comm *comm, const jf_srs *key_slice, size_t slice_start, int shard_round3) {
    JF_GUARD(ctx);
    if (!pk) return fail(ctx, JF_ERR_INVALID_ARG, "shard_commits: null proving key");
    if (!comm) {  // back to one GPU
        pk->comm = nullptr;
        pk->srs_slice = nullptr;
        pk->shard_start = 0;
        pk->shard_rows = pk->rows_local = 0;
        return JF_OK;
    }
    if (!key_slice || comm_ctx(comm) != ctx) return fail(ctx, JF_ERR_INVALID_ARG, "shard_commits: null key slice or a comm of another context");
    if (key_slice->curve != pk->curve) return fail(ctx, JF_ERR_INVALID_ARG, "shard_commits: the key slice is over another curve");
    if (comm_size(comm) > 16) return fail(ctx, JF_ERR_INVALID_ARG, "shard_commits: at most 16 ranks");
    if (!pk->d_parts) {
        const size_t pt = (size_t)key_slice->limbs64 * 32;
        JF_TRY(dalloc(ctx, pk, pt * 32 * 16, &pk->d_parts));
    }
    pk->comm = comm;
    pk->srs_slice = key_slice;
    pk->shard_start = slice_start;
    pk->shard_rows = pk->rows_local = 0;
    // round 3 by sub-coset: only in the sub-coset form, and not with resident coset evaluations (they are laid out for all rows)
    if (shard_round3 && pk->sub && !pk->cache_coset && comm_size(comm) > 1) {
        pk->shard_rows = 1;
        for (int r = comm_rank(comm); r < pk->sub; r += comm_size(comm)) {
            pk->row_map[pk->rows_local] = r;
            memcpy(pk->sub_off_local + 4 * pk->rows_local, pk->sub_off + 4 * r, 32);
            pk->rows_local++;
        }
    }
    return JF_OK;
}

int jf_plonk_link_hint(jf_ctx *ctx, const jf_plonk_pk *pk, uint64_t *out_poly, size_t cap, size_t *out_len) {
    JF_GUARD(ctx);
    if (!pk || !out_poly || !out_len) return fail(ctx, JF_ERR_INVALID_ARG, "link_hint: null argument");
    if (!pk->proofs_done) return fail(ctx, JF_ERR_INVALID_ARG, "link_hint: no proof has been made with this key yet");
    const size_t len = pk->n + 2;  // the masked first wire polynomial (prover.rs:82)
    if (cap < len) return fail(ctx, JF_ERR_INVALID_ARG, "link_hint: the buffer holds fewer than n + 2 coefficients");
    JF_CUDA(ctx, cudaMemcpyAsync(out_poly, pk->d_w, 32 * len, cudaMemcpyDeviceToHost, ctx->stream));
    JF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out_len = len;
    return JF_OK;
}

int jf_plonk_link_proofs(jf_ctx *ctx, const jf_srs *srs, const uint64_t *a1, size_t len1, const uint64_t *a1_comm_xy, int a1_inf,
                         const uint64_t *a2, size_t len2, const uint64_t *a2_comm_xy, int a2_inf, unsigned alignment, size_t offset,
                         size_t size, int transcript_kind, int flags, jf_link_proof *out) {
    JF_GUARD(ctx);
    if (!srs || !out || !a1_comm_xy || !a2_comm_xy || (len1 && !a1) || (len2 && !a2)) return fail(ctx, JF_ERR_INVALID_ARG, "link_proofs: null argument");
    if (transcript_kind != 0 && transcript_kind != 1) return fail(ctx, JF_ERR_INVALID_ARG, "link_proofs: unknown transcript kind");
    // DensePolynomial semantics: trailing (high-degree) zero coefficients do not count
    auto trim = [](const uint64_t *p, size_t len) {
        while (len && !(p[4 * (len - 1)] | p[4 * (len - 1) + 1] | p[4 * (len - 1) + 2] | p[4 * (len - 1) + 3])) len--;
        return len;
    };
    len1 = trim(a1, len1);
    len2 = trim(a2, len2);
    void *d1, *d2;
    JF_TRY(scratch(ctx, "link_in1", 32 * (len1 + 1), &d1));
    JF_TRY(scratch(ctx, "link_in2", 32 * (len2 + 1), &d2));
    if (len1) JF_CUDA(ctx, cudaMemcpyAsync(d1, a1, 32 * len1, cudaMemcpyHostToDevice, ctx->stream));
    if (len2) JF_CUDA(ctx, cudaMemcpyAsync(d2, a2, 32 * len2, cudaMemcpyHostToDevice, ctx->stream));
    if (srs->curve == JF_BN254)
        return link_entry<Bn254Plonk>(ctx, srs, d1, len1, a1_comm_xy, a1_inf, d2, len2, a2_comm_xy, a2_inf, alignment, offset, size,
                                      transcript_kind, flags, out);
    return link_entry<Bls12381Plonk>(ctx, srs, d1, len1, a1_comm_xy, a1_inf, d2, len2, a2_comm_xy, a2_inf, alignment, offset, size,
                                     transcript_kind, flags, out);
}

int jf_plonk_link_proofs_resident(jf_ctx *ctx, const jf_plonk_pk *lhs, const jf_plonk_proof *lhs_proof, const jf_plonk_pk *rhs,
                                  const jf_plonk_proof *rhs_proof, unsigned alignment, size_t offset, size_t size, int transcript_kind,
                                  int flags, jf_link_proof *out) {
    JF_GUARD(ctx);
    if (!lhs || !rhs || !lhs_proof || !rhs_proof || !out) return fail(ctx, JF_ERR_INVALID_ARG, "link_proofs: null argument");
    if (transcript_kind != 0 && transcript_kind != 1) return fail(ctx, JF_ERR_INVALID_ARG, "link_proofs: unknown transcript kind");
    if (lhs->nw != jf::NW_TURBO || rhs->nw != jf::NW_TURBO) return fail(ctx, JF_ERR_INVALID_ARG, "link_proofs: only TurboPlonk supports link groups");
    if (!lhs->proofs_done || !rhs->proofs_done) return fail(ctx, JF_ERR_INVALID_ARG, "link_proofs: no proof has been made with one of the keys");
    if (lhs->srs != rhs->srs) return fail(ctx, JF_ERR_INVALID_ARG, "link_proofs: the two proofs must share one commit key");
    const jf_srs *srs = lhs->srs;
    if (srs->curve == JF_BN254)
        return link_entry<Bn254Plonk>(ctx, srs, lhs->d_w, lhs->n + 2, lhs_proof->wires_poly_comms, lhs_proof->wires_inf[0], rhs->d_w, rhs->n + 2,
                                      rhs_proof->wires_poly_comms, rhs_proof->wires_inf[0], alignment, offset, size, transcript_kind, flags, out);
    return link_entry<Bls12381Plonk>(ctx, srs, lhs->d_w, lhs->n + 2, lhs_proof->wires_poly_comms, lhs_proof->wires_inf[0], rhs->d_w, rhs->n + 2,
                                     rhs_proof->wires_poly_comms, rhs_proof->wires_inf[0], alignment, offset, size, transcript_kind, flags, out);
}

int jf_poly_div_link_domain(jf_ctx *ctx, int field, const uint64_t *const *polys, const size_t *lens, size_t batch, unsigned alignment,
                            size_t offset, size_t size, int flags, uint64_t *const *out_quotients) {
    JF_GUARD(ctx);
    if (batch && (!polys || !lens || !out_quotients)) return fail(ctx, JF_ERR_INVALID_ARG, "div_link_domain: null argument");
    for (size_t i = 0; i < batch; i++)
        if (lens[i] && (!polys[i] || (lens[i] > size && !out_quotients[i]))) return fail(ctx, JF_ERR_INVALID_ARG, "div_link_domain: null polynomial");
    int rc;
    if (field == JF_BN254_FR) rc = Plonk<Bn254Plonk>::div_link_domain(ctx, polys, lens, batch, alignment, offset, size, flags, out_quotients);
    else if (field == JF_BLS12_381_FR) rc = Plonk<Bls12381Plonk>::div_link_domain(ctx, polys, lens, batch, alignment, offset, size, flags, out_quotients);
    else return fail(ctx, JF_ERR_INVALID_ARG, "div_link_domain: field must be a scalar field");
    if (rc != JF_OK) cudaStreamSynchronize(ctx->stream);
    return rc;
}

long jf_link_proof_serialize(const jf_link_proof *proof, uint8_t *out, size_t cap) {
    if (!proof || !out) return JF_ERR_INVALID_ARG;
    const size_t pt = proof->curve == JF_BN254 ? 32 : 48;
    if (cap < 2 * pt) return JF_ERR_INVALID_ARG;
    if (proof->curve == JF_BN254) {
        HostCurve<Bn254Plonk>::g1_bytes(proof->quotient_commitment, proof->quotient_inf, out);
        HostCurve<Bn254Plonk>::g1_bytes(proof->opening_proof, proof->opening_inf, out + pt);
    } else {
        HostCurve<Bls12381Plonk>::g1_bytes(proof->quotient_commitment, proof->quotient_inf, out);
        HostCurve<Bls12381Plonk>::g1_bytes(proof->opening_proof, proof->opening_inf, out + pt);
    }
    return (long)(2 * pt);
}

long jf_plonk_proof_serialize(const jf_plonk_proof *proof, uint8_t *out, size_t cap) {
    if (!proof || !out) return JF_ERR_INVALID_ARG;
    const size_t need = proof->curve == JF_BN254 ? 769 : 8 * 4 + 48 * 13 + 32 * 10 + 1;
    if (cap < need) return JF_ERR_INVALID_ARG;
    if (proof->curve == JF_BN254) return (long)Plonk<Bn254Plonk>::serialize(proof, out);
    return (long)Plonk<Bls12381Plonk>::serialize(proof, out);
}

// ---- host-only transcript entry points (no GPU needed; exercised by the CPU test-suite) ----------
void jf_keccak256(const uint8_t *data, size_t len, uint8_t out[32]) { keccak256(data, len, out); }

void *jf_transcript_new(int kind, const char *label) {
    if (kind != 0 && kind != 1) return nullptr;
    return new Transcript(kind, label ? label : "");
}
void jf_transcript_free(void *t) { delete static_cast<Transcript *>(t); }
void jf_transcript_append(void *t, const char *label, const uint8_t *msg, size_t len) {
    static_cast<Transcript *>(t)->append_message(label, msg, len);
}
/* field: JF_BN254_FR or JF_BLS12_381_FR; out = the challenge, 4 Montgomery limbs */
int jf_transcript_challenge(void *t, int field, const char *label, uint64_t *out) {
    if (!t || !out) return JF_ERR_INVALID_ARG;
    Transcript &tr = *static_cast<Transcript *>(t);
    if (field == JF_BN254_FR) {
        auto c = Plonk<Bn254Plonk>::challenge(tr, label);
        HostCurve<Bn254Plonk>::fr_to_limbs(c, out);
        return JF_OK;
    }
    if (field == JF_BLS12_381_FR) {
        auto c = Plonk<Bls12381Plonk>::challenge(tr, label);
        HostCurve<Bls12381Plonk>::fr_to_limbs(c, out);
        return JF_OK;
    }
    return JF_ERR_INVALID_ARG;
}

}  // extern "C"
