// Lagrange-basis commit key: L[j] = [L_j(beta)] G for the size-n domain H = <w_n>, from the monomial key P_k = [beta^k] G.
//
// `UnivariateKzgPCS::commit` (primitives/src/pcs/univariate_kzg/mod.rs:90-116) is linear in the polynomial, and the prover's wire
// polynomials are interpolants of the witness VALUES over H (relation/src/constraint_system.rs:1225-1247), so
//     commit(w) = sum_j value_j * L[j]      with      L_j(X) = (1/n) sum_k w_n^(-jk) X^k,
// i.e. L = the inverse DFT of (P_0 .. P_(n-1)) taken in the group: n/2 log n butterflies (u, v) -> (u + v, w^t (u - v)), each with
// one 254-bit scalar multiplication.  That is a one-off cost per (key, n) -- about a second at n = 2^20 -- after which a wire
// commitment is an MSM whose scalars are witness values: zeros cost nothing and small values a few digits.  The commitments, and
// so every proof byte, are the same.  The two masking terms of a wire polynomial, (b_0 + b_1 X)(X^n - 1) (prover.rs:463-486), are
// commit-key points too: the key built here carries P_n - P_0 and P_(n+1) - P_1 behind the n Lagrange points.
#include "common.cuh"
#include "ec.cuh"
#include "field.cuh"

namespace jf {

template <class F> __device__ __forceinline__ Fp<F> lag_pow(Fp<F> base, uint32_t e) {
    Fp<F> r = Fp<F>::one();
    while (e) {
        if (e & 1u) r = Fp<F>::mul(r, base);
        base = Fp<F>::sqr(base);
        e >>= 1;
    }
    return r;
}

// tw[t] = w^t as a canonical scalar (8 x 32-bit words, also for the 255-bit field), t < count
template <class Fr> __global__ void lag_twiddle_kernel(uint32_t *tw, Fp<Fr> w_mont, uint32_t count) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const Fp<Fr> c = Fp<Fr>::from_mont(lag_pow(w_mont, t));
#pragma unroll
    for (int i = 0; i < 8; i++) tw[8 * (size_t)t + i] = c.v[i];
}

// k * P for a point in XYZZ form: fixed 4-bit windows, left to right.  Every lane of a warp carries its own scalar, so a
// bit-by-bit double-and-add pays for an addition at almost every bit (some lane always has the bit set); with a 15-entry table
// (local memory, 128 B per entry) the warp does 254 doublings and 64 additions whatever the scalars are.
template <class Fq> __device__ XYZZ<Fq> xyzz_mul(const XYZZ<Fq> &p, const uint32_t *k) {
    XYZZ<Fq> acc = XYZZ<Fq>::inf();
    if (p.is_inf()) return acc;
    XYZZ<Fq> tab[15];  // tab[i] = (i + 1) P
    tab[0] = p;
#pragma unroll 1
    for (int i = 1; i < 15; i++) {
        if (i & 1) tab[i] = tab[i >> 1].dbl();  // (i + 1) even: 2 * ((i + 1) / 2) P
        else {
            tab[i] = tab[i - 1];
            tab[i].add(p);
        }
    }
    int w = 63;
    while (w >= 0 && !((k[w >> 3] >> ((w & 7) * 4)) & 15u)) w--;
#pragma unroll 1
    for (; w >= 0; w--) {
        if (!acc.is_inf()) {
            acc = acc.dbl();
            acc = acc.dbl();
            acc = acc.dbl();
            acc = acc.dbl();
        }
        const uint32_t d = (k[w >> 3] >> ((w & 7) * 4)) & 15u;
        if (d) acc.add(tab[d - 1]);
    }
    return acc;
}

// X[k] = ninv * P_k (the 1/n of the inverse transform goes in with the inputs)
template <class Fq> __global__ void __launch_bounds__(64) lag_load_kernel(const Affine<Fq> *pts, XYZZ<Fq> *X, uint32_t n, const uint32_t *ninv) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t s[8];
#pragma unroll
    for (int i = 0; i < 8; i++) s[i] = ninv[i];
    X[k] = scalar_mul(pts[k], s, 8);
}

// one decimation-in-frequency stage: blocks of 2h, (u, v) = (X[j], X[j + h]) -> (u + v, w^(j stride) (u - v))
template <class Fq> __global__ void __launch_bounds__(64) lag_stage_kernel(XYZZ<Fq> *X, uint32_t n, uint32_t h, const uint32_t *tw, uint32_t stride) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n / 2) return;
    const uint32_t j = idx % h, i0 = (idx / h) * 2 * h + j, i1 = i0 + h;
    const XYZZ<Fq> u = X[i0], v = X[i1];
    XYZZ<Fq> s = u, d = u;
    s.add(v);
    d.add(v.neg());
    const uint32_t t = j * stride;
    if (t) d = xyzz_mul(d, tw + 8 * (size_t)t);
    X[i0] = s;
    X[i1] = d;
}

// natural order out of the bit-reversed result, normalised (`into_affine`)
template <class Fq> __global__ void __launch_bounds__(64) lag_finish_kernel(const XYZZ<Fq> *X, Affine<Fq> *out, uint32_t n, int log_n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t r = log_n ? __brev(i) >> (32 - log_n) : 0;
    out[r] = X[i].to_affine();
}

// out[t] = P_(n + t) - P_t, t < 2: the commit-key side of the masking terms b_t X^t (X^n - 1)
template <class Fq> __global__ void lag_mask_points_kernel(const Affine<Fq> *pts, Affine<Fq> *out, uint32_t n) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2) return;
    XYZZ<Fq> a = XYZZ<Fq>::from_affine(pts[n + t]);
    a.add(XYZZ<Fq>::from_affine(pts[t]).neg());
    out[t] = a.to_affine();
}

template <class C> static int srs_lagrange_t(jf_ctx *ctx, const jf_srs *mono, unsigned log_n, int mask_points, jf_srs **out) {
    using Fq = typename C::Fq;
    using Fr = typename C::Fr;
    using E = Fp<Fr>;
    if (log_n > (unsigned)Fr::TWO_ADICITY || log_n > 25) return fail(ctx, JF_ERR_DOMAIN_TOO_LARGE, "srs_lagrange: unsupported domain size");
    const size_t n = (size_t)1 << log_n, total = n + (mask_points ? 2 : 0);
    if (mono->n < total) return fail(ctx, JF_ERR_INVALID_ARG, "srs_lagrange: the commit key holds fewer than n (+ 2) points");
    cudaStream_t st = ctx->stream;
    // w_n^-1 and n^-1
    uint32_t e[8];
    Limbs<Fr>::p(e);
    e[0] -= 1;
    for (int s = 0; s < Fr::TWO_ADICITY; s++)
        for (int i = 0; i < 8; i++) e[i] = (e[i] >> 1) | (i < 7 ? e[i + 1] << 31 : 0);
    E w = E::pow(E::from_u32(Fr::GENERATOR), e, 8);
    for (unsigned s = 0; s < Fr::TWO_ADICITY - log_n; s++) w = E::sqr(w);
    const E w_inv = E::inv(w);
    const E ninv = E::from_mont(E::inv(E::from_u32((uint32_t)n)));
    void *dX, *dtw, *dout, *dninv;
    JF_TRY(scratch(ctx, "lag_x", sizeof(XYZZ<Fq>) * n, &dX));
    JF_TRY(scratch(ctx, "lag_tw", 32 * (n / 2 + 1), &dtw));
    JF_TRY(scratch(ctx, "lag_out", sizeof(Affine<Fq>) * total, &dout));
    JF_TRY(scratch(ctx, "lag_ninv", 64, &dninv));
    const Affine<Fq> *pts = (const Affine<Fq> *)mono->d_points;  // table 0 = the key points themselves
    XYZZ<Fq> *X = (XYZZ<Fq> *)dX;
    JF_CUDA(ctx, cudaMemcpyAsync(dninv, ninv.v, 32, cudaMemcpyHostToDevice, st));
    if (n > 1) JF_LAUNCH(ctx, "lag_twiddle", lag_twiddle_kernel<Fr><<<(unsigned)((n / 2 + 127) / 128), 128, 0, st>>>((uint32_t *)dtw, w_inv, (uint32_t)(n / 2)));
    JF_LAUNCH(ctx, "lag_load", lag_load_kernel<Fq><<<(unsigned)((n + 63) / 64), 64, 0, st>>>(pts, X, (uint32_t)n, (const uint32_t *)dninv));
    for (size_t h = n / 2; h >= 1; h >>= 1)
        JF_LAUNCH(ctx, "lag_stage", lag_stage_kernel<Fq><<<(unsigned)((n / 2 + 63) / 64), 64, 0, st>>>(X, (uint32_t)n, (uint32_t)h, (const uint32_t *)dtw,
                                                                                            (uint32_t)(n / (2 * h))));
    JF_LAUNCH(ctx, "lag_finish", lag_finish_kernel<Fq><<<(unsigned)((n + 63) / 64), 64, 0, st>>>(X, (Affine<Fq> *)dout, (uint32_t)n, (int)log_n));
    if (mask_points) JF_LAUNCH(ctx, "lag_mask_points", lag_mask_points_kernel<Fq><<<1, 32, 0, st>>>(pts, (Affine<Fq> *)dout + n, (uint32_t)n));
    // same window as the monomial key: the two keys then share the MSM's bucket geometry
    JF_TRY(srs_build(ctx, mono->curve, dout, total, mono->window_bits, mono->tables > 1 ? 1 : 0, out));
    (*out)->skew = 1;  // witness values, not random-looking coefficients, meet this key
    return JF_OK;
}

int srs_lagrange(jf_ctx *ctx, const jf_srs *mono, unsigned log_n, int mask_points, jf_srs **out) {
    if (mono->curve == JF_BN254) return srs_lagrange_t<Bn254G1>(ctx, mono, log_n, mask_points, out);
    if (mono->curve == JF_BLS12_381) return srs_lagrange_t<Bls12381G1>(ctx, mono, log_n, mask_points, out);
    return fail(ctx, JF_ERR_INVALID_ARG, "srs_lagrange: unknown curve");
}

}  // namespace jf

using namespace jf;

extern "C" int jf_srs_lagrange(jf_ctx *ctx, const jf_srs *srs, unsigned log_n, int mask_points, jf_srs **out) {
    if (!ctx) return JF_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    cudaSetDevice(ctx->device);
    if (!srs || !out) return fail(ctx, JF_ERR_INVALID_ARG, "srs_lagrange: null argument");
    *out = nullptr;
    int rc = srs_lagrange(ctx, srs, log_n, mask_points, out);
    if (rc != JF_OK) cudaStreamSynchronize(ctx->stream);
    return rc;
}
