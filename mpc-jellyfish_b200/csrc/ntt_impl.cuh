// ntt.cu -- radix-2 NTT / coset NTT over BN254 Fr and BLS12-381 Fr (K3/K4 of SURVEY.md §2.3).
//
// Replaces ark-poly 0.4.2 `Radix2EvaluationDomain::{fft,ifft}_in_place` and the
// `get_coset(GENERATOR)` variants as called from
//   relation/src/constraint_system.rs:1172,1189,1221,1240,1257   (ifft, size n)
//   plonk/src/proof_system/prover.rs:545,552-567                 (25 coset ffts, size 8n)
//   plonk/src/proof_system/prover.rs:672                         (coset ifft, size 8n)
// Natural order in, natural order out.
//
// Algorithm: Stockham autosort, 1-4 passes of radix R = 2^K (K <= 9), ping-ponging between the
// caller's vector and one scratch vector.  One pass = for every j in [0, n/R): gather the R
// inputs x[j + r*n/R], multiply by the inter-pass twiddle w_{Ns*R}^{(j mod Ns) r}, do an
// R-point DIF transform and scatter to y[expand(j) + f*Ns].  A CTA owns a tile of R rows x B
// adjacent columns (R*B = 2048 elements, 64 KB of shared memory): every global access is a run
// of B*32 contiguous bytes, every thread keeps 8 elements (64 registers) and does three
// butterfly layers per shared-memory round trip.  Coset scaling (x_j * g^j) is fused into the
// first pass' loads, zero padding (in_len < n) costs no reads, and the inverse transform's
// n^-1 * g^-j scaling is fused into the last pass' stores.
#pragma once
#include <atomic>
#include "common.cuh"
#include "field.cuh"

namespace jf {

static constexpr int NTT_TILE_LOG = 11;  // R * B = 2048 elements per CTA
static constexpr int NTT_MAX_K = 9;
static constexpr int NTT_MIN_K = 3;

struct NttPass {
    int k;       // log2 radix
    int log_ns;  // log2 of the product of the previous radices
};

struct NttPlan {
    int field;
    unsigned log_n;
    bool inverse;
    bool has_offset;
    std::vector<NttPass> passes;
    int split;                    // two-level power tables: e = hi * 2^split + lo
    void *w_lo = nullptr, *w_hi = nullptr;      // powers of w_n (or its inverse)
    void *off_lo = nullptr, *off_hi = nullptr;  // powers of offset (fwd) / offset^-1 scaled by n^-1 (inv)
    void *tw[NTT_MAX_K + 1] = {nullptr};        // tw[k][i] = w_{2^k}^i, i < 2^(k-1)
    void *n_inv = nullptr;                      // single element n^-1 (plain inverse)
    // inter-pass twiddles, one multiplication per element: ptw[i][(r << log_ns) | k] = w^((k r) n / (Ns R)),
    // times n^-1 on the last pass of a plain inverse transform (the scaling rides along for free)
    std::vector<void *> ptw;
    void *off_full = nullptr;                   // offset^j (fwd) / n^-1 offset^-j (inv), j < n; built on first dense use
    bool ninv_folded = false;
    // ntt_run_cosets: `rows` offset tables of n entries in off_full, X^n on each coset in fold_c
    int rows = 0;
    void *fold_c = nullptr;
    uint64_t last_use = 0;  // cache clock of the last call that used this plan (LRU eviction)
};

inline void free_plan(NttPlan *pl) {
    cudaFree(pl->w_lo);
    cudaFree(pl->w_hi);
    cudaFree(pl->off_lo);
    cudaFree(pl->off_hi);
    cudaFree(pl->n_inv);
    cudaFree(pl->off_full);
    cudaFree(pl->fold_c);
    for (auto &t : pl->ptw) cudaFree(t);
    for (auto &t : pl->tw) cudaFree(t);
    delete pl;
}

// The cache is keyed by the coset offset as well, so a caller that varies the offset per call would otherwise grow it
// without bound: beyond NTT_PLAN_CAP plans the least recently used ones are dropped (after a device synchronisation:
// kernels of any lane may still read their tables).  `keep` is never dropped.
static constexpr size_t NTT_PLAN_CAP = 48;
inline NttPlan *plan_touch(jf_ctx *ctx, NttPlan *pl) {
    pl->last_use = ++ctx->ntt_plan_clock;
    return pl;
}
inline void plan_cache_insert(jf_ctx *ctx, const std::string &key, NttPlan *pl, const NttPlan *keep = nullptr) {
    plan_touch(ctx, pl);
    ctx->ntt_plans[key] = pl;
    if (ctx->ntt_plans.size() <= NTT_PLAN_CAP) return;
    cudaDeviceSynchronize();
    while (ctx->ntt_plans.size() > NTT_PLAN_CAP * 3 / 4) {
        auto victim = ctx->ntt_plans.end();
        for (auto it = ctx->ntt_plans.begin(); it != ctx->ntt_plans.end(); ++it)
            if (it->second != pl && it->second != keep && (victim == ctx->ntt_plans.end() || it->second->last_use < victim->second->last_use))
                victim = it;
        if (victim == ctx->ntt_plans.end()) break;
        free_plan(victim->second);
        ctx->ntt_plans.erase(victim);
    }
}

template <class F> struct El {  // 32-byte element as two 16-byte halves for vector loads
    uint4 lo, hi;
};

template <class F> __device__ __forceinline__ Fp<F> ld_el(const Fp<F> *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 a = q[0], b = q[1];
    Fp<F> r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
template <class F> __device__ __forceinline__ Fp<F> ldg_el(const Fp<F> *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    Fp<F> r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
template <class F> __device__ __forceinline__ void st_el(Fp<F> *p, const Fp<F> &r) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

// g^e from a two-level table
template <class F>
__device__ __forceinline__ Fp<F> pow2l(const Fp<F> *lo, const Fp<F> *hi, uint64_t e, int split) {
    uint32_t l = (uint32_t)(e & ((1ull << split) - 1));
    uint32_t h = (uint32_t)(e >> split);
    Fp<F> a = ldg_el(lo + l);
    if (h == 0) return a;
    return Fp<F>::mul(a, ldg_el(hi + h));
}

struct PassArgs {
    const void *src;
    void *dst;
    size_t stride;     // elements between batch entries
    size_t in_len;     // first pass only: entries >= in_len read as zero
    uint32_t batch;
    uint32_t log_n;
    uint32_t log_ns;
    int first, last;   // pass position
    int scale_mode;    // on last pass: 0 none, 1 multiply by *n_inv, 2 multiply by off table at output index
    int coset_in;      // on first pass: multiply input j by off table
    int split;
    const void *w_lo, *w_hi, *off_lo, *off_hi, *tw, *n_inv;
    const void *ptw;       // this pass' inter-pass twiddle table (passes > 0)
    const void *off_full;  // full offset power table or NULL (then the two-level tables are used)
    // several cosets of one polynomial (ntt_run_cosets): batch entry `bid` is coset row bid % off_rows, whose
    // offset powers are off_full + row * n; on the first pass src_group consecutive entries read the same
    // source vector, and coefficients n.. are folded in with X^n = fold_c[row] (reduction mod X^n - c)
    size_t src_stride;
    uint32_t src_group, off_rows;
    const void *fold_c;
};

template <int K> __device__ __forceinline__ uint32_t bitrev_k(uint32_t x) { return __brev(x) >> (32 - K); }

// One Stockham pass of radix 2^K.  blockDim = 256 = (R/8) * B.
// FIRST / LAST are compile-time so that every instantiation carries only the twiddle / scaling code it uses
// (the fully unrolled kernel otherwise overflows the instruction cache: 11 % no-instruction stalls in ncu).
// MULTI: a batch of several cosets per polynomial (ntt_run_cosets); the plain transforms compile without that code.
template <class F, int K, bool FIRST, bool LAST, bool MULTI> __global__ void __launch_bounds__(256, 2) ntt_pass_kernel(PassArgs a) {
    using E = Fp<F>;
    constexpr int R = 1 << K;
    constexpr int LOGB = NTT_TILE_LOG - K;
    constexpr int B = 1 << LOGB;
    constexpr int G = (K + 2) / 3;          // register groups
    constexpr int PLANE = (R + R / 8) * B;  // uint4 per plane, one padding row per 8 rows
    extern __shared__ uint4 smem[];
    uint4 *pl0 = smem, *pl1 = smem + PLANE;

    const uint32_t t = threadIdx.x;
    const uint32_t b = t & (B - 1);
    const uint32_t q = t >> LOGB;  // < R/8
    const uint32_t log_cols = a.log_n - K;
    const uint64_t cols = 1ull << log_cols;
    // consecutive CTAs work on the same column tile of different vectors: they share the twiddle tile in L2
    const uint64_t bid = blockIdx.x % a.batch;
    const uint64_t j = (uint64_t)(blockIdx.x / a.batch) * B + b;
    const bool valid = j < cols;
    const E *src = reinterpret_cast<const E *>(a.src) + bid * a.stride;
    E *dst = reinterpret_cast<E *>(a.dst) + bid * a.stride;
    const E *tw = reinterpret_cast<const E *>(a.tw);
    // several cosets of one polynomial (only the first / last pass of such a transform takes this branch):
    // this entry's offset table and, on the first pass, the source vector its group shares
    const E *off_full = reinterpret_cast<const E *>(a.off_full);
    uint32_t orow = 0;
    if (MULTI) {
        const uint32_t bid32 = (uint32_t)bid;
        orow = bid32 % a.off_rows;
        off_full += (size_t)orow << a.log_n;
        if (FIRST) src = reinterpret_cast<const E *>(a.src) + (size_t)(bid32 / a.src_group) * a.src_stride;
    }

    E v[8];
    // ---- gather (group 0 mapping: row = e << (K-3) | q) ------------------------------------
    {
        constexpr int S0 = K - 3;
        const uint64_t k = j & ((1ull << a.log_ns) - 1);
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const uint32_t r = ((uint32_t)e << S0) | q;
            const uint64_t idx = j + (uint64_t)r * cols;
            if (valid && (!FIRST || idx < a.in_len)) {
                v[e] = ld_el(src + idx);
                if (FIRST) {
                    if (MULTI && a.fold_c && idx + (1ull << a.log_n) < a.in_len)
                        v[e] = E::add(v[e], E::mul(ld_el(src + idx + (1ull << a.log_n)), ldg_el((const E *)a.fold_c + orow)));
                    if (a.coset_in)
                        v[e] = E::mul(v[e], a.off_full ? ldg_el(off_full + idx)
                                                       : pow2l((const E *)a.off_lo, (const E *)a.off_hi, idx, a.split));
                } else {
                    v[e] = E::mul(v[e], ldg_el((const E *)a.ptw + (((uint64_t)r << a.log_ns) | k)));
                }
            } else {
                v[e] = E::zero();
            }
        }
    }
    // ---- butterfly groups --------------------------------------------------------------
#pragma unroll
    for (int g = 0; g < G; g++) {
        const int S = (g < G - 1 || K % 3 == 0) ? K - 3 * (g + 1) : 0;      // row bit of element-index bit 0
        const int C = (g < G - 1 || K % 3 == 0) ? 3 : K - 3 * (G - 1);      // layers in this group
        const uint32_t q_lo = q & ((1u << S) - 1), q_hi = q >> S;
        if (g > 0) {
            // load this group's rows from shared memory
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const uint32_t r = (q_hi << (S + 3)) | ((uint32_t)e << S) | q_lo;
                const uint32_t ad = (r + (r >> 3)) * B + b;
                uint4 x = pl0[ad], y = pl1[ad];
                v[e].v[0] = x.x; v[e].v[1] = x.y; v[e].v[2] = x.z; v[e].v[3] = x.w;
                v[e].v[4] = y.x; v[e].v[5] = y.y; v[e].v[6] = y.z; v[e].v[7] = y.w;
            }
        }
#pragma unroll
        for (int bit = C - 1; bit >= 0; bit--) {
            const int p = S + bit;  // row bit handled by this layer; partner rows differ by 2^p
#pragma unroll
            for (int e0 = 0; e0 < 8; e0++) {
                if (e0 & (1 << bit)) continue;
                const int e1 = e0 | (1 << bit);
                E lo = v[e0], hi = v[e1];
                v[e0] = E::add(lo, hi);
                E d = E::sub(lo, hi);
                if (p == 0) {
                    v[e1] = d;
                } else {
                    const uint32_t rmod = (((uint32_t)e0 & ((1u << bit) - 1)) << S) | q_lo;  // row mod 2^p
                    const uint32_t ti = rmod << (K - 1 - p);
                    v[e1] = ti == 0 ? d : E::mul(d, ldg_el(tw + ti));
                }
            }
        }
        if (g < G - 1) {
            if (g > 0) __syncthreads();  // everyone has read the previous exchange
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const uint32_t r = (q_hi << (S + 3)) | ((uint32_t)e << S) | q_lo;
                const uint32_t ad = (r + (r >> 3)) * B + b;
                pl0[ad] = make_uint4(v[e].v[0], v[e].v[1], v[e].v[2], v[e].v[3]);
                pl1[ad] = make_uint4(v[e].v[4], v[e].v[5], v[e].v[6], v[e].v[7]);
            }
            __syncthreads();
        }
    }
    // ---- scatter: last group's mapping is row = 8 q + e (S = 0) ------------------------------
    if (!valid) return;
    {
        const uint64_t ns_mask = (1ull << a.log_ns) - 1;
        const uint64_t j0 = ((j >> a.log_ns) << (a.log_ns + K)) | (j & ns_mask);
        const E *ninv = reinterpret_cast<const E *>(a.n_inv);
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const uint32_t r = (q << 3) | (uint32_t)e;
            const uint32_t f = bitrev_k<K>(r);
            const uint64_t idx = j0 + ((uint64_t)f << a.log_ns);
            E o = v[e];
            if (LAST) {
                if (a.scale_mode == 1) o = E::mul(o, ldg_el(ninv));
                else if (a.scale_mode == 2)
                    o = E::mul(o, ldg_el(off_full + idx));  // ntt_run always builds it for inverses
            }
            st_el(dst + idx, o);
        }
    }
}

// Sizes 1, 2, 4: direct evaluation, one thread per vector.
template <class F> __global__ void ntt_tiny_kernel(Fp<F> *data, size_t stride, uint32_t batch, uint32_t log_n,
                                                  size_t in_len, int inverse, const Fp<F> *w_lo,
                                                  const Fp<F> *off_lo, int has_off, const Fp<F> *n_inv) {
    using E = Fp<F>;
    uint32_t bid = blockIdx.x * blockDim.x + threadIdx.x;
    if (bid >= batch) return;
    E *x = data + (size_t)bid * stride;
    const uint32_t n = 1u << log_n;
    E in[4], out[4];
    for (uint32_t i = 0; i < n; i++) {
        in[i] = i < in_len ? x[i] : E::zero();
        if (!inverse && has_off) in[i] = E::mul(in[i], off_lo[i]);
    }
    for (uint32_t i = 0; i < n; i++) {
        E acc = E::zero();
        for (uint32_t jx = 0; jx < n; jx++) acc = E::add(acc, E::mul(in[jx], w_lo[(i * jx) & (n - 1)]));
        out[i] = acc;
    }
    for (uint32_t i = 0; i < n; i++) {
        if (inverse) out[i] = has_off ? E::mul(out[i], off_lo[i]) : E::mul(out[i], *n_inv);
        x[i] = out[i];
    }
}

// table[i] = scale * base^(i * step_mul), i < count
template <class F>
__global__ void pow_table_kernel(Fp<F> *table, Fp<F> base, Fp<F> scale, uint64_t step_mul, uint32_t count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    table[i] = Fp<F>::mul(scale, Fp<F>::pow_u64(base, (uint64_t)i * step_mul));
}

// ptw[(r << log_ns) | k] = scale * w^((k r) << tw_shift), from the two-level power tables
template <class F>
__global__ void pass_table_kernel(Fp<F> *out, const Fp<F> *w_lo, const Fp<F> *w_hi, int split, uint32_t log_ns, uint32_t k_bits,
                                  uint32_t tw_shift, Fp<F> scale, int use_scale) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >> (log_ns + k_bits)) return;
    const uint64_t k = i & ((1ull << log_ns) - 1), r = i >> log_ns;
    Fp<F> t = pow2l(w_lo, w_hi, (k * r) << tw_shift, split);
    if (use_scale) t = Fp<F>::mul(t, scale);
    st_el(out + i, t);
}
// full[i] = lo[i & mask] * hi[i >> split]  (lo already carries any constant factor)
template <class F>
__global__ void full_table_kernel(Fp<F> *out, const Fp<F> *lo, const Fp<F> *hi, int split, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st_el(out + i, pow2l(lo, hi, i, split));
}

// ---- host-side field helpers (serial, a handful of operations per plan) -----------------------
template <class F> static Fp<F> host_root_of_unity(unsigned log_n) {
    // two_adic_root = GENERATOR^((p-1)/2^s); w_n = two_adic_root^(2^(s - log_n))
    uint32_t e[8];
    Limbs<F>::p(e);
    e[0] -= 1;  // p - 1 (p odd, low limb >= 1)
    for (int k = 0; k < F::TWO_ADICITY; k++) {
        for (int i = 0; i < 8; i++) e[i] = (e[i] >> 1) | (i < 7 ? e[i + 1] << 31 : 0);
    }
    Fp<F> g = Fp<F>::from_u32(F::GENERATOR);
    Fp<F> w = Fp<F>::pow(g, e, 8);
    for (unsigned k = 0; k < F::TWO_ADICITY - log_n; k++) w = Fp<F>::sqr(w);
    return w;
}

template <class F> static int build_plan(jf_ctx *ctx, NttPlan *pl, const uint64_t *coset_offset) {
    using E = Fp<F>;
    const unsigned log_n = pl->log_n;
    E w = host_root_of_unity<F>(log_n);
    if (pl->inverse) w = E::inv(w);
    E ninv = E::one();  // n^-1 = (2^-1)^log_n
    {
        E half = E::inv(E::from_u32(2));
        for (unsigned i = 0; i < log_n; i++) ninv = E::mul(ninv, half);
    }
    E off = E::one();
    if (pl->has_offset) {
        for (int i = 0; i < 4; i++) {
            off.v[2 * i] = (uint32_t)coset_offset[i];
            off.v[2 * i + 1] = (uint32_t)(coset_offset[i] >> 32);
        }
        if (pl->inverse) off = E::inv(off);
    }
    auto table = [&](void **dptr, E base, E scale, uint64_t step, uint32_t count) -> int {
        JF_CUDA(ctx, cudaMalloc(dptr, sizeof(E) * (size_t)count));
        JF_LAUNCH(ctx, "pow_table", pow_table_kernel<F><<<(count + 127) / 128, 128, 0, ctx->stream>>>((E *)*dptr, base, scale, step, count));
        return JF_OK;
    };
    // the inverse transform's n^-1 rides in the low table of the output scaling (hi[0] = 1 is skipped)
    const E out_scale = pl->inverse ? ninv : E::one();
    pl->passes.clear();
    if (log_n < NTT_MIN_K) {  // tiny path: full tables of n entries
        pl->split = log_n;
        JF_TRY(table(&pl->w_lo, w, E::one(), 1, 1u << log_n));
        if (pl->has_offset) JF_TRY(table(&pl->off_lo, off, out_scale, 1, 1u << log_n));
    } else {
        const int np = (log_n + NTT_MAX_K - 1) / NTT_MAX_K;
        const int base = log_n / np, rem = log_n % np;
        int acc = 0;
        for (int i = 0; i < np; i++) {
            int k = base + (i < rem ? 1 : 0);
            pl->passes.push_back({k, acc});
            acc += k;
        }
        pl->split = (log_n + 1) / 2;
        const uint32_t lo_n = 1u << pl->split, hi_n = 1u << (log_n - pl->split);
        JF_TRY(table(&pl->w_lo, w, E::one(), 1, lo_n));
        JF_TRY(table(&pl->w_hi, w, E::one(), lo_n, hi_n));
        if (pl->has_offset) {
            JF_TRY(table(&pl->off_lo, off, out_scale, 1, lo_n));
            JF_TRY(table(&pl->off_hi, off, E::one(), lo_n, hi_n));
        }
        const int npass = (int)pl->passes.size();
        pl->ptw.assign(npass, nullptr);
        pl->ninv_folded = pl->inverse && !pl->has_offset && npass >= 2;
        for (int i = 1; i < npass; i++) {
            const NttPass &ps = pl->passes[i];
            const uint64_t count = 1ull << (ps.log_ns + ps.k);
            JF_CUDA(ctx, cudaMalloc(&pl->ptw[i], sizeof(E) * count));
            const bool fold = pl->ninv_folded && i == npass - 1;
            JF_LAUNCH(ctx, "pass_table", pass_table_kernel<F><<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(
                (E *)pl->ptw[i], (const E *)pl->w_lo, (const E *)pl->w_hi, pl->split, (uint32_t)ps.log_ns, (uint32_t)ps.k,
                (uint32_t)(log_n - ps.log_ns - ps.k), ninv, fold ? 1 : 0));
        }
        for (auto &ps : pl->passes) {
            if (pl->tw[ps.k]) continue;
            // w_{2^k}^i = w^(i * n / 2^k)
            JF_TRY(table(&pl->tw[ps.k], w, E::one(), 1ull << (log_n - ps.k), 1u << (ps.k - 1)));
        }
    }
    JF_CUDA(ctx, cudaMalloc(&pl->n_inv, sizeof(E)));
    JF_CUDA(ctx, cudaMemcpyAsync(pl->n_inv, &ninv, sizeof(E), cudaMemcpyHostToDevice, ctx->stream));
    JF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // ninv is a stack temporary
    return JF_OK;
}

template <class F, int K, bool FIRST, bool LAST, bool MULTI> static int launch_pass_fl(jf_ctx *ctx, const PassArgs &a) {
    constexpr int R = 1 << K;
    constexpr int B = 1 << (NTT_TILE_LOG - K);
    constexpr size_t smem = (size_t)2 * (R + R / 8) * B * sizeof(uint4);
    // the attribute is per device: one bit per device, set after the (idempotent) call so that concurrent first calls are harmless
    static std::atomic<uint64_t> configured{0};
    const uint64_t bit = (uint64_t)1 << (ctx->device & 63);
    if (!(configured.load(std::memory_order_acquire) & bit)) {
        JF_CUDA(ctx, cudaFuncSetAttribute(ntt_pass_kernel<F, K, FIRST, LAST, MULTI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.fetch_or(bit, std::memory_order_release);
    }
    const uint64_t cols = (uint64_t)1 << (a.log_n - K);
    const uint64_t blocks = ((cols + B - 1) / B) * a.batch;
    if (blocks > 0x7fffffffull) return fail(ctx, JF_ERR_INVALID_ARG, "ntt: batch too large");
    JF_LAUNCH(ctx, "ntt_pass", ntt_pass_kernel<F, K, FIRST, LAST, MULTI><<<(unsigned)blocks, 256, smem, ctx->stream>>>(a));
    return JF_OK;
}

template <class F, int K> static int launch_pass(jf_ctx *ctx, const PassArgs &a) {
    if (a.off_rows != 0) {  // only the first and last pass differ for a multi-coset batch
        if (a.first && a.last) return launch_pass_fl<F, K, true, true, true>(ctx, a);
        if (a.first) return launch_pass_fl<F, K, true, false, true>(ctx, a);
        if (a.last) return launch_pass_fl<F, K, false, true, true>(ctx, a);
    }
    if (a.first && a.last) return launch_pass_fl<F, K, true, true, false>(ctx, a);
    if (a.first) return launch_pass_fl<F, K, true, false, false>(ctx, a);
    if (a.last) return launch_pass_fl<F, K, false, true, false>(ctx, a);
    return launch_pass_fl<F, K, false, false, false>(ctx, a);
}

template <class F> static int launch_pass_k(jf_ctx *ctx, int k, const PassArgs &a) {
    switch (k) {
    case 3: return launch_pass<F, 3>(ctx, a);
    case 4: return launch_pass<F, 4>(ctx, a);
    case 5: return launch_pass<F, 5>(ctx, a);
    case 6: return launch_pass<F, 6>(ctx, a);
    case 7: return launch_pass<F, 7>(ctx, a);
    case 8: return launch_pass<F, 8>(ctx, a);
    case 9: return launch_pass<F, 9>(ctx, a);
    }
    return fail(ctx, JF_ERR_INVALID_ARG, "ntt: bad radix");
}

// cached plan of (field, log n, direction, offset); an offset equal to one is the plain domain
template <class F>
static int get_plan(jf_ctx *ctx, int field, unsigned log_n, int inverse, const uint64_t *coset_offset, NttPlan **out, bool *out_has_off) {
    using E = Fp<F>;
    bool has_off = false;
    if (coset_offset) {
        E one = E::one();
        for (int i = 0; i < 4; i++)
            if ((uint32_t)coset_offset[i] != one.v[2 * i] || (uint32_t)(coset_offset[i] >> 32) != one.v[2 * i + 1]) has_off = true;
    }
    char key[160];
    snprintf(key, sizeof key, "%d/%u/%d/%d/%016llx%016llx%016llx%016llx", field, log_n, inverse ? 1 : 0, has_off ? 1 : 0,
             has_off ? (unsigned long long)coset_offset[0] : 0ull, has_off ? (unsigned long long)coset_offset[1] : 0ull,
             has_off ? (unsigned long long)coset_offset[2] : 0ull, has_off ? (unsigned long long)coset_offset[3] : 0ull);
    NttPlan *pl;
    auto it = ctx->ntt_plans.find(key);
    if (it == ctx->ntt_plans.end()) {
        pl = new NttPlan();
        pl->field = field;
        pl->log_n = log_n;
        pl->inverse = inverse != 0;
        pl->has_offset = has_off;
        int rc = build_plan<F>(ctx, pl, coset_offset);
        if (rc != JF_OK) {
            free_plan(pl);
            return rc;
        }
        plan_cache_insert(ctx, key, pl);
    } else {
        pl = plan_touch(ctx, it->second);
    }
    *out = pl;
    *out_has_off = has_off;
    return JF_OK;
}

template <class F>
static int ntt_run_t(jf_ctx *ctx, int field, void *d_data, void *d_out, size_t in_len, unsigned log_n, int inverse,
                     const uint64_t *coset_offset, size_t batch, size_t batch_stride) {
    using E = Fp<F>;
    if (log_n > (unsigned)F::TWO_ADICITY) return fail(ctx, JF_ERR_DOMAIN_TOO_LARGE, "ntt: log_n exceeds the field's two-adicity");
    if (log_n > 30) return fail(ctx, JF_ERR_NOMEM, "ntt: log_n > 30 is not supported");
    const size_t n = (size_t)1 << log_n;
    if (in_len > n) return fail(ctx, JF_ERR_INVALID_ARG, "ntt: in_len > domain size");
    if (batch == 0) return JF_OK;
    if (batch > 1 && batch_stride < n) return fail(ctx, JF_ERR_INVALID_ARG, "ntt: batch_stride < domain size");
    if (batch >= (1u << 30)) return fail(ctx, JF_ERR_INVALID_ARG, "ntt: batch too large");
    NttPlan *pl;
    bool has_off;
    JF_TRY(get_plan<F>(ctx, field, log_n, inverse, coset_offset, &pl, &has_off));
    E *data = reinterpret_cast<E *>(d_data);
    if (log_n < NTT_MIN_K) {
        JF_LAUNCH(ctx, "ntt_tiny", ntt_tiny_kernel<F><<<(unsigned)((batch + 63) / 64), 64, 0, ctx->stream>>>(
            data, batch_stride, (uint32_t)batch, log_n, in_len, inverse, (const E *)pl->w_lo, (const E *)pl->off_lo,
            has_off ? 1 : 0, (const E *)pl->n_inv));
        if (d_out != d_data)
            JF_CUDA(ctx, cudaMemcpy2DAsync(d_out, batch_stride * sizeof(E), d_data, batch_stride * sizeof(E), n * sizeof(E),
                                           batch, cudaMemcpyDeviceToDevice, ctx->stream));
        return JF_OK;
    }
    // Buffer chain: pass i reads hop[i] and writes hop[i+1]; hop[0] = input, hop[np] = output.
    // Out of place (input may be clobbered): odd np alternates in/out, even np detours through one
    // scratch vector.  In place: even np bounces off one scratch vector, odd np needs two.
    const int np = (int)pl->passes.size();
    const size_t span = (batch - 1) * batch_stride + n;
    void *t1 = nullptr, *t2 = nullptr;
    std::vector<void *> hop(np + 1);
    hop[0] = d_data;
    hop[np] = d_out;
    const bool in_place = d_out == d_data;
    if (in_place) {
        JF_TRY(scratch(ctx, "ntt_t1", span * sizeof(E), &t1));
        if (np % 2 == 1) JF_TRY(scratch(ctx, "ntt_t2", span * sizeof(E), &t2));
        if (np == 1) {
            hop[1] = t1;
        } else if (np % 2 == 0) {
            for (int i = 1; i < np; i++) hop[i] = (i % 2 == 1) ? t1 : d_data;
        } else {
            for (int i = 1; i < np; i++) hop[i] = (i % 2 == 1) ? t1 : t2;
        }
    } else if (np % 2 == 1) {
        for (int i = 1; i < np; i++) hop[i] = (i % 2 == 1) ? d_out : d_data;
    } else {
        JF_TRY(scratch(ctx, "ntt_t1", span * sizeof(E), &t1));
        for (int i = 1; i < np; i++) hop[i] = (i % 2 == 1) ? t1 : d_data;
    }
    // dense coset input / coset output scaling: one multiplication per element from a full table
    if (has_off && !pl->off_full && (inverse || in_len > n / 4)) {
        JF_CUDA(ctx, cudaMalloc(&pl->off_full, sizeof(E) * n));
        JF_LAUNCH(ctx, "full_table", full_table_kernel<F><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
            (E *)pl->off_full, (const E *)pl->off_lo, (const E *)pl->off_hi, pl->split, (uint64_t)n));
        // the plan is shared by every lane (stream) of the context: complete the table before anyone else can see the pointer
        JF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    for (int i = 0; i < np; i++) {
        PassArgs a;
        a.src = hop[i];
        a.dst = hop[i + 1];
        a.stride = batch_stride;
        a.in_len = in_len;
        a.batch = (uint32_t)batch;
        a.log_n = log_n;
        a.log_ns = pl->passes[i].log_ns;
        a.first = i == 0;
        a.last = i == np - 1;
        a.coset_in = (!inverse && has_off) ? 1 : 0;
        a.scale_mode = inverse ? (has_off ? 2 : (pl->ninv_folded ? 0 : 1)) : 0;
        a.ptw = pl->ptw[i];
        a.off_full = (inverse || in_len > n / 4) ? pl->off_full : nullptr;
        a.split = pl->split;
        a.w_lo = pl->w_lo;
        a.w_hi = pl->w_hi;
        a.off_lo = pl->off_lo;
        a.off_hi = pl->off_hi;
        a.tw = pl->tw[pl->passes[i].k];
        a.n_inv = pl->n_inv;
        a.src_stride = batch_stride;
        a.src_group = 1;
        a.off_rows = 0;  // not a multi-coset batch
        a.fold_c = nullptr;
        JF_TRY(launch_pass_k<F>(ctx, pl->passes[i].k, a));
    }
    if (in_place && np == 1)
        JF_CUDA(ctx, cudaMemcpy2DAsync(d_data, batch_stride * sizeof(E), t1, batch_stride * sizeof(E), n * sizeof(E), batch,
                                       cudaMemcpyDeviceToDevice, ctx->stream));
    return JF_OK;
}

// Several cosets of the same size-n domain in one batch (the quotient domain of the prover taken coset by coset:
// the 8n-point coset g<w_8n> is the union of the eight cosets (g w_8n^r)<w_n>, and a polynomial of degree < 2n
// is evaluated on one of them by reducing it mod X^n - (g w_8n^r)^n and transforming n coefficients).
//   forward: polynomial p at d_src + p * src_stride (in_len <= 2n coefficients)
//            -> d_dst[(p * rows + r) * n + i] = poly(offsets[r] * w_n^i)
//   inverse: in place on d_dst: row (p * rows + r) holds values on coset r and becomes the n coefficients of
//            the interpolant of degree < n
// All rows share the twiddle tables of the plain size-n plan; each coset has its own table of offset powers.
template <class F>
static int ntt_run_cosets_t(jf_ctx *ctx, int field, const void *d_src, size_t src_stride, size_t in_len, void *d_dst, unsigned log_n,
                            int inverse, const uint64_t *offsets, int rows, size_t polys) {
    using E = Fp<F>;
    if (log_n > (unsigned)F::TWO_ADICITY) return fail(ctx, JF_ERR_DOMAIN_TOO_LARGE, "ntt: log_n exceeds the field's two-adicity");
    if (log_n > 30) return fail(ctx, JF_ERR_NOMEM, "ntt: log_n > 30 is not supported");
    if (log_n < (unsigned)NTT_MIN_K) return fail(ctx, JF_ERR_INVALID_ARG, "ntt_cosets: domain smaller than 8");
    if (rows < 1 || rows > 16 || !offsets) return fail(ctx, JF_ERR_INVALID_ARG, "ntt_cosets: 1..16 cosets");
    const size_t n = (size_t)1 << log_n;
    if (!inverse && in_len > 2 * n) return fail(ctx, JF_ERR_INVALID_ARG, "ntt_cosets: in_len > 2 n");
    if (polys == 0) return JF_OK;
    const size_t total = polys * (size_t)rows;
    if (total >= (1u << 30)) return fail(ctx, JF_ERR_INVALID_ARG, "ntt: batch too large");
    NttPlan *base;
    bool unused;
    JF_TRY(get_plan<F>(ctx, field, log_n, inverse, nullptr, &base, &unused));
    std::string key = "cosets/" + std::to_string(field) + "/" + std::to_string(log_n) + "/" + std::to_string(inverse ? 1 : 0) + "/";
    for (int i = 0; i < 4 * rows; i++) {
        char h[20];
        snprintf(h, sizeof h, "%016llx", (unsigned long long)offsets[i]);
        key += h;
    }
    NttPlan *pl;
    auto it = ctx->ntt_plans.find(key);
    if (it == ctx->ntt_plans.end()) {
        pl = new NttPlan();
        pl->field = field;
        pl->log_n = log_n;
        pl->inverse = inverse != 0;
        pl->has_offset = true;
        pl->rows = rows;
        const int brc = [&]() -> int {
            E ninv = E::one();
            if (inverse && !base->ninv_folded) {
                E half = E::inv(E::from_u32(2));
                for (unsigned i = 0; i < log_n; i++) ninv = E::mul(ninv, half);
            }
            JF_CUDA(ctx, cudaMalloc(&pl->off_full, sizeof(E) * n * rows));
            JF_CUDA(ctx, cudaMalloc(&pl->fold_c, sizeof(E) * rows));
            std::vector<E> c(rows);
            for (int r = 0; r < rows; r++) {
                E off;
                for (int i = 0; i < 4; i++) {
                    off.v[2 * i] = (uint32_t)offsets[4 * r + i];
                    off.v[2 * i + 1] = (uint32_t)(offsets[4 * r + i] >> 32);
                }
                c[r] = off;
                for (unsigned i = 0; i < log_n; i++) c[r] = E::sqr(c[r]);
                if (inverse) off = E::inv(off);
                JF_LAUNCH(ctx, "pow_table", pow_table_kernel<F><<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(
                    (E *)pl->off_full + (size_t)r * n, off, ninv, 1, (uint32_t)n));
            }
            JF_CUDA(ctx, cudaMemcpyAsync(pl->fold_c, c.data(), sizeof(E) * rows, cudaMemcpyHostToDevice, ctx->stream));
            JF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // c is a local
            return JF_OK;
        }();
        if (brc != JF_OK) {
            free_plan(pl);
            return brc;
        }
        plan_cache_insert(ctx, key, pl, base);  // owned by the cache from here on (freed with the context)
    } else {
        pl = plan_touch(ctx, it->second);
    }
    const int np = (int)base->passes.size();
    void *t1 = nullptr, *t2 = nullptr;
    std::vector<const void *> hop(np + 1);
    hop[np] = d_dst;
    if (!inverse) {
        if (np >= 2) JF_TRY(scratch(ctx, "ntt_c1", total * n * sizeof(E), &t1));
        hop[0] = d_src;
        for (int i = np - 1; i >= 1; i--) hop[i] = ((np - i) % 2 == 1) ? t1 : d_dst;
    } else {
        JF_TRY(scratch(ctx, "ntt_c1", total * n * sizeof(E), &t1));
        if (np % 2 == 1 && np >= 3) JF_TRY(scratch(ctx, "ntt_c2", total * n * sizeof(E), &t2));
        hop[0] = d_dst;
        if (np == 1) hop[1] = t1;
        for (int i = 1; i < np; i++) hop[i] = (i % 2 == 1) ? t1 : (np % 2 == 0 ? (const void *)d_dst : (const void *)t2);
    }
    for (int i = 0; i < np; i++) {
        PassArgs a;
        a.src = hop[i];
        a.dst = const_cast<void *>(hop[i + 1]);
        a.stride = n;
        a.in_len = inverse ? n : in_len;
        a.batch = (uint32_t)total;
        a.log_n = log_n;
        a.log_ns = base->passes[i].log_ns;
        a.first = i == 0;
        a.last = i == np - 1;
        a.coset_in = inverse ? 0 : 1;
        a.scale_mode = inverse ? 2 : 0;
        a.ptw = base->ptw[i];
        a.off_full = pl->off_full;
        a.split = base->split;
        a.w_lo = base->w_lo;
        a.w_hi = base->w_hi;
        a.off_lo = nullptr;
        a.off_hi = nullptr;
        a.tw = base->tw[base->passes[i].k];
        a.n_inv = base->n_inv;
        a.src_stride = inverse ? n : src_stride;
        a.src_group = inverse ? 1u : (uint32_t)rows;
        a.off_rows = (uint32_t)rows;
        a.fold_c = (!inverse && in_len > n) ? pl->fold_c : nullptr;
        JF_TRY(launch_pass_k<F>(ctx, base->passes[i].k, a));
    }
    if (inverse && np == 1)
        JF_CUDA(ctx, cudaMemcpyAsync(d_dst, t1, total * n * sizeof(E), cudaMemcpyDeviceToDevice, ctx->stream));
    return JF_OK;
}

}  // namespace jf
