// api.cu -- the extern "C" boundary declared in include/jf_b200.h.
//
// Host buffers in, host buffers out; every call stages through device memory owned by the
// context.  No CPU fallback exists: if CUDA is unusable jf_ctx_create fails and nothing else
// can be called.
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "field.cuh"

using namespace jf;

namespace jf {

template <class F> __global__ void field_op_kernel(int op, const Fp<F> *a, const Fp<F> *b, Fp<F> *out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fp<F> x = a[i], y = b ? b[i] : Fp<F>::zero(), r;
    switch (op) {
    case 0: r = Fp<F>::mul(x, y); break;
    case 1: r = Fp<F>::add(x, y); break;
    case 2: r = Fp<F>::sub(x, y); break;
    case 3: r = Fp<F>::sqr(x); break;
    case 4: r = Fp<F>::inv(x); break;
    case 5: r = Fp<F>::to_mont(x); break;
    case 6: r = Fp<F>::from_mont(x); break;
    default: r = Fp<F>::neg(x); break;
    }
    out[i] = r;
}

template <class F> static int field_op_t(jf_ctx *ctx, int op, const void *a, const void *b, void *out, size_t n) {
    if (n == 0) return JF_OK;
    JF_LAUNCH(ctx, "field_op", field_op_kernel<F><<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(op, (const Fp<F> *)a, (const Fp<F> *)b,
                                                                             (Fp<F> *)out, n));
    return JF_OK;
}

int field_op(jf_ctx *ctx, int field, int op, const void *d_a, const void *d_b, void *d_out, size_t n) {
    if (op < 0 || op > 7) return fail(ctx, JF_ERR_INVALID_ARG, "field_op: unknown op");
    switch (field) {
    case JF_BN254_FR: return field_op_t<Bn254Fr>(ctx, op, d_a, d_b, d_out, n);
    case JF_BN254_FQ: return field_op_t<Bn254Fq>(ctx, op, d_a, d_b, d_out, n);
    case JF_BLS12_381_FR: return field_op_t<Bls12381Fr>(ctx, op, d_a, d_b, d_out, n);
    case JF_BLS12_381_FQ: return field_op_t<Bls12381Fq>(ctx, op, d_a, d_b, d_out, n);
    }
    return fail(ctx, JF_ERR_INVALID_ARG, "field_op: unknown field");
}

static int field_limbs64(int field) { return field == JF_BLS12_381_FQ ? 6 : 4; }
static int curve_limbs64(int curve) { return curve == JF_BLS12_381 ? 6 : 4; }

// surface a sticky device-side error (scalar out of range) after a synchronisation point
int check_dev_err(jf_ctx *ctx) {
    int h = 0;
    JF_CUDA(ctx, cudaMemcpyAsync(&h, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    JF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (h != 0) {
        cudaMemsetAsync(ctx->d_err, 0, sizeof(int), ctx->stream);
        if (h == JF_ERR_COMM) return fail(ctx, h, "msm_sharded: a peer did not deliver its partial sum within 2 s");
        return fail(ctx, h, "msm: a scalar is not below the group order (expected canonical BigInts / reduced field elements)");
    }
    return JF_OK;
}

// rows of `width` bytes between pitched buffers; cudaMemcpy2DAsync refuses pitches of 2 GiB and more (a batch of
// 2^24-element vectors dealt out over 8 GPUs is 4 GiB apart on the host), so those go row by row
int copy_rows(jf_ctx *ctx, void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows, cudaMemcpyKind kind,
              cudaStream_t st) {
    if (rows == 0 || width == 0) return JF_OK;
    if (dpitch < ((size_t)1 << 31) && spitch < ((size_t)1 << 31)) {
        JF_CUDA(ctx, cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, kind, st));
        return JF_OK;
    }
    for (size_t r = 0; r < rows; r++)
        JF_CUDA(ctx, cudaMemcpyAsync((char *)dst + r * dpitch, (const char *)src + r * spitch, width, kind, st));
    return JF_OK;
}

bool zero_copy_enabled() {
    static const bool on = [] {
        const char *e = getenv("JF_MSM_ZEROCOPY");
        return !(e && e[0] == '0');
    }();
    return on;
}
// device-visible alias of a page-locked host buffer, nullptr for pageable memory
const void *pinned_device_view(const void *host) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
    return at.devicePointer;
}


}  // namespace jf

#define JF_GUARD(ctx)                              \
    if (!(ctx)) return JF_ERR_INVALID_ARG;         \
    std::lock_guard<std::mutex> lock_((ctx)->mu);  \
    cudaSetDevice((ctx)->device)

extern "C" {

int jf_ctx_create(int device, jf_ctx **out) {
    if (!out) return JF_ERR_INVALID_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) return JF_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return JF_ERR_CUDA;
    jf_ctx *ctx = new jf_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete ctx;
        return JF_ERR_CUDA;
    }
    ctx->sm_count = prop.multiProcessorCount;
    // highest priority: the library's own side streams (lowest) only fill what this stream leaves idle
    int prio_least = 0, prio_greatest = 0;
    cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
    if (cudaStreamCreateWithPriority(&ctx->own_stream, cudaStreamNonBlocking, prio_greatest) != cudaSuccess ||
        cudaMalloc((void **)&ctx->d_err, 64) != cudaSuccess || cudaMemset(ctx->d_err, 0, 64) != cudaSuccess) {
        delete ctx;
        return JF_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return JF_OK;
}

void jf_ctx_destroy(jf_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ntt_free_plans(ctx);
    for (auto &kv : ctx->scratch) cudaFree(kv.second.ptr);
    for (auto &r : ctx->prof) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    for (auto &e : ctx->event_pool) cudaEventDestroy(e);
    for (auto &e : ctx->sync_events) cudaEventDestroy(e);
    if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
    if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    cudaFree(ctx->d_err);
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int jf_ctx_set_stream(jf_ctx *ctx, void *cuda_stream) {
    JF_GUARD(ctx);
    JF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return JF_OK;
}

int jf_ctx_sync(jf_ctx *ctx) {
    JF_GUARD(ctx);
    return check_dev_err(ctx);
}

const char *jf_last_error(const jf_ctx *ctx) {
    // the text is rewritten by whichever thread fails next: hand every caller its own copy, taken under the lock
    static thread_local std::string mine;
    if (!ctx) return "null context";
    {
        std::lock_guard<std::mutex> lock(const_cast<jf_ctx *>(ctx)->mu);
        mine = ctx->err;
    }
    return mine.c_str();
}

uint64_t jf_ctx_launch_count(const jf_ctx *ctx) { return ctx ? ctx->launches : 0; }

// ---- commit key -------------------------------------------------------------------------------
int jf_srs_load(jf_ctx *ctx, int curve, const void *affine_pts, size_t n, size_t stride_bytes, long inf_flag_offset,
                int window_bits, int precompute, jf_srs **out) {
    JF_GUARD(ctx);
    if (!out || (n && !affine_pts)) return fail(ctx, JF_ERR_INVALID_ARG, "srs_load: null argument");
    if (curve != JF_BN254 && curve != JF_BLS12_381) return fail(ctx, JF_ERR_INVALID_ARG, "srs_load: unknown curve");
    const size_t rec = (size_t)curve_limbs64(curve) * 16;  // x || y
    if (stride_bytes < rec) return fail(ctx, JF_ERR_INVALID_ARG, "srs_load: stride smaller than a point");
    if (inf_flag_offset >= 0 && (size_t)inf_flag_offset >= stride_bytes)
        return fail(ctx, JF_ERR_INVALID_ARG, "srs_load: infinity flag outside the record");
    void *d_base = nullptr;
    JF_TRY(scratch(ctx, "srs_stage", rec * (n ? n : 1), &d_base));
    if (n) {
        if (stride_bytes == rec && inf_flag_offset < 0) {
            JF_CUDA(ctx, cudaMemcpyAsync(d_base, affine_pts, rec * n, cudaMemcpyHostToDevice, ctx->stream));
        } else {
            // repack once on the host: ark-ec's `Affine { x, y, infinity }` is 72 / 104 bytes with padding
            std::vector<unsigned char> packed(rec * n);
            const unsigned char *src = (const unsigned char *)affine_pts;
            for (size_t i = 0; i < n; i++) {
                const unsigned char *r = src + i * stride_bytes;
                if (inf_flag_offset >= 0 && r[inf_flag_offset]) memset(&packed[i * rec], 0, rec);
                else memcpy(&packed[i * rec], r, rec);
            }
            JF_CUDA(ctx, cudaMemcpyAsync(d_base, packed.data(), rec * n, cudaMemcpyHostToDevice, ctx->stream));
            JF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
    }
    return srs_build(ctx, curve, d_base, n, window_bits, precompute, out);
}

int jf_srs_generate_for_testing(jf_ctx *ctx, int curve, const uint64_t *beta, size_t first_power, size_t n, int window_bits,
                                int precompute, jf_srs **out) {
    JF_GUARD(ctx);
    if (!out || !beta) return fail(ctx, JF_ERR_INVALID_ARG, "srs_generate: null argument");
    if (curve != JF_BN254 && curve != JF_BLS12_381) return fail(ctx, JF_ERR_INVALID_ARG, "srs_generate: unknown curve");
    if (n >= (1ull << 32)) return fail(ctx, JF_ERR_INVALID_ARG, "srs_generate: too many points");
    const size_t rec = (size_t)curve_limbs64(curve) * 16;
    void *d_base = nullptr;
    JF_TRY(scratch(ctx, "srs_stage", rec * (n ? n : 1), &d_base));
    JF_TRY(srs_generate(ctx, curve, beta, first_power, n, d_base));
    return srs_build(ctx, curve, d_base, n, window_bits, precompute, out);
}

int jf_srs_read(jf_ctx *ctx, const jf_srs *srs, size_t first, size_t count, uint64_t *out_xy) {
    JF_GUARD(ctx);
    if (!srs || !out_xy) return fail(ctx, JF_ERR_INVALID_ARG, "srs_read: null argument");
    if (first > srs->n || count > srs->n - first) return fail(ctx, JF_ERR_INVALID_ARG, "srs_read: range outside the key");
    const size_t rec = (size_t)srs->limbs64 * 16;
    JF_CUDA(ctx, cudaMemcpyAsync(out_xy, (const char *)srs->d_points + first * rec, count * rec, cudaMemcpyDeviceToHost,
                                 ctx->stream));
    JF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return JF_OK;
}

size_t jf_srs_len(const jf_srs *srs) { return srs ? srs->n : 0; }
int jf_srs_window_bits(const jf_srs *srs) { return srs ? srs->window_bits : 0; }

void jf_srs_free(jf_ctx *ctx, jf_srs *srs) {
    if (!srs) return;
    if (ctx) {
        std::lock_guard<std::mutex> lock(ctx->mu);
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        cudaFree(srs->d_points);
    } else {
        cudaFree(srs->d_points);
    }
    delete srs;
}

// ---- MSM ----------------------------------------------------------------------------------------
// `sum`: the batch is ONE MSM cut into consecutive parts; out_xy / out_inf receive the single sum of the parts' points
static int msm_batch_locked(jf_ctx *ctx, const jf_srs *srs, const uint64_t *const *scalars, const size_t *lens,
                            const size_t *base_offsets, size_t batch, int mont, uint64_t *out_xy, int *out_inf, bool sum = false) {
    if (!srs || !out_xy || !out_inf || (batch && (!scalars || !lens))) return fail(ctx, JF_ERR_INVALID_ARG, "msm: null argument");
    const int L = srs->limbs64;
    const size_t pt = (size_t)L * 4 * 8;  // XYZZ bytes
    size_t max_len = 0;
    for (size_t i = 0; i < batch; i++) {
        if (lens[i] && !scalars[i]) return fail(ctx, JF_ERR_INVALID_ARG, "msm: null scalar vector");
        max_len = lens[i] > max_len ? lens[i] : max_len;
    }
    if (batch == 0) return JF_OK;
    // A batch runs in groups of up to GROUP MSMs (msm_run_many): the scalars of vector i+1 cross PCIe on the copy stream
    // while vector i goes through its bulk phases (digits, sort, bucket accumulation), and the latency-bound bucket
    // reduction runs once per group over all its bucket sets.  One staging buffer per group member.
    constexpr int GROUP = 8;
    const int G = batch < (size_t)GROUP ? (int)batch : GROUP;
    void *d_sc[GROUP], *d_res, *h_res;
    for (int k = 0; k < G; k++) {
        char name[32];
        snprintf(name, sizeof name, "msm_scalars%d", k);
        JF_TRY(scratch(ctx, name, 32 * (max_len ? max_len : 1), &d_sc[k]));
    }
    JF_TRY(scratch(ctx, "msm_results", pt * batch, &d_res));
    JF_TRY(pinned(ctx, pt * batch, &h_res));
    const bool piped = batch > 1;
    cudaStream_t main_stream = ctx->stream;
    enum { EV_READY = 0, EV_DONE = GROUP, EV_START = 2 * GROUP, EV_COUNT = 2 * GROUP + 1 };
    if (piped) {
        if (!ctx->copy_in) {
            JF_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
            JF_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
        }
        while (ctx->sync_events.size() < EV_COUNT) {
            cudaEvent_t e;
            JF_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ctx->sync_events.push_back(e);
        }
        // the copy stream starts after whatever already reads the staging buffers on the compute stream
        JF_CUDA(ctx, cudaEventRecord(ctx->sync_events[EV_START], main_stream));
        JF_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_in, ctx->sync_events[EV_START], 0));
    }
    int rc = JF_OK;
    if (!piped) {
        const size_t off = base_offsets ? base_offsets[0] : 0;
        if (off > srs->n) return fail(ctx, JF_ERR_INVALID_ARG, "msm: base_offset beyond the commit key");
        const size_t n = lens[0] < srs->n - off ? lens[0] : srs->n - off;
        // Page-locked scalars (jf_host_alloc / cudaHostRegister) that the sort reads exactly once are read by the sort kernel
        // straight from host memory: the PCIe transfer and the sort become one step instead of two (JF_MSM_ZEROCOPY=0: copy).
        const void *src = d_sc[0];
        if (n && zero_copy_enabled() && msm_reads_scalars_once(srs, n) && (src = pinned_device_view(scalars[0])) != nullptr) {
        } else {
            src = d_sc[0];
            if (n) JF_CUDA(ctx, cudaMemcpyAsync(d_sc[0], scalars[0], 32 * n, cudaMemcpyHostToDevice, main_stream));
        }
        rc = msm_run(ctx, srs, off, src, n, mont, d_res);
    }
    for (size_t g0 = 0; piped && g0 < batch && rc == JF_OK; g0 += G) {
        const int cnt = batch - g0 < (size_t)G ? (int)(batch - g0) : G;
        MsmJob jobs[GROUP];
        cudaEvent_t *ev = ctx->sync_events.data();
        for (int k = 0; k < cnt && rc == JF_OK; k++) {
            const size_t i = g0 + k, off = base_offsets ? base_offsets[i] : 0;
            if (off > srs->n) {
                rc = fail(ctx, JF_ERR_INVALID_ARG, "msm: base_offset beyond the commit key");
                break;
            }
            const size_t n = lens[i] < srs->n - off ? lens[i] : srs->n - off;
            jobs[k] = MsmJob{off, d_sc[k], n, mont, (char *)d_res + pt * i, ev[EV_READY + k], ev[EV_DONE + k]};
        }
        // member k's upload is issued right before its kernels are enqueued: the kernels of the members before it are
        // then already running while the host stages a pageable source buffer
        struct Up { jf_ctx *ctx; const MsmJob *jobs; const uint64_t *const *src; bool wait_done; } up{ctx, jobs, scalars + g0, g0 > 0};
        auto prepare = [](void *user, int k) -> int {
            Up *u = (Up *)user;
            const MsmJob &j = u->jobs[k];
            cudaError_t ce = cudaSuccess;
            if (u->wait_done) ce = cudaStreamWaitEvent(u->ctx->copy_in, j.done, 0);  // the previous group's member k has read this buffer
            if (ce == cudaSuccess && j.n) ce = cudaMemcpyAsync(const_cast<void *>(j.d_scalars), u->src[k], 32 * j.n, cudaMemcpyHostToDevice, u->ctx->copy_in);
            if (ce == cudaSuccess) ce = cudaEventRecord(j.ready, u->ctx->copy_in);
            if (ce != cudaSuccess) return fail(u->ctx, JF_ERR_CUDA, std::string("msm_batch: ") + cudaGetErrorString(ce));
            return JF_OK;
        };
        if (rc == JF_OK) rc = msm_run_many(ctx, srs, jobs, cnt, prepare, &up);
    }
    JF_TRY(rc);
    JF_CUDA(ctx, cudaMemcpyAsync(h_res, d_res, pt * batch, cudaMemcpyDeviceToHost, ctx->stream));
    JF_TRY(check_dev_err(ctx));  // synchronises
    if (sum) return msm_finish_host(ctx, srs->curve, (const uint64_t *)h_res, batch, out_xy, out_inf);
    for (size_t i = 0; i < batch; i++)
        JF_TRY(msm_finish_host(ctx, srs->curve, (const uint64_t *)((const char *)h_res + pt * i), 1,
                               out_xy + (size_t)2 * L * i, out_inf + i));
    return JF_OK;
}

int jf_msm(jf_ctx *ctx, const jf_srs *srs, size_t base_offset, const uint64_t *scalars, size_t n, int scalars_in_montgomery,
           uint64_t *out_xy, int *out_infinity) {
    JF_GUARD(ctx);
    // A large MSM whose scalars must be copied first (the sort reads them more than once) runs as two halves over consecutive
    // base ranges: the second half crosses PCIe beside the kernels of the first, both share one bucket reduction, and the two
    // points are added on the host.  Measured end to end (tools/time_msm_split.py): 2^22 12.5 -> 12.0 ms, 2^24 48.2 -> 45.1 ms;
    // no gain at 2^20, where the sort reads page-locked scalars in place (JF_MSM_SPLIT=0: never split).
    static const bool split_enabled = [] { const char *e = getenv("JF_MSM_SPLIT"); return !(e && e[0] == '0'); }();
    if (split_enabled && srs && scalars && base_offset <= srs->n) {
        const size_t len = n < srs->n - base_offset ? n : srs->n - base_offset;
        if (len >= ((size_t)1 << 21) && !(zero_copy_enabled() && msm_reads_scalars_once(srs, len) && pinned_device_view(scalars))) {
            const size_t h = len / 2;
            const uint64_t *parts[2] = {scalars, scalars + 4 * h};
            const size_t lens[2] = {h, len - h}, offs[2] = {base_offset, base_offset + h};
            return msm_batch_locked(ctx, srs, parts, lens, offs, 2, scalars_in_montgomery, out_xy, out_infinity, true);
        }
    }
    return msm_batch_locked(ctx, srs, &scalars, &n, &base_offset, 1, scalars_in_montgomery, out_xy, out_infinity);
}

int jf_msm_batch(jf_ctx *ctx, const jf_srs *srs, const uint64_t *const *scalars, const size_t *lens,
                 const size_t *base_offsets, size_t batch, int scalars_in_montgomery, uint64_t *out_xy, int *out_infinity) {
    JF_GUARD(ctx);
    return msm_batch_locked(ctx, srs, scalars, lens, base_offsets, batch, scalars_in_montgomery, out_xy, out_infinity);
}

int jf_msm_device(jf_ctx *ctx, const jf_srs *srs, size_t base_offset, const void *d_scalars, size_t n,
                  int scalars_in_montgomery, void *d_out_xyzz) {
    JF_GUARD(ctx);
    if (!srs || !d_out_xyzz || (n && !d_scalars)) return fail(ctx, JF_ERR_INVALID_ARG, "msm_device: null argument");
    return msm_run(ctx, srs, base_offset, d_scalars, n, scalars_in_montgomery, d_out_xyzz);
}

int jf_msm_combine(jf_ctx *ctx, int curve, const uint64_t *xyzz_parts, size_t parts, uint64_t *out_xy, int *out_infinity) {
    /* pure host code: ctx may be NULL (it only receives the error text) */
    if (!xyzz_parts || !out_xy || !out_infinity) return fail(ctx, JF_ERR_INVALID_ARG, "msm_combine: null argument");
    return msm_finish_host(ctx, curve, xyzz_parts, parts, out_xy, out_infinity);
}

// ---- NTT ----------------------------------------------------------------------------------------
int jf_ntt_device(jf_ctx *ctx, int field, void *d_data, size_t in_len, unsigned log_n, int inverse,
                  const uint64_t *coset_offset, size_t batch, size_t batch_stride) {
    JF_GUARD(ctx);
    if (!d_data) return fail(ctx, JF_ERR_INVALID_ARG, "ntt: null data");
    return ntt_run(ctx, field, d_data, d_data, in_len, log_n, inverse, coset_offset, batch, batch_stride);
}

int jf_ntt(jf_ctx *ctx, int field, uint64_t *data, size_t in_len, unsigned log_n, int inverse, const uint64_t *coset_offset,
           size_t batch, size_t batch_stride) {
    JF_GUARD(ctx);
    if (!data) return fail(ctx, JF_ERR_INVALID_ARG, "ntt: null data");
    if (field != JF_BN254_FR && field != JF_BLS12_381_FR)
        return fail(ctx, JF_ERR_INVALID_ARG, "ntt: field must be BN254 Fr or BLS12-381 Fr");
    if (log_n > (field == JF_BN254_FR ? 28u : 32u))
        return fail(ctx, JF_ERR_DOMAIN_TOO_LARGE, "ntt: log_n exceeds the field's two-adicity");
    if (log_n > 30) return fail(ctx, JF_ERR_NOMEM, "ntt: log_n > 30 is not supported");
    const size_t n = (size_t)1 << log_n;
    if (batch == 0) return JF_OK;
    if (in_len > n) return fail(ctx, JF_ERR_INVALID_ARG, "ntt: in_len > domain size");
    if (batch > 1 && batch_stride < n) return fail(ctx, JF_ERR_INVALID_ARG, "ntt: batch_stride < domain size");
    // device staging is dense: vector b lives at b * n
    void *d_in = nullptr, *d_out = nullptr;
    JF_TRY(scratch(ctx, "ntt_in", 32 * n * batch, &d_in));
    JF_TRY(scratch(ctx, "ntt_out", 32 * n * batch, &d_out));
    // Large batches are pipelined: the upload of group g+1 and the download of group g-1 run on their own
    // streams beside the kernels of group g (PCIe is full duplex), so the call costs about one direction
    // of the transfer instead of both plus the transform.
    size_t groups = 1;
    if (batch > 1 && 32 * n * batch >= ((size_t)64 << 20)) {
        groups = batch < 16 ? batch : 16;
        while (batch % groups) groups--;
    }
    if (groups == 1) {
        // only the first in_len entries are read by the kernels: upload just those
        if (in_len)
            JF_TRY(copy_rows(ctx, d_in, 32 * n, data, 32 * batch_stride, 32 * in_len, batch, cudaMemcpyHostToDevice,
                                           ctx->stream));
        JF_TRY(ntt_run(ctx, field, d_in, d_out, in_len, log_n, inverse, coset_offset, batch, n));
        JF_TRY(copy_rows(ctx, data, 32 * batch_stride, d_out, 32 * n, 32 * n, batch, cudaMemcpyDeviceToHost, ctx->stream));
        JF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return JF_OK;
    }
    if (!ctx->copy_in) {
        JF_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
        JF_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
    }
    while (ctx->sync_events.size() < 2 * groups + 1) {
        cudaEvent_t e;
        JF_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->sync_events.push_back(e);
    }
    const size_t per = batch / groups;
    // the copy streams start after whatever is already queued on the compute stream
    JF_CUDA(ctx, cudaEventRecord(ctx->sync_events[2 * groups], ctx->stream));
    JF_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_in, ctx->sync_events[2 * groups], 0));
    JF_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_out, ctx->sync_events[2 * groups], 0));
    const int prc = [&]() -> int {
        for (size_t g = 0; g < groups; g++) {
            char *di = (char *)d_in + 32 * n * per * g, *d_o = (char *)d_out + 32 * n * per * g;
            uint64_t *h = data + 4 * batch_stride * per * g;
            if (in_len)
                JF_TRY(copy_rows(ctx, di, 32 * n, h, 32 * batch_stride, 32 * in_len, per, cudaMemcpyHostToDevice, ctx->copy_in));
            JF_CUDA(ctx, cudaEventRecord(ctx->sync_events[2 * g], ctx->copy_in));
            JF_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->sync_events[2 * g], 0));
            JF_TRY(ntt_run(ctx, field, di, d_o, in_len, log_n, inverse, coset_offset, per, n));
            JF_CUDA(ctx, cudaEventRecord(ctx->sync_events[2 * g + 1], ctx->stream));
            JF_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_out, ctx->sync_events[2 * g + 1], 0));
            JF_TRY(copy_rows(ctx, h, 32 * batch_stride, d_o, 32 * n, 32 * n, per, cudaMemcpyDeviceToHost, ctx->copy_out));
        }
        return JF_OK;
    }();
    if (prc != JF_OK) {  // copies may still be queued against the caller's buffer: drain all three streams before reporting
        cudaStreamSynchronize(ctx->copy_in);
        cudaStreamSynchronize(ctx->copy_out);
        cudaStreamSynchronize(ctx->stream);
        return prc;
    }
    JF_CUDA(ctx, cudaStreamSynchronize(ctx->copy_out));
    JF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return JF_OK;
}

int jf_ntt_cosets(jf_ctx *ctx, int field, const uint64_t *polys_in, size_t in_len, size_t in_stride, size_t polys,
                  unsigned log_n, int inverse, const uint64_t *offsets, int rows, uint64_t *out) {
    JF_GUARD(ctx);
    if (!out || !offsets || (!inverse && !polys_in)) return fail(ctx, JF_ERR_INVALID_ARG, "ntt_cosets: null argument");
    if (field != JF_BN254_FR && field != JF_BLS12_381_FR)
        return fail(ctx, JF_ERR_INVALID_ARG, "ntt: field must be BN254 Fr or BLS12-381 Fr");
    if (log_n > (field == JF_BN254_FR ? 28u : 32u))
        return fail(ctx, JF_ERR_DOMAIN_TOO_LARGE, "ntt: log_n exceeds the field's two-adicity");
    if (log_n > 30) return fail(ctx, JF_ERR_NOMEM, "ntt: log_n > 30 is not supported");
    if (rows < 1 || rows > 16) return fail(ctx, JF_ERR_INVALID_ARG, "ntt_cosets: 1..16 cosets");
    if (polys == 0) return JF_OK;
    const size_t n = (size_t)1 << log_n, total = polys * (size_t)rows;
    if (!inverse && (in_len > 2 * n || in_stride < in_len)) return fail(ctx, JF_ERR_INVALID_ARG, "ntt_cosets: in_len > 2 n or in_stride < in_len");
    void *d_in = nullptr, *d_out = nullptr;
    JF_TRY(scratch(ctx, "nttc_out", 32 * n * total, &d_out));
    if (!inverse) {
        const size_t dstride = in_len ? in_len : 1;
        JF_TRY(scratch(ctx, "nttc_in", 32 * dstride * polys, &d_in));
        if (in_len)
            JF_TRY(copy_rows(ctx, d_in, 32 * dstride, polys_in, 32 * in_stride, 32 * in_len, polys, cudaMemcpyHostToDevice,
                                           ctx->stream));
        JF_TRY(ntt_run_cosets(ctx, field, d_in, dstride, in_len, d_out, log_n, 0, offsets, rows, polys));
    } else {
        JF_CUDA(ctx, cudaMemcpyAsync(d_out, out, 32 * n * total, cudaMemcpyHostToDevice, ctx->stream));
        JF_TRY(ntt_run_cosets(ctx, field, d_out, n, n, d_out, log_n, 1, offsets, rows, polys));
    }
    JF_CUDA(ctx, cudaMemcpyAsync(out, d_out, 32 * n * total, cudaMemcpyDeviceToHost, ctx->stream));
    JF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return JF_OK;
}

// ---- device buffers -----------------------------------------------------------------------------
int jf_dev_alloc(jf_ctx *ctx, size_t bytes, void **out) {
    JF_GUARD(ctx);
    if (!out) return fail(ctx, JF_ERR_INVALID_ARG, "dev_alloc: null out");
    JF_CUDA(ctx, cudaMalloc(out, bytes ? bytes : 1));
    return JF_OK;
}
int jf_dev_free(jf_ctx *ctx, void *ptr) {
    JF_GUARD(ctx);
    JF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    JF_CUDA(ctx, cudaFree(ptr));
    return JF_OK;
}
int jf_dev_upload(jf_ctx *ctx, void *dst, const void *src, size_t bytes) {
    JF_GUARD(ctx);
    JF_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    JF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return JF_OK;
}
int jf_dev_download(jf_ctx *ctx, void *dst, const void *src, size_t bytes) {
    JF_GUARD(ctx);
    JF_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    JF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return JF_OK;
}
int jf_host_alloc(jf_ctx *ctx, size_t bytes, void **out) {
    JF_GUARD(ctx);
    if (!out) return fail(ctx, JF_ERR_INVALID_ARG, "host_alloc: null out");
    JF_CUDA(ctx, cudaMallocHost(out, bytes ? bytes : 1));
    return JF_OK;
}
int jf_host_free(jf_ctx *ctx, void *ptr) {
    JF_GUARD(ctx);
    JF_CUDA(ctx, cudaFreeHost(ptr));
    return JF_OK;
}

// ---- element-wise helpers -----------------------------------------------------------------------
int jf_field_op(jf_ctx *ctx, int field, int op, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n) {
    JF_GUARD(ctx);
    if (field < 0 || field > 3) return fail(ctx, JF_ERR_INVALID_ARG, "field_op: unknown field");
    if (!a || !out) return fail(ctx, JF_ERR_INVALID_ARG, "field_op: null argument");
    const size_t bytes = (size_t)field_limbs64(field) * 8 * n;
    void *da, *db, *dout;
    JF_TRY(scratch(ctx, "fop_a", bytes + 1, &da));
    JF_TRY(scratch(ctx, "fop_b", bytes + 1, &db));
    JF_TRY(scratch(ctx, "fop_o", bytes + 1, &dout));
    JF_CUDA(ctx, cudaMemcpyAsync(da, a, bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (b) JF_CUDA(ctx, cudaMemcpyAsync(db, b, bytes, cudaMemcpyHostToDevice, ctx->stream));
    JF_TRY(field_op(ctx, field, op, da, b ? db : nullptr, dout, n));
    JF_CUDA(ctx, cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    JF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return JF_OK;
}

int jf_fixed_base_mul(jf_ctx *ctx, int curve, const uint64_t *scalars, size_t n, uint64_t *out_xy) {
    JF_GUARD(ctx);
    if (curve != JF_BN254 && curve != JF_BLS12_381) return fail(ctx, JF_ERR_INVALID_ARG, "fixed_base_mul: unknown curve");
    if (n && (!scalars || !out_xy)) return fail(ctx, JF_ERR_INVALID_ARG, "fixed_base_mul: null argument");
    if (n == 0) return JF_OK;
    const size_t rec = (size_t)curve_limbs64(curve) * 16;
    void *ds, *dp;
    JF_TRY(scratch(ctx, "fbm_s", 32 * n, &ds));
    JF_TRY(scratch(ctx, "fbm_p", rec * n, &dp));
    JF_CUDA(ctx, cudaMemcpyAsync(ds, scalars, 32 * n, cudaMemcpyHostToDevice, ctx->stream));
    JF_TRY(fixed_base_mul(ctx, curve, ds, n, dp));
    JF_CUDA(ctx, cudaMemcpyAsync(out_xy, dp, rec * n, cudaMemcpyDeviceToHost, ctx->stream));
    JF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return JF_OK;
}

// ---- measurement hooks ----------------------------------------------------------------------------
int jf_profile_enable(jf_ctx *ctx, int on) {
    JF_GUARD(ctx);
    ctx->prof_on = on != 0;
    ctx->prof_dominant_only = on == 2;
    return JF_OK;
}

long jf_profile_collect(jf_ctx *ctx, char *buf, size_t cap) {
    JF_GUARD(ctx);
    if (!buf || cap == 0) return JF_ERR_INVALID_ARG;
    JF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::map<std::string, std::pair<long, double>> agg;
    for (auto &r : ctx->prof) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            auto &e = agg[r.name];
            e.first++;
            e.second += ms;
        }
        ctx->event_pool.push_back(r.a);
        ctx->event_pool.push_back(r.b);
    }
    ctx->prof.clear();
    std::string s;
    char line[160];
    for (auto &kv : agg) {
        snprintf(line, sizeof line, "%s %ld %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        s += line;
    }
    if (s.size() + 1 > cap) return fail(ctx, JF_ERR_INVALID_ARG, "profile_collect: buffer too small");
    memcpy(buf, s.c_str(), s.size() + 1);
    return (long)s.size();
}

int jf_microbench(jf_ctx *ctx, int kind, double *out_rate) {
    JF_GUARD(ctx);
    if (!out_rate || kind < 0 || kind > 1) return fail(ctx, JF_ERR_INVALID_ARG, "microbench: bad argument");
    return microbench(ctx, kind, out_rate);
}

}  // extern "C"
