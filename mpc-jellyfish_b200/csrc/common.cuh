// common.cuh -- context, error plumbing and workspace shared by the translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <mutex>
#include <string>
#include <vector>
#include <map>

#include "../../include/jf_b200.h"

namespace jf {

// A grow-only device scratch buffer.
struct DevBuf {
    void *ptr = nullptr;
    size_t cap = 0;
};

struct NttPlan;  // ntt.cu

}  // namespace jf

struct jf_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;  // lazily created: host<->device copies beside the kernels
    std::vector<cudaEvent_t> sync_events;                // event pool for the copy pipeline
    int lane = 0;  // 1 while work is being issued on a secondary stream: scratch buffers are kept apart per lane
    std::mutex mu;
    std::string err;
    uint64_t launches = 0;
    int *d_err = nullptr;  // device-side sticky error flag (scalar range), checked at sync points
    // named scratch buffers (grow-only)
    std::map<std::string, jf::DevBuf> scratch;
    void *pinned = nullptr;  // small pinned staging area for results
    size_t pinned_cap = 0;
    // optional per-kernel event timing (jf_profile_*): (name, start, stop) per launch
    bool prof_on = false;
    bool prof_dominant_only = false;  // jf_profile_enable(ctx, 2): only msm_accumulate / ntt_pass launches are bracketed
    bool prof_open = false;           // the current launch has a start event
    struct ProfRec { const char *name; cudaEvent_t a, b; };
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> event_pool;
    // cached NTT plans keyed by (field, log_n, inverse, coset offset limbs)
    std::map<std::string, jf::NttPlan *> ntt_plans;
    uint64_t ntt_plan_clock = 0;  // LRU clock of the plan cache (ntt_impl.cuh: plan_cache_insert)
};

struct jf_srs {
    int curve = 0;
    size_t n = 0;          // number of points
    int limbs64 = 4;       // u64 limbs per coordinate
    int window_bits = 16;  // c
    int windows = 16;      // W
    int tables = 1;        // T: 1 (no precompute) or W
    void *d_points = nullptr;  // tables * n affine points, table-major
    int skew = 0;  // scalars against this key are expected to pile onto few buckets (a Lagrange-basis key meets witness VALUES:
                   // constant columns, small integers): compact sort with warp-aggregated atomics from the start
};

namespace jf {

inline int fail(jf_ctx *ctx, int code, const std::string &msg) {
    if (ctx) ctx->err = msg;
    return code;
}

#define JF_CUDA(ctx, expr)                                                                         \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            return jf::fail(ctx, e_ == cudaErrorMemoryAllocation ? JF_ERR_NOMEM : JF_ERR_CUDA,     \
                            std::string(#expr) + ": " + cudaGetErrorString(e_));                   \
        }                                                                                          \
    } while (0)

#define JF_TRY(expr)                \
    do {                            \
        int rc_ = (expr);           \
        if (rc_ != JF_OK) return rc_; \
    } while (0)

// scratch buffer of at least `bytes`, named so that independent users do not alias
inline int scratch(jf_ctx *ctx, const char *name, size_t bytes, void **out) {
    DevBuf &b = ctx->scratch[ctx->lane ? std::string(name) + "#1" : std::string(name)];
    if (b.cap < bytes) {
        if (b.ptr) {
            JF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            JF_CUDA(ctx, cudaFree(b.ptr));
            b.ptr = nullptr;
            b.cap = 0;
        }
        size_t want = bytes + bytes / 8 + 256;
        JF_CUDA(ctx, cudaMalloc(&b.ptr, want));
        b.cap = want;
    }
    *out = b.ptr;
    return JF_OK;
}

inline int pinned(jf_ctx *ctx, size_t bytes, void **out) {
    if (ctx->pinned_cap < bytes) {
        if (ctx->pinned) cudaFreeHost(ctx->pinned);
        ctx->pinned = nullptr;
        ctx->pinned_cap = 0;
        JF_CUDA(ctx, cudaMallocHost(&ctx->pinned, bytes + 4096));
        ctx->pinned_cap = bytes + 4096;
    }
    *out = ctx->pinned;
    return JF_OK;
}

#define JF_LAUNCH_CHECK(ctx)                                                            \
    do {                                                                                \
        (ctx)->launches++;                                                              \
        cudaError_t e_ = cudaGetLastError();                                            \
        if (e_ != cudaSuccess)                                                          \
            return jf::fail(ctx, JF_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e_)); \
    } while (0)

inline cudaEvent_t prof_event(jf_ctx *ctx) {
    cudaEvent_t e;
    if (!ctx->event_pool.empty()) {
        e = ctx->event_pool.back();
        ctx->event_pool.pop_back();
        return e;
    }
    cudaEventCreate(&e);
    return e;
}
inline void prof_begin(jf_ctx *ctx, const char *name) {
    ctx->prof_open = false;
    if (!ctx->prof_on) return;
    if (ctx->prof_dominant_only && strcmp(name, "msm_accumulate") != 0 && strcmp(name, "ntt_pass") != 0) return;
    ctx->prof_open = true;
    jf_ctx::ProfRec r{name, prof_event(ctx), prof_event(ctx)};
    cudaEventRecord(r.a, ctx->stream);
    ctx->prof.push_back(r);
}
inline void prof_end(jf_ctx *ctx) {
    if (!ctx->prof_open) return;
    cudaEventRecord(ctx->prof.back().b, ctx->stream);
}

// launch a kernel: optional event bracket, launch counter, launch-error check
#define JF_LAUNCH(ctx, name, ...)   \
    do {                            \
        jf::prof_begin(ctx, name);  \
        __VA_ARGS__;                \
        jf::prof_end(ctx);          \
        JF_LAUNCH_CHECK(ctx);       \
    } while (0)

// entry points implemented per translation unit
// d_out == d_data: in place.  Otherwise the result lands in d_out and d_data is clobbered.
int ntt_run(jf_ctx *ctx, int field, void *d_data, void *d_out, size_t in_len, unsigned log_n, int inverse,
            const uint64_t *coset_offset, size_t batch, size_t batch_stride);
// `rows` cosets offsets[r] * <w_n> of `polys` vectors in one batch (ntt_impl.cuh: ntt_run_cosets_t)
int ntt_run_cosets(jf_ctx *ctx, int field, const void *d_src, size_t src_stride, size_t in_len, void *d_dst, unsigned log_n,
                   int inverse, const uint64_t *offsets, int rows, size_t polys);
void ntt_free_plans(jf_ctx *ctx);
int msm_run(jf_ctx *ctx, const jf_srs *srs, size_t base_offset, const void *d_scalars, size_t n, int mont,
            void *d_out_xyzz);
// A group of MSMs over the same commit key on ctx->stream: the bulk phases run one after the other, the bucket
// reduction (latency-bound: ~20 dependent levels whatever the number of bucket sets) once for all of them.
struct MsmJob {
    size_t base_offset;
    const void *d_scalars;
    size_t n;
    int mont;
    void *d_out_xyzz;
    cudaEvent_t ready = nullptr;  // optional: the scalars are complete once this event has fired (upload on another stream)
    cudaEvent_t done = nullptr;   // optional: recorded on ctx->stream once this member no longer reads its scalars
};
// prepare(user, i), when given, is called right before member i is enqueued (e.g. to issue the upload its `ready`
// event follows: a copy from pageable memory blocks the host, so it must not delay the kernels of members < i).
int msm_run_many(jf_ctx *ctx, const jf_srs *srs, const MsmJob *jobs, int count, int (*prepare)(void *user, int i) = nullptr,
                 void *user = nullptr);
// true when an MSM of n pairs over this key sorts in ONE pass over the scalars (then they may be read straight from pinned host memory)
bool msm_reads_scalars_once(const jf_srs *srs, size_t n);
bool zero_copy_enabled();                          // JF_MSM_ZEROCOPY != 0
const void *pinned_device_view(const void *host);  // device-visible alias of a page-locked host buffer, nullptr for pageable memory
int msm_finish_host(jf_ctx *ctx, int curve, const uint64_t *xyzz_parts, size_t parts, uint64_t *out_xy, int *out_inf);
// comm.cu: all-gather of one XYZZ partial per rank (mailboxes over peer memory, or ncclAllGather); see comm_exchange_from there
int comm_exchange_from(jf_ctx *ctx, jf_comm *c, const void *d_part, size_t pt, void *d_out_parts);
int comm_size(const jf_comm *c);
int comm_rank(const jf_comm *c);
int comm_bcast_rows(jf_ctx *ctx, jf_comm *c, const void *d_local_rows, void *d_all_rows, size_t row_bytes, int rows);
jf_ctx *comm_ctx(const jf_comm *c);
int srs_build(jf_ctx *ctx, int curve, const void *d_base_points /* n affine, device */, size_t n, int window_bits,
              int precompute, jf_srs **out);
int srs_generate(jf_ctx *ctx, int curve, const uint64_t *beta, size_t first_power, size_t n, void *d_out_points);
// lagrange.cu: [L_j(beta)] G, j < 2^log_n, from the monomial key (inverse DFT in the group); mask_points: + P_n - P_0, P_(n+1) - P_1
int srs_lagrange(jf_ctx *ctx, const jf_srs *mono, unsigned log_n, int mask_points, jf_srs **out);
int microbench(jf_ctx *ctx, int kind, double *out_rate);
int fixed_base_mul(jf_ctx *ctx, int curve, const void *d_scalars, size_t n, void *d_out_points);
int field_op(jf_ctx *ctx, int field, int op, const void *d_a, const void *d_b, void *d_out, size_t n);
// copies the sticky device-side error flag back (synchronises the stream) and turns it into a status
int check_dev_err(jf_ctx *ctx);
int copy_rows(jf_ctx *ctx, void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows, cudaMemcpyKind kind,
              cudaStream_t st);

}  // namespace jf
