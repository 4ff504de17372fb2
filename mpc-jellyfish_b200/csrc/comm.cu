// comm.cu -- multi-GPU forms of the MSM (and of batched NTTs), SURVEY.md §8e.
//
// `UnivariateKzgPCS::commit` (primitives/src/pcs/univariate_kzg/mod.rs:106-111) is a sum over (coefficient, key point)
// pairs, so it shards by point range with ONE exchange step: the per-GPU partial sums (one XYZZ point each).
//
//  jf_comm  : one process per GPU.  NCCL (loaded at run time from libnccl.so.2) bootstraps the group.  The exchange is either
//             `ncclAllGather` of the 128 / 192-byte partials, or -- transport 2 -- peer-memory mailboxes: each rank maps every
//             peer's mailbox through CUDA IPC and ONE kernel at the tail of the MSM stores the partial into all peers' HBM
//             over NVLink, raises a sequence flag and waits for the peers' flags.  No collective launch, no extra stream.
//  jf_group : one process, several GPUs (what a Rust prover process is).  One worker thread per device uploads that device's
//             scalar slice, runs the slice, and brings its partial back; the partials meet in host memory.  Batched NTTs are
//             dealt out by polynomial and need no exchange at all (`prover.rs:552-567` is a par_iter over polynomials).
//
// The G-1 final additions and the `into_affine` inversion run on the host in both forms: they are ~400 dependent field
// multiplications, ~20 us on a CPU core against ~0.2 ms for a lone GPU thread.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <condition_variable>
#include <functional>
#include <thread>
#include <unistd.h>
#include "common.cuh"
#include "ec.cuh"

namespace jf {

// ---- NCCL, bound at run time ------------------------------------------------------------------------------------
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};

static NcclApi *nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // JF_NCCL_LIB names a specific copy.  Otherwise RTLD_NOLOAD first: inside a process that already carries an NCCL
        // (e.g. torch's bundled one) use that very copy -- a process holds one library per SONAME.
        if (const char *path = getenv("JF_NCCL_LIB")) api.handle = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names)
            if (!api.handle) api.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        for (const char *n : names)
            if (!api.handle) api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (!api.handle) {
            api.err = std::string("cannot load libnccl.so.2: ") + dlerror();
            return;
        }
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))dlsym(api.handle, "ncclAllGather");
        api.Broadcast = (decltype(api.Broadcast))dlsym(api.handle, "ncclBroadcast");
        api.GroupStart = (decltype(api.GroupStart))dlsym(api.handle, "ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))dlsym(api.handle, "ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
        if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.GetErrorString || !api.Broadcast ||
            !api.GroupStart || !api.GroupEnd) {
            api.err = "libnccl.so.2 lacks an expected symbol";
            api.handle = nullptr;
        }
    });
    return api.handle ? &api : nullptr;
}

#define JF_NCCL(ctx, api, expr)                                                                                  \
    do {                                                                                                         \
        ncclResult_t r_ = (expr);                                                                                \
        if (r_ != ncclSuccess) return jf::fail(ctx, JF_ERR_COMM, std::string(#expr) + ": " + (api)->GetErrorString(r_)); \
    } while (0)

// ---- mailboxes ---------------------------------------------------------------------------------------------------
// Mailbox of rank r (in r's HBM): entry [parity][source rank] = 192 bytes of XYZZ data + the sequence number of the call that
// wrote it.  Call k uses parity k & 1: a rank can only be one call ahead of its slowest peer (its call k+1 cannot finish
// before every peer has published k+1, which each peer does after reading call k), so two parities never collide.
static constexpr int MBOX_ENTRY = 256, MBOX_FLAG = 192, MAX_RANKS = 16;
struct PeerPtrs {
    unsigned char *p[MAX_RANKS];
};

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// One CTA.  (1) store this rank's partial into every rank's mailbox (peer stores over NVLink; the own one is local),
// (2) release the sequence flag at every destination, (3) wait until all sources have released theirs here, (4) copy the
// nranks partials, in rank order, to a dense array.
__global__ void __launch_bounds__(256)
msm_exchange_kernel(const uint4 *my_part, int vec_per_part, PeerPtrs mbox, int rank, int nranks, unsigned long long seq,
                    uint4 *out_parts, int *err) {
    const int par = (int)(seq & 1), t = threadIdx.x;
    const size_t my_entry = (size_t)(par * nranks + rank) * MBOX_ENTRY;
    for (int i = t; i < nranks * vec_per_part; i += blockDim.x) {
        const int dst = i / vec_per_part, k = i % vec_per_part;
        reinterpret_cast<uint4 *>(mbox.p[dst] + my_entry)[k] = my_part[k];
    }
    __threadfence_system();
    __syncthreads();
    if (t < nranks) {
        unsigned long long *f = reinterpret_cast<unsigned long long *>(mbox.p[t] + my_entry + MBOX_FLAG);
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(seq) : "memory");
    }
    if (t < nranks) {
        const unsigned long long *f =
            reinterpret_cast<const unsigned long long *>(mbox.p[rank] + (size_t)(par * nranks + t) * MBOX_ENTRY + MBOX_FLAG);
        const unsigned long long t0 = global_ns();
        unsigned long long v;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
            if (v == seq) break;
            if (global_ns() - t0 > 2000000000ull) {  // a peer never arrived: report instead of hanging the GPU
                *err = JF_ERR_COMM;
                break;
            }
            __nanosleep(200);
        }
    }
    __syncthreads();
    for (int i = t; i < nranks * vec_per_part; i += blockDim.x) {
        const int src = i / vec_per_part, k = i % vec_per_part;
        const uint4 *e = reinterpret_cast<const uint4 *>(mbox.p[rank] + (size_t)(par * nranks + src) * MBOX_ENTRY);
        out_parts[i] = __ldcv(e + k);
    }
}

}  // namespace jf

struct jf_comm {
    jf_ctx *ctx = nullptr;
    int rank = 0, nranks = 1, transport = 1;
    ncclComm_t nccl = nullptr;
    unsigned char *mbox = nullptr;  // this rank's mailbox
    jf::PeerPtrs peers = {};        // every rank's mailbox as seen from this process
    unsigned long long seq = 0;
    void *d_part = nullptr;   // this rank's partial (XYZZ)
    void *d_parts = nullptr;  // nranks partials
    bool via_ipc[jf::MAX_RANKS] = {};  // peers.p[r] was opened with cudaIpcOpenMemHandle (a rank of the SAME process is mapped directly)
};

// ---- the one-process form ------------------------------------------------------------------------------------------
struct jf_group {
    std::vector<jf_ctx *> ctx;
    std::vector<int> dev;
    std::string err;
    // one worker per member: jobs are posted to all of them and awaited together
    struct Worker {
        std::thread th;
        std::mutex mu;
        std::condition_variable cv;
        std::function<int()> job;
        bool has_job = false, done = false, quit = false;
        int rc = 0;
    };
    std::vector<Worker *> workers;
    std::mutex call_mu;  // one group call at a time
};

struct jf_group_srs {
    int curve = 0;
    size_t n = 0;
    std::vector<jf_srs *> slice;
    std::vector<size_t> start;  // slice g covers points [start[g], start[g + 1])
};

namespace jf {

static void worker_main(jf_group::Worker *w) {
    std::unique_lock<std::mutex> lk(w->mu);
    for (;;) {
        w->cv.wait(lk, [&] { return w->has_job || w->quit; });
        if (w->quit) return;
        std::function<int()> job = std::move(w->job);
        w->has_job = false;
        lk.unlock();
        int rc = job();
        lk.lock();
        w->rc = rc;
        w->done = true;
        w->cv.notify_all();
    }
}

// run f(g) on every member's worker thread; returns the first non-zero status and records that member's error text
static int group_run(jf_group *grp, const std::function<int(int)> &f) {
    const int G = (int)grp->ctx.size();
    for (int g = 0; g < G; g++) {
        jf_group::Worker *w = grp->workers[g];
        std::lock_guard<std::mutex> lk(w->mu);
        w->job = [&f, g] { return f(g); };
        w->has_job = true;
        w->done = false;
        w->cv.notify_all();
    }
    int rc = JF_OK;
    for (int g = 0; g < G; g++) {
        jf_group::Worker *w = grp->workers[g];
        std::unique_lock<std::mutex> lk(w->mu);
        w->cv.wait(lk, [&] { return w->done; });
        if (w->rc != JF_OK && rc == JF_OK) {
            rc = w->rc;
            grp->err = "gpu " + std::to_string(grp->dev[g]) + ": " + jf_last_error(grp->ctx[g]);
        }
    }
    return rc;
}

// balanced contiguous split of [0, n): the first n % G members get one extra point
static void split_range(size_t n, int G, std::vector<size_t> &start) {
    start.resize(G + 1);
    const size_t base = n / G, rem = n % G;
    size_t s = 0;
    for (int g = 0; g < G; g++) {
        start[g] = s;
        s += base + ((size_t)g < rem ? 1 : 0);
    }
    start[G] = n;
}

// one member's slice of an MSM: host scalars in, XYZZ partial out (host)
static int msm_partial_to_host(jf_ctx *ctx, const jf_srs *srs, size_t off, const uint64_t *scalars, size_t n, int mont,
                               uint64_t *h_out_xyzz) {
    std::lock_guard<std::mutex> lock(ctx->mu);
    cudaSetDevice(ctx->device);
    const size_t pt = (size_t)srs->limbs64 * 32;
    void *d_sc, *d_res, *h_res;
    JF_TRY(scratch(ctx, "msm_scalars0", 32 * (n ? n : 1), &d_sc));
    JF_TRY(scratch(ctx, "msm_results", pt, &d_res));
    JF_TRY(pinned(ctx, pt, &h_res));
    if (n) JF_CUDA(ctx, cudaMemcpyAsync(d_sc, scalars, 32 * n, cudaMemcpyHostToDevice, ctx->stream));
    JF_TRY(msm_run(ctx, srs, off, d_sc, n, mont, d_res));
    JF_CUDA(ctx, cudaMemcpyAsync(h_res, d_res, pt, cudaMemcpyDeviceToHost, ctx->stream));
    JF_TRY(check_dev_err(ctx));  // synchronises
    memcpy(h_out_xyzz, h_res, pt);
    return JF_OK;
}

// every rank's `d_part` (pt bytes: one XYZZ point) -> d_out_parts (nranks x pt, rank order) on every rank; ctx->stream.
// Calls must be issued in the same order on every rank and on ONE stream (the mailbox parities rely on it).
int comm_exchange_from(jf_ctx *ctx, jf_comm *c, const void *d_part, size_t pt, void *d_out_parts) {
    if (c->transport == 2) {
        c->seq++;
        JF_LAUNCH(ctx, "msm_exchange", msm_exchange_kernel<<<1, 256, 0, ctx->stream>>>((const uint4 *)d_part, (int)(pt / 16), c->peers, c->rank,
                                                                          c->nranks, c->seq, (uint4 *)d_out_parts, ctx->d_err));
        return JF_OK;
    }
    NcclApi *api = nccl_api();
    if (!api) return fail(ctx, JF_ERR_COMM, "NCCL is not available");
    JF_NCCL(ctx, api, api->AllGather(d_part, d_out_parts, pt, ncclChar, c->nccl, ctx->stream));
    return JF_OK;  // NCCL's kernel, not ours: not counted in jf_ctx_launch_count
}
int comm_size(const jf_comm *c) { return c->nranks; }
int comm_rank(const jf_comm *c) { return c->rank; }
// Rows dealt out round-robin (row r lives on rank r mod nranks, as that rank's local row r / nranks) -> all `rows` rows, in order,
// on every rank: one grouped NCCL broadcast per row over NVLink (bulk data: the mailboxes only carry 192-byte partials).
int comm_bcast_rows(jf_ctx *ctx, jf_comm *c, const void *d_local_rows, void *d_all_rows, size_t row_bytes, int rows) {
    NcclApi *api = nccl_api();
    if (!api) return fail(ctx, JF_ERR_COMM, "NCCL is not available");
    JF_NCCL(ctx, api, api->GroupStart());
    for (int r = 0; r < rows; r++) {
        const int root = r % c->nranks;
        char *dst = (char *)d_all_rows + (size_t)r * row_bytes;
        const void *src = root == c->rank ? (const char *)d_local_rows + (size_t)(r / c->nranks) * row_bytes : dst;
        const ncclResult_t r_ = api->Broadcast(src, dst, row_bytes, ncclChar, root, c->nccl, ctx->stream);
        if (r_ != ncclSuccess) {
            api->GroupEnd();  // do not leave the library inside a group
            return fail(ctx, JF_ERR_COMM, std::string("ncclBroadcast: ") + api->GetErrorString(r_));
        }
    }
    JF_NCCL(ctx, api, api->GroupEnd());
    return JF_OK;
}
jf_ctx *comm_ctx(const jf_comm *c) { return c->ctx; }
static int comm_exchange(jf_ctx *ctx, jf_comm *c, size_t pt, void *d_out_parts) { return comm_exchange_from(ctx, c, c->d_part, pt, d_out_parts); }

}  // namespace jf

using namespace jf;

extern "C" {

int jf_comm_unique_id(uint8_t id[JF_COMM_ID_BYTES]) {
    static_assert(sizeof(ncclUniqueId) == JF_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    NcclApi *api = nccl_api();
    if (!api || !id) return JF_ERR_COMM;
    ncclUniqueId u;
    if (api->GetUniqueId(&u) != ncclSuccess) return JF_ERR_COMM;
    memcpy(id, &u, sizeof u);
    return JF_OK;
}

void jf_comm_destroy(jf_comm *c) {
    if (!c) return;
    jf_ctx *ctx = c->ctx;
    std::lock_guard<std::mutex> lock(ctx->mu);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int r = 0; r < c->nranks; r++)
        if (r != c->rank && c->peers.p[r] && c->via_ipc[r]) cudaIpcCloseMemHandle(c->peers.p[r]);
    NcclApi *api = nccl_api();
    if (api && c->nccl) api->CommDestroy(c->nccl);
    if (c->mbox) cudaFree(c->mbox);
    if (c->d_part) cudaFree(c->d_part);
    if (c->d_parts) cudaFree(c->d_parts);
    delete c;
}

int jf_comm_init(jf_ctx *ctx, int rank, int nranks, const uint8_t id[JF_COMM_ID_BYTES], int transport, jf_comm **out) {
    if (!ctx) return JF_ERR_INVALID_ARG;
    if (!out || !id || nranks < 1 || nranks > MAX_RANKS || rank < 0 || rank >= nranks || transport < 0 || transport > 2)
        return fail(ctx, JF_ERR_INVALID_ARG, "comm_init: bad argument (1..16 ranks; transport 0, 1 or 2)");
    *out = nullptr;
    NcclApi *api = nccl_api();
    if (!api) {
        nccl_api();
        return fail(ctx, JF_ERR_COMM, "comm_init: NCCL is not available");
    }
    if (const char *e = getenv("JF_COMM_TRANSPORT")) {
        if (!strcmp(e, "nccl")) transport = 1;
        else if (!strcmp(e, "p2p")) transport = 2;
    }
    jf_comm *c = new jf_comm();
    c->ctx = ctx;
    c->rank = rank;
    c->nranks = nranks;
    int rc = [&]() -> int {
        std::lock_guard<std::mutex> lock(ctx->mu);
        cudaSetDevice(ctx->device);
        ncclUniqueId u;
        memcpy(&u, id, sizeof u);
        JF_NCCL(ctx, api, api->CommInitRank(&c->nccl, nranks, u, rank));
        const size_t mbytes = (size_t)2 * nranks * MBOX_ENTRY;
        JF_CUDA(ctx, cudaMalloc((void **)&c->mbox, mbytes));
        JF_CUDA(ctx, cudaMemset(c->mbox, 0, mbytes));
        JF_CUDA(ctx, cudaMalloc(&c->d_part, MBOX_ENTRY));
        JF_CUDA(ctx, cudaMalloc(&c->d_parts, (size_t)nranks * MBOX_ENTRY));
        // exchange the mailbox handles (64 bytes each + this rank's verdict on opening them) through NCCL itself
        struct Rec {
            cudaIpcMemHandle_t h;
            int ok;
            int pid, device;  // a rank of the same process (one thread per GPU) is reached through its pointer, not through IPC
            int pad0;
            unsigned long long ptr;
            int pad[10];
        };
        static_assert(sizeof(Rec) == 128, "Rec");
        Rec mine = {};
        mine.ok = cudaIpcGetMemHandle(&mine.h, c->mbox) == cudaSuccess ? 1 : 0;
        cudaGetLastError();
        mine.pid = (int)getpid();
        mine.device = ctx->device;
        mine.ptr = (unsigned long long)(uintptr_t)c->mbox;
        Rec *d_all;
        std::vector<Rec> all(nranks);
        JF_CUDA(ctx, cudaMalloc((void **)&d_all, sizeof(Rec) * (nranks + 1)));
        auto gather = [&]() -> int {
            JF_CUDA(ctx, cudaMemcpyAsync(d_all + nranks, &mine, sizeof(Rec), cudaMemcpyHostToDevice, ctx->stream));
            JF_NCCL(ctx, api, api->AllGather(d_all + nranks, d_all, sizeof(Rec), ncclChar, c->nccl, ctx->stream));
            JF_CUDA(ctx, cudaMemcpyAsync(all.data(), d_all, sizeof(Rec) * nranks, cudaMemcpyDeviceToHost, ctx->stream));
            JF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            return JF_OK;
        };
        int grc = gather();
        bool p2p_ok = grc == JF_OK && transport != 1;
        for (int r = 0; r < nranks && p2p_ok; r++) p2p_ok = all[r].ok != 0;
        if (p2p_ok) {
            for (int r = 0; r < nranks; r++) {
                if (r == rank) {
                    c->peers.p[r] = c->mbox;
                    continue;
                }
                void *p = nullptr;
                if (all[r].pid == mine.pid) {
                    // same process: CUDA IPC cannot map it; the pointer is valid here once peer access is on
                    if (all[r].device != ctx->device) {
                        int can = 0;
                        cudaDeviceCanAccessPeer(&can, ctx->device, all[r].device);
                        cudaError_t e = can ? cudaDeviceEnablePeerAccess(all[r].device, 0) : cudaErrorInvalidDevice;
                        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                            cudaGetLastError();
                            p2p_ok = false;
                            break;
                        }
                        cudaGetLastError();
                    }
                    c->peers.p[r] = (unsigned char *)(uintptr_t)all[r].ptr;
                    continue;
                }
                if (cudaIpcOpenMemHandle(&p, all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                    cudaGetLastError();
                    p2p_ok = false;
                    break;
                }
                c->peers.p[r] = (unsigned char *)p;
                c->via_ipc[r] = true;
            }
        }
        // second round: p2p only if EVERY rank could map every mailbox
        if (grc == JF_OK) {
            mine.ok = p2p_ok ? 1 : 0;
            grc = gather();
            for (int r = 0; r < nranks && grc == JF_OK; r++) p2p_ok = p2p_ok && all[r].ok != 0;
        }
        cudaFree(d_all);
        JF_TRY(grc);
        if (transport == 2 && !p2p_ok) return fail(ctx, JF_ERR_COMM, "comm_init: peer-memory transport requested but a peer's mailbox cannot be mapped");
        c->transport = p2p_ok ? 2 : 1;
        if (!p2p_ok)
            for (int r = 0; r < nranks; r++) {
                if (r != rank && c->peers.p[r] && c->via_ipc[r]) cudaIpcCloseMemHandle(c->peers.p[r]);
                c->peers.p[r] = nullptr;
                c->via_ipc[r] = false;
            }
        return JF_OK;
    }();
    if (rc != JF_OK) {
        std::string keep = ctx->err;
        jf_comm_destroy(c);
        ctx->err = keep;
        return rc;
    }
    *out = c;
    return JF_OK;
}

int jf_comm_transport(const jf_comm *c) { return c ? c->transport : 0; }

int jf_msm_sharded_device(jf_ctx *ctx, jf_comm *c, const jf_srs *srs, size_t base_offset, const void *d_scalars, size_t n_local,
                          int mont, void *d_out_parts) {
    if (!ctx) return JF_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    cudaSetDevice(ctx->device);
    if (!c || c->ctx != ctx || !srs || !d_out_parts || (n_local && !d_scalars))
        return fail(ctx, JF_ERR_INVALID_ARG, "msm_sharded_device: null argument or a comm of another context");
    JF_TRY(msm_run(ctx, srs, base_offset, d_scalars, n_local, mont, c->d_part));
    return comm_exchange(ctx, c, (size_t)srs->limbs64 * 32, d_out_parts);
}

int jf_msm_sharded(jf_ctx *ctx, jf_comm *c, const jf_srs *srs, size_t base_offset, const uint64_t *scalars, size_t n_local,
                   int mont, uint64_t *out_xy, int *out_inf) {
    if (!ctx) return JF_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    cudaSetDevice(ctx->device);
    if (!c || c->ctx != ctx || !srs || !out_xy || !out_inf || (n_local && !scalars))
        return fail(ctx, JF_ERR_INVALID_ARG, "msm_sharded: null argument or a comm of another context");
    if (base_offset > srs->n) return fail(ctx, JF_ERR_INVALID_ARG, "msm: base_offset beyond the commit key");
    const size_t n = n_local < srs->n - base_offset ? n_local : srs->n - base_offset;
    const size_t pt = (size_t)srs->limbs64 * 32;
    void *d_sc, *h_res;
    JF_TRY(scratch(ctx, "msm_scalars0", 32 * (n ? n : 1), &d_sc));
    JF_TRY(pinned(ctx, pt * c->nranks, &h_res));
    const void *src = nullptr;
    if (n && zero_copy_enabled() && msm_reads_scalars_once(srs, n)) src = pinned_device_view(scalars);  // see jf_msm
    if (!src) {
        src = d_sc;
        if (n) JF_CUDA(ctx, cudaMemcpyAsync(d_sc, scalars, 32 * n, cudaMemcpyHostToDevice, ctx->stream));
    }
    JF_TRY(msm_run(ctx, srs, base_offset, src, n, mont, c->d_part));
    JF_TRY(comm_exchange(ctx, c, pt, c->d_parts));
    JF_CUDA(ctx, cudaMemcpyAsync(h_res, c->d_parts, pt * c->nranks, cudaMemcpyDeviceToHost, ctx->stream));
    JF_TRY(check_dev_err(ctx));  // synchronises
    return msm_finish_host(ctx, srs->curve, (const uint64_t *)h_res, c->nranks, out_xy, out_inf);
}

// ---- one process, several GPUs ---------------------------------------------------------------------------------------
int jf_group_create(const int *devices, int n_dev, jf_group **out) {
    if (!out || !devices || n_dev < 1 || n_dev > 64) return JF_ERR_INVALID_ARG;
    *out = nullptr;
    jf_group *g = new jf_group();
    for (int i = 0; i < n_dev; i++) {
        jf_ctx *c = nullptr;
        int rc = jf_ctx_create(devices[i], &c);
        if (rc != JF_OK) {
            for (jf_ctx *x : g->ctx) jf_ctx_destroy(x);
            delete g;
            return rc;
        }
        g->ctx.push_back(c);
        g->dev.push_back(devices[i]);
    }
    for (int i = 0; i < n_dev; i++) {
        jf_group::Worker *w = new jf_group::Worker();
        w->th = std::thread(worker_main, w);
        g->workers.push_back(w);
    }
    *out = g;
    return JF_OK;
}

void jf_group_destroy(jf_group *g) {
    if (!g) return;
    for (jf_group::Worker *w : g->workers) {
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->quit = true;
            w->cv.notify_all();
        }
        w->th.join();
        delete w;
    }
    for (jf_ctx *c : g->ctx) jf_ctx_destroy(c);
    delete g;
}

int jf_group_size(const jf_group *g) { return g ? (int)g->ctx.size() : 0; }
jf_ctx *jf_group_ctx(jf_group *g, int i) { return g && i >= 0 && i < (int)g->ctx.size() ? g->ctx[i] : nullptr; }
const char *jf_group_last_error(const jf_group *g) { return g ? g->err.c_str() : "null group"; }

void jf_group_srs_free(jf_group *g, jf_group_srs *s) {
    if (!s) return;
    for (size_t i = 0; i < s->slice.size(); i++)
        if (s->slice[i]) jf_srs_free(g && i < g->ctx.size() ? g->ctx[i] : nullptr, s->slice[i]);
    delete s;
}

int jf_group_srs_load(jf_group *g, int curve, const void *affine_pts, size_t n, size_t stride_bytes, long inf_flag_offset,
                      int window_bits, int precompute, jf_group_srs **out) {
    if (!g || !out || (n && !affine_pts)) return JF_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(g->call_mu);
    const int G = (int)g->ctx.size();
    jf_group_srs *s = new jf_group_srs();
    s->curve = curve;
    s->n = n;
    s->slice.assign(G, nullptr);
    split_range(n, G, s->start);
    int rc = group_run(g, [&](int i) {
        const char *base = (const char *)affine_pts + s->start[i] * stride_bytes;
        return jf_srs_load(g->ctx[i], curve, base, s->start[i + 1] - s->start[i], stride_bytes, inf_flag_offset, window_bits, precompute,
                           &s->slice[i]);
    });
    if (rc != JF_OK) {
        jf_group_srs_free(g, s);
        return rc;
    }
    *out = s;
    return JF_OK;
}

int jf_group_srs_generate_for_testing(jf_group *g, int curve, const uint64_t *beta, size_t n, int window_bits, int precompute,
                                      jf_group_srs **out) {
    if (!g || !out || !beta) return JF_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(g->call_mu);
    const int G = (int)g->ctx.size();
    jf_group_srs *s = new jf_group_srs();
    s->curve = curve;
    s->n = n;
    s->slice.assign(G, nullptr);
    split_range(n, G, s->start);
    int rc = group_run(g, [&](int i) {
        return jf_srs_generate_for_testing(g->ctx[i], curve, beta, s->start[i], s->start[i + 1] - s->start[i], window_bits, precompute,
                                           &s->slice[i]);
    });
    if (rc != JF_OK) {
        jf_group_srs_free(g, s);
        return rc;
    }
    *out = s;
    return JF_OK;
}

int jf_group_msm(jf_group *g, const jf_group_srs *s, size_t base_offset, const uint64_t *scalars, size_t n, int mont,
                 uint64_t *out_xy, int *out_inf) {
    if (!g || !s || !out_xy || !out_inf || (n && !scalars)) return JF_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(g->call_mu);
    const int G = (int)g->ctx.size();
    if ((int)s->slice.size() != G) {
        g->err = "group_msm: the key belongs to another group";
        return JF_ERR_INVALID_ARG;
    }
    if (base_offset > s->n) {
        g->err = "msm: base_offset beyond the commit key";
        return JF_ERR_INVALID_ARG;
    }
    const size_t end = base_offset + (n < s->n - base_offset ? n : s->n - base_offset);  // arkworks: min(len(bases), len(scalars))
    const int L = s->curve == JF_BLS12_381 ? 6 : 4;
    std::vector<uint64_t> parts((size_t)G * 4 * L, 0);
    int rc = group_run(g, [&](int i) {
        const size_t lo = std::max(base_offset, s->start[i]), hi = std::min(end, s->start[i + 1]);
        if (lo >= hi) return (int)JF_OK;  // nothing of the range on this GPU: its partial stays the identity (all zero)
        return msm_partial_to_host(g->ctx[i], s->slice[i], lo - s->start[i], scalars + 4 * (lo - base_offset), hi - lo, mont,
                                   parts.data() + (size_t)i * 4 * L);
    });
    if (rc != JF_OK) return rc;
    return msm_finish_host(nullptr, s->curve, parts.data(), G, out_xy, out_inf);
}

int jf_group_ntt(jf_group *g, int field, uint64_t *data, size_t in_len, unsigned log_n, int inverse, const uint64_t *coset_offset,
                 size_t batch, size_t batch_stride) {
    if (!g || !data) return JF_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(g->call_mu);
    const size_t G = g->ctx.size();
    if (batch > 1 && log_n < 63 && batch_stride < ((size_t)1 << log_n)) {
        g->err = "ntt: batch_stride < domain size";
        return JF_ERR_INVALID_ARG;
    }
    // vector b belongs to member b mod G: member i sees a batch of ceil((batch - i) / G) vectors, G * batch_stride apart
    return group_run(g, [&](int i) {
        if ((size_t)i >= batch) return (int)JF_OK;
        const size_t mine = (batch - i + G - 1) / G;
        return jf_ntt(g->ctx[i], field, data + 4 * batch_stride * i, in_len, log_n, inverse, coset_offset, mine, batch_stride * G);
    });
}

}  // extern "C"
