// Multi-GPU plumbing of the sharded-state path behind the C-ABI: one communicator per rank, NCCL over
// NVLink 5 / NVSwitch.  The reference is single-process, single-device (models/adapt_vqe.py:156), so nothing here replaces
// reference code; it carries the global<->local qubit swap of the 32-qubit 4x4 configuration (BASELINE cfg 5).
//
// NCCL is resolved at run time (dlopen of libnccl.so.2): the library keeps loading on boxes without NCCL, and inside a
// torch process the already-loaded NCCL (same SONAME) is the one that gets used, so there is exactly one NCCL in the
// process.  All collectives run on a dedicated communication stream; ordering against the context's compute stream is by
// events, which is what lets the exchange of a swap overlap the local bit-permutation that feeds it:
//
//   fh_comm_swap_exchange:  for every peer p (in the order rank^1, rank^2, ...):
//        compute stream:  permute the chunk of the slab that is destined to p        (k_swap_bits on a sub-range)
//        comm stream:     wait for that chunk, ncclSend it to p, ncclRecv p's chunk   (one group per peer)
//   so chunk k+1 is being permuted while chunk k is on the wire; the chunk that stays on this rank is a local copy.
#include <dlfcn.h>
#include <string.h>

#include <new>
#include <vector>

#include "common.cuh"

// ---- the slice of the NCCL API we use (types as in nccl.h 2.x; resolved with dlsym) --------------------------------
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;          // ncclSuccess == 0
enum { NCCL_FLOAT64 = 8, NCCL_SUM = 0 };       // ncclDataType_t ncclFloat64 = 8 (ncclDouble), ncclRedOp_t ncclSum = 0

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

static NcclApi *nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.ok ? &api : nullptr;
    tried = true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) return nullptr;
#define FH_NCCL_SYM(field, sym)                                                   \
    do {                                                                          \
        *reinterpret_cast<void **>(&api.field) = dlsym(api.handle, sym);          \
        if (!api.field) return nullptr;                                           \
    } while (0)
    FH_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
    FH_NCCL_SYM(CommInitRank, "ncclCommInitRank");
    FH_NCCL_SYM(CommDestroy, "ncclCommDestroy");
    FH_NCCL_SYM(GroupStart, "ncclGroupStart");
    FH_NCCL_SYM(GroupEnd, "ncclGroupEnd");
    FH_NCCL_SYM(Send, "ncclSend");
    FH_NCCL_SYM(Recv, "ncclRecv");
    FH_NCCL_SYM(AllReduce, "ncclAllReduce");
    FH_NCCL_SYM(AllGather, "ncclAllGather");
    FH_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef FH_NCCL_SYM
    api.ok = true;
    return &api;
}

#define FH_NCCL(call)                                                                                   \
    do {                                                                                                \
        ncclResult_t _r = (call);                                                                       \
        if (_r != 0) {                                                                                  \
            fh_set_error("%s failed: %s (%s:%d)", #call, api->GetErrorString(_r), __FILE__, __LINE__);  \
            return FH_ECUDA;                                                                            \
        }                                                                                               \
    } while (0)

struct fh_comm {
    fh_ctx *ctx = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    cudaStream_t stream = nullptr;             // communication stream
    cudaEvent_t ready[64];                     // chunk k of the source is ready (recorded on the compute stream)
    cudaEvent_t done = nullptr;                // the whole exchange has landed (recorded on the comm stream)
    double *d_scal = nullptr, *h_scal = nullptr;
    int scal_cap = 0;
    double last_ms = 0.0;                      // device time of the most recent exchange (comm stream)
    cudaEvent_t t0 = nullptr, t1 = nullptr;
};

void launch_swap_bits_range(cudaStream_t s, int sm, const double2 *src, double2 *dst, int n, int npairs, const int *a,
                            const int *b, u64 first, u64 count);

extern "C" int fh_comm_unique_id(unsigned char *out128) {
    FH_REQUIRE(out128, "fh_comm_unique_id: NULL argument");
    NcclApi *api = nccl_api();
    FH_REQUIRE(api, "fh_comm_unique_id: libnccl.so.2 could not be loaded");
    ncclUniqueId id;
    FH_NCCL(api->GetUniqueId(&id));
    memcpy(out128, id.internal, 128);
    return FH_OK;
}

extern "C" int fh_comm_init(fh_ctx *ctx, const unsigned char *id128, int rank, int world, fh_comm **out) {
    FH_REQUIRE(ctx && id128 && out, "fh_comm_init: NULL argument");
    FH_REQUIRE(world >= 1 && world <= 64 && (world & (world - 1)) == 0, "fh_comm_init: world size %d must be a power of two <= 64", world);
    FH_REQUIRE(rank >= 0 && rank < world, "fh_comm_init: rank %d outside [0, %d)", rank, world);
    NcclApi *api = nccl_api();
    FH_REQUIRE(api, "fh_comm_init: libnccl.so.2 could not be loaded");
    FH_CUDA(cudaSetDevice(ctx->device));
    fh_comm *c = new (std::nothrow) fh_comm();
    if (!c) return FH_ENOMEM;
    c->ctx = ctx;
    c->rank = rank;
    c->world = world;
    for (int k = 0; k < 64; ++k) c->ready[k] = nullptr;
    ncclUniqueId id;
    memcpy(id.internal, id128, 128);
    ncclResult_t r = api->CommInitRank(&c->comm, world, id, rank);
    if (r != 0) {
        fh_set_error("ncclCommInitRank failed: %s", api->GetErrorString(r));
        delete c;
        return FH_ECUDA;
    }
    FH_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (int k = 0; k < world; ++k) FH_CUDA(cudaEventCreateWithFlags(&c->ready[k], cudaEventDisableTiming));
    FH_CUDA(cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming));
    FH_CUDA(cudaEventCreate(&c->t0));
    FH_CUDA(cudaEventCreate(&c->t1));
    c->scal_cap = 4096;
    FH_CUDA(cudaMalloc(&c->d_scal, sizeof(double) * c->scal_cap * (size_t)(world + 1)));
    FH_CUDA(cudaMallocHost(&c->h_scal, sizeof(double) * c->scal_cap * (size_t)(world + 1)));
    *out = c;
    return FH_OK;
}

extern "C" int fh_comm_destroy(fh_comm *c) {
    if (!c) return FH_OK;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->ctx->stream);
    NcclApi *api = nccl_api();
    if (api && c->comm) api->CommDestroy(c->comm);
    for (int k = 0; k < 64; ++k)
        if (c->ready[k]) cudaEventDestroy(c->ready[k]);
    if (c->done) cudaEventDestroy(c->done);
    if (c->t0) cudaEventDestroy(c->t0);
    if (c->t1) cudaEventDestroy(c->t1);
    cudaFree(c->d_scal);
    cudaFreeHost(c->h_scal);
    cudaStreamDestroy(c->stream);
    delete c;
    return FH_OK;
}

extern "C" int fh_comm_info(const fh_comm *c, int *rank, int *world, double *last_exchange_ms) {
    FH_REQUIRE(c, "fh_comm_info: comm is NULL");
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    if (last_exchange_ms) *last_exchange_ms = c->last_ms;
    return FH_OK;
}

// dst chunk p <- chunk `rank` of rank p's src, for every p; optional local bit permutation of src first (staged in
// `scratch`, chunk by chunk, each chunk sent as soon as it is ready).  With n_pairs == 0 scratch may be NULL and src is
// sent as it is.  src, scratch and dst are three different slabs of the same size.  Returns after the exchange has been
// ENQUEUED; the context's compute stream waits for it (stream-ordered), the host does not.
extern "C" int fh_comm_swap_exchange(fh_comm *c, const fh_state *src, fh_state *scratch, fh_state *dst, int n_pairs,
                                     const int32_t *a, const int32_t *b) {
    FH_REQUIRE(c && src && dst, "fh_comm_swap_exchange: NULL argument");
    FH_REQUIRE(src->n == dst->n && src->d != dst->d, "fh_comm_swap_exchange: src and dst must be distinct slabs of equal size");
    FH_REQUIRE(n_pairs >= 0 && n_pairs <= 8, "fh_comm_swap_exchange: at most 8 bit pairs");
    FH_REQUIRE(n_pairs == 0 || (scratch && a && b && scratch->n == src->n && scratch->d != src->d && scratch->d != dst->d),
               "fh_comm_swap_exchange: a bit permutation needs a third slab as scratch");
    NcclApi *api = nccl_api();
    FH_REQUIRE(api, "fh_comm_swap_exchange: NCCL not available");
    fh_ctx *ctx = c->ctx;
    FH_CUDA(cudaSetDevice(ctx->device));
    const int world = c->world, rank = c->rank, n = src->n;
    int g = 0;
    while ((1 << g) < world) ++g;
    FH_REQUIRE(n > g, "fh_comm_swap_exchange: slab smaller than the world size");
    const u64 chunk = 1ull << (n - g);                 // amplitudes per peer
    int aa[8], bb[8];
    for (int k = 0; k < n_pairs; ++k) {
        FH_REQUIRE(a[k] >= 0 && a[k] < n && b[k] >= 0 && b[k] < n && a[k] != b[k], "fh_comm_swap_exchange: bad bit pair");
        aa[k] = a[k];
        bb[k] = b[k];
    }
    const double2 *send_base = n_pairs ? scratch->d : src->d;
    // the comm stream must not start before everything already enqueued on the compute stream (producers of src)
    FH_CUDA(cudaEventRecord(c->ready[rank], ctx->stream));
    FH_CUDA(cudaStreamWaitEvent(c->stream, c->ready[rank], 0));
    FH_CUDA(cudaEventRecord(c->t0, c->stream));
    // own chunk first (no wire): permuted straight into dst
    if (n_pairs)
        launch_swap_bits_range(ctx->stream, ctx->sm_count, src->d, dst->d, n, n_pairs, aa, bb, (u64)rank * chunk, chunk);
    else
        FH_CUDA(cudaMemcpyAsync(dst->d + (u64)rank * chunk, src->d + (u64)rank * chunk, chunk * sizeof(double2),
                                cudaMemcpyDeviceToDevice, ctx->stream));
    for (int s = 1; s < world; ++s) {
        const int peer = rank ^ s;
        if (n_pairs) {
            launch_swap_bits_range(ctx->stream, ctx->sm_count, src->d, scratch->d, n, n_pairs, aa, bb, (u64)peer * chunk, chunk);
            FH_CUDA(cudaEventRecord(c->ready[peer], ctx->stream));
            FH_CUDA(cudaStreamWaitEvent(c->stream, c->ready[peer], 0));
        }
        FH_NCCL(api->GroupStart());
        FH_NCCL(api->Send(send_base + (u64)peer * chunk, chunk * 2, NCCL_FLOAT64, peer, c->comm, c->stream));
        FH_NCCL(api->Recv(dst->d + (u64)peer * chunk, chunk * 2, NCCL_FLOAT64, peer, c->comm, c->stream));
        FH_NCCL(api->GroupEnd());
    }
    FH_CUDA(cudaEventRecord(c->t1, c->stream));
    FH_CUDA(cudaEventRecord(c->done, c->stream));
    FH_CUDA(cudaStreamWaitEvent(ctx->stream, c->done, 0));       // consumers on the compute stream see the new slab
    FH_CUDA(cudaGetLastError());
    return FH_OK;
}

// device time (ms) of the most recent exchange on the communication stream; synchronises that stream
extern "C" int fh_comm_last_exchange_ms(fh_comm *c, double *ms) {
    FH_REQUIRE(c && ms, "fh_comm_last_exchange_ms: NULL argument");
    FH_CUDA(cudaEventSynchronize(c->t1));
    float f = 0.f;
    FH_CUDA(cudaEventElapsedTime(&f, c->t0, c->t1));
    *ms = c->last_ms = f;
    return FH_OK;
}

// values[count] <- sum over ranks (host in/out); ordered after the compute stream
extern "C" int fh_comm_all_reduce_sum(fh_comm *c, double *values, int count) {
    FH_REQUIRE(c && (values || count == 0), "fh_comm_all_reduce_sum: NULL argument");
    FH_REQUIRE(count >= 0 && count <= c->scal_cap, "fh_comm_all_reduce_sum: at most %d values", c->scal_cap);
    if (count == 0) return FH_OK;
    NcclApi *api = nccl_api();
    FH_REQUIRE(api, "fh_comm_all_reduce_sum: NCCL not available");
    fh_ctx *ctx = c->ctx;
    FH_CUDA(cudaSetDevice(ctx->device));
    memcpy(c->h_scal, values, sizeof(double) * count);
    FH_CUDA(cudaMemcpyAsync(c->d_scal, c->h_scal, sizeof(double) * count, cudaMemcpyHostToDevice, ctx->stream));
    FH_NCCL(api->AllReduce(c->d_scal, c->d_scal, (size_t)count, NCCL_FLOAT64, NCCL_SUM, c->comm, ctx->stream));
    FH_CUDA(cudaMemcpyAsync(c->h_scal, c->d_scal, sizeof(double) * count, cudaMemcpyDeviceToHost, ctx->stream));
    FH_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(values, c->h_scal, sizeof(double) * count);
    return FH_OK;
}

// out[world * count] <- concatenation over ranks of in[count] (host buffers)
extern "C" int fh_comm_all_gather(fh_comm *c, const double *in, int count, double *out) {
    FH_REQUIRE(c && in && out, "fh_comm_all_gather: NULL argument");
    FH_REQUIRE(count >= 1 && count <= c->scal_cap, "fh_comm_all_gather: 1..%d values per rank", c->scal_cap);
    NcclApi *api = nccl_api();
    FH_REQUIRE(api, "fh_comm_all_gather: NCCL not available");
    fh_ctx *ctx = c->ctx;
    FH_CUDA(cudaSetDevice(ctx->device));
    double *d_in = c->d_scal, *d_out = c->d_scal + c->scal_cap, *h_out = c->h_scal + c->scal_cap;
    memcpy(c->h_scal, in, sizeof(double) * count);
    FH_CUDA(cudaMemcpyAsync(d_in, c->h_scal, sizeof(double) * count, cudaMemcpyHostToDevice, ctx->stream));
    FH_NCCL(api->AllGather(d_in, d_out, (size_t)count, NCCL_FLOAT64, c->comm, ctx->stream));
    FH_CUDA(cudaMemcpyAsync(h_out, d_out, sizeof(double) * count * c->world, cudaMemcpyDeviceToHost, ctx->stream));
    FH_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(out, h_out, sizeof(double) * count * c->world);
    return FH_OK;
}

// Device-side gather of one slice of pool gradients per rank, for a pool-sharded screening of ONE state: every rank has
// evaluated outputs [first_r, first_r + count_r) into its own fh_pool; this all-gathers `count` doubles per rank
// (ranks with fewer pad) from the device buffer d_in into host out[world * count] with one NCCL all-gather and one
// device-to-host copy.
extern "C" int fh_comm_barrier(fh_comm *c) {
    FH_REQUIRE(c, "fh_comm_barrier: comm is NULL");
    double v = 0.0;
    return fh_comm_all_reduce_sum(c, &v, 1);
}
