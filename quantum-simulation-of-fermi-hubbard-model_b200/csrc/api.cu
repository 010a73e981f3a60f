// C-ABI of libfhsim.so: context, states, observables (K2), pool screening (K3), immediate-mode ops (K1).
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <new>

#include "common.cuh"
#include "sector_eval.cuh"

static thread_local char g_err[512] = "";

void fh_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *fh_last_error(void) { return g_err; }
extern "C" int fh_version(void) { return 100; }

static inline int popcnt(u64 v) { return __builtin_popcountll(v); }

u64 fh_next_uid() {
    static u64 next = 1;
    return __atomic_fetch_add(&next, 1, __ATOMIC_RELAXED);
}

// ----------------------------------------------------------------------------------------------
// context
// ----------------------------------------------------------------------------------------------
extern "C" int fh_ctx_create(int device, void *stream, fh_ctx **out) {
    FH_REQUIRE(out != nullptr, "fh_ctx_create: out is NULL");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        fh_set_error("fh_ctx_create: no CUDA device available (%s); libfhsim has no CPU fallback",
                     e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return FH_ECUDA;
    }
    FH_REQUIRE(device >= 0 && device < count, "fh_ctx_create: device %d out of range (have %d)", device, count);
    FH_CUDA(cudaSetDevice(device));
    fh_ctx *ctx = new (std::nothrow) fh_ctx();
    if (!ctx) return FH_ENOMEM;
    ctx->device = device;
    if (stream) {
        ctx->stream = reinterpret_cast<cudaStream_t>(stream);
    } else {
        FH_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    cudaDeviceProp prop;
    FH_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    FH_TRY(fh_kernels_init_device());
    FH_CUDA(cudaMalloc(&ctx->d_partials, sizeof(double) * 2 * FH_MAX_PARTIALS));
    FH_CUDA(cudaMalloc(&ctx->d_result, sizeof(double) * 64));
    FH_CUDA(cudaMalloc(&ctx->d_diag, fh_diag_scratch_bytes()));
    FH_CUDA(cudaMalloc(&ctx->d_counter, sizeof(unsigned int) * 4));
    FH_CUDA(cudaMemset(ctx->d_counter, 0, sizeof(unsigned int) * 4));
    FH_CUDA(cudaMallocHost(&ctx->h_result, sizeof(double) * 64));
    *out = ctx;
    return FH_OK;
}

extern "C" int fh_ctx_destroy(fh_ctx *ctx) {
    if (!ctx) return FH_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_partials);
    cudaFree(ctx->d_result);
    cudaFreeHost(ctx->h_result);
    if (ctx->d_flush) cudaFree(ctx->d_flush);
    for (auto &blk : ctx->scratch_cache) cudaFree(blk.second);
    ctx->scratch_cache.clear();
    if (ctx->d_diag) cudaFree(ctx->d_diag);
    if (ctx->d_counter) cudaFree(ctx->d_counter);
    if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
    if (ctx->ev_stop) cudaEventDestroy(ctx->ev_stop);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return FH_OK;
}

int fh_ctx_scratch_get(fh_ctx *ctx, size_t bytes, void **out) {
    for (size_t k = 0; k < ctx->scratch_cache.size(); ++k)
        if (ctx->scratch_cache[k].first == bytes) {
            *out = ctx->scratch_cache[k].second;
            ctx->scratch_cached_bytes -= bytes;
            ctx->scratch_cache.erase(ctx->scratch_cache.begin() + (long)k);
            return FH_OK;
        }
    cudaError_t e = cudaMalloc(out, bytes);
    if (e != cudaSuccess && !ctx->scratch_cache.empty()) {
        // out of memory with blocks of other sizes parked in the cache: release them and retry once
        cudaGetLastError();
        for (auto &blk : ctx->scratch_cache) cudaFree(blk.second);
        ctx->scratch_cache.clear();
        ctx->scratch_cached_bytes = 0;
        e = cudaMalloc(out, bytes);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        fh_set_error("scratch allocation of %llu bytes failed: %s", (unsigned long long)bytes, cudaGetErrorString(e));
        return FH_ECUDA;
    }
    return FH_OK;
}

void fh_ctx_scratch_put(fh_ctx *ctx, size_t bytes, void *ptr) {
    if (!ptr) return;
    if (ctx->scratch_cache.size() < FH_SCRATCH_CACHE_MAX_BLOCKS &&
        ctx->scratch_cached_bytes + bytes <= FH_SCRATCH_CACHE_MAX_BYTES) {
        ctx->scratch_cache.emplace_back(bytes, ptr);
        ctx->scratch_cached_bytes += bytes;
    } else {
        cudaFree(ptr);
    }
}

extern "C" int fh_ctx_sync(fh_ctx *ctx) {
    FH_REQUIRE(ctx, "fh_ctx_sync: ctx is NULL");
    FH_CUDA(cudaStreamSynchronize(ctx->stream));
    return FH_OK;
}

extern "C" int fh_ctx_info(fh_ctx *ctx, int *sm_count, size_t *free_bytes, size_t *total_bytes) {
    FH_REQUIRE(ctx, "fh_ctx_info: ctx is NULL");
    FH_CUDA(cudaSetDevice(ctx->device));
    if (sm_count) *sm_count = ctx->sm_count;
    size_t f = 0, t = 0;
    FH_CUDA(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return FH_OK;
}

extern "C" int fh_ctx_flush_l2(fh_ctx *ctx, size_t bytes) {
    FH_REQUIRE(ctx, "fh_ctx_flush_l2: ctx is NULL");
    FH_CUDA(cudaSetDevice(ctx->device));
    bytes = (bytes + 31) & ~(size_t)31;
    if (bytes > ctx->flush_bytes) {
        if (ctx->d_flush) cudaFree(ctx->d_flush);
        ctx->d_flush = nullptr;
        ctx->flush_bytes = 0;
        FH_CUDA(cudaMalloc(&ctx->d_flush, bytes));
        ctx->flush_bytes = bytes;
    }
    launch_flush(ctx->stream, ctx->d_flush, bytes);
    FH_CUDA(cudaGetLastError());
    return FH_OK;
}

extern "C" int fh_ctx_timer_start(fh_ctx *ctx) {
    FH_REQUIRE(ctx, "fh_ctx_timer_start: ctx is NULL");
    if (!ctx->ev_start) {
        FH_CUDA(cudaEventCreate(&ctx->ev_start));
        FH_CUDA(cudaEventCreate(&ctx->ev_stop));
    }
    FH_CUDA(cudaEventRecord(ctx->ev_start, ctx->stream));
    return FH_OK;
}

extern "C" int fh_ctx_timer_stop(fh_ctx *ctx, double *elapsed_ms) {
    FH_REQUIRE(ctx && elapsed_ms, "fh_ctx_timer_stop: NULL argument");
    FH_REQUIRE(ctx->ev_start, "fh_ctx_timer_stop: timer not started");
    FH_CUDA(cudaEventRecord(ctx->ev_stop, ctx->stream));
    FH_CUDA(cudaEventSynchronize(ctx->ev_stop));
    float ms = 0.f;
    FH_CUDA(cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_stop));
    *elapsed_ms = ms;
    return FH_OK;
}

// ----------------------------------------------------------------------------------------------
// states
// ----------------------------------------------------------------------------------------------
extern "C" int fh_state_create(fh_ctx *ctx, int n_qubits, fh_state **out) {
    FH_REQUIRE(ctx && out, "fh_state_create: NULL argument");
    FH_REQUIRE(n_qubits >= 1 && n_qubits <= 33, "fh_state_create: n_qubits=%d outside [1, 33]", n_qubits);
    FH_CUDA(cudaSetDevice(ctx->device));
    fh_state *st = new (std::nothrow) fh_state();
    if (!st) return FH_ENOMEM;
    st->ctx = ctx;
    st->n = n_qubits;
    st->dim = 1ull << n_qubits;
    st->owned = true;
    cudaError_t e = cudaMalloc(&st->d, st->dim * sizeof(double2));
    if (e != cudaSuccess) {
        fh_set_error("fh_state_create: cudaMalloc of %llu bytes failed: %s", st->dim * 16ull, cudaGetErrorString(e));
        cudaGetLastError();
        delete st;
        return FH_ENOMEM;
    }
    launch_set_basis(ctx->stream, st->d, st->dim, 0);
    *out = st;
    return FH_OK;
}

extern "C" int fh_state_wrap(fh_ctx *ctx, int n_qubits, void *device_ptr, fh_state **out) {
    FH_REQUIRE(ctx && out && device_ptr, "fh_state_wrap: NULL argument");
    FH_REQUIRE(n_qubits >= 1 && n_qubits <= 33, "fh_state_wrap: n_qubits=%d outside [1, 33]", n_qubits);
    FH_REQUIRE((reinterpret_cast<uintptr_t>(device_ptr) & 15) == 0, "fh_state_wrap: pointer must be 16-byte aligned");
    fh_state *st = new (std::nothrow) fh_state();
    if (!st) return FH_ENOMEM;
    st->ctx = ctx;
    st->n = n_qubits;
    st->dim = 1ull << n_qubits;
    st->owned = false;
    st->d = reinterpret_cast<double2 *>(device_ptr);
    *out = st;
    return FH_OK;
}

extern "C" int fh_state_destroy(fh_state *st) {
    if (!st) return FH_OK;
    if (st->owned) {
        cudaSetDevice(st->ctx->device);
        cudaStreamSynchronize(st->ctx->stream);
        cudaFree(st->d);
    }
    delete st;
    return FH_OK;
}

extern "C" int fh_state_set_basis(fh_state *st, uint64_t index) {
    FH_REQUIRE(st, "fh_state_set_basis: state is NULL");
    FH_REQUIRE(index < st->dim, "fh_state_set_basis: index %llu >= 2^%d", (u64)index, st->n);
    launch_set_basis(st->ctx->stream, st->d, st->dim, index);
    FH_CUDA(cudaGetLastError());
    return FH_OK;
}

extern "C" int fh_state_copy(fh_state *dst, const fh_state *src) {
    FH_REQUIRE(dst && src, "fh_state_copy: NULL argument");
    FH_REQUIRE(dst->n == src->n, "fh_state_copy: qubit counts differ (%d vs %d)", dst->n, src->n);
    FH_CUDA(cudaMemcpyAsync(dst->d, src->d, src->dim * sizeof(double2), cudaMemcpyDeviceToDevice, dst->ctx->stream));
    return FH_OK;
}

extern "C" int fh_state_to_host(const fh_state *st, double *out) {
    FH_REQUIRE(st && out, "fh_state_to_host: NULL argument");
    FH_CUDA(cudaMemcpyAsync(out, st->d, st->dim * sizeof(double2), cudaMemcpyDeviceToHost, st->ctx->stream));
    FH_CUDA(cudaStreamSynchronize(st->ctx->stream));
    return FH_OK;
}

extern "C" int fh_state_from_host(fh_state *st, const double *in) {
    FH_REQUIRE(st && in, "fh_state_from_host: NULL argument");
    FH_CUDA(cudaMemcpyAsync(st->d, in, st->dim * sizeof(double2), cudaMemcpyHostToDevice, st->ctx->stream));
    FH_CUDA(cudaStreamSynchronize(st->ctx->stream));
    return FH_OK;
}

extern "C" int fh_state_device_ptr(const fh_state *st, void **ptr) {
    FH_REQUIRE(st && ptr, "fh_state_device_ptr: NULL argument");
    *ptr = st->d;
    return FH_OK;
}

extern "C" int fh_state_inner(const fh_state *a, const fh_state *b, double *re, double *im) {
    FH_REQUIRE(a && b && re && im, "fh_state_inner: NULL argument");
    FH_REQUIRE(a->n == b->n, "fh_state_inner: qubit counts differ");
    fh_ctx *ctx = a->ctx;
    launch_inner(ctx->stream, ctx->sm_count, a->d, b->d, a->dim, ctx->d_partials, ctx->d_result);
    FH_CUDA(cudaGetLastError());
    FH_CUDA(cudaMemcpyAsync(ctx->h_result, ctx->d_result, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FH_CUDA(cudaStreamSynchronize(ctx->stream));
    *re = ctx->h_result[0];
    *im = ctx->h_result[1];
    return FH_OK;
}

extern "C" int fh_state_norm2(const fh_state *st, double *out) {
    double im = 0.0;
    return fh_state_inner(st, st, out, &im);
}

// ----------------------------------------------------------------------------------------------
// pair / diag descriptors from user masks
// ----------------------------------------------------------------------------------------------
int fh_fill_pair(PairOp *op, int n, u64 x, u64 fixmask, u64 fixval, u64 zeta, const double m[8]) {
    const u64 full = n >= 64 ? ~0ull : ((1ull << n) - 1ull);
    FH_REQUIRE(x != 0 && (x & ~full) == 0, "pair op: x-mask 0x%llx invalid for %d qubits", x, n);
    FH_REQUIRE((fixmask & ~full) == 0 && (zeta & ~full) == 0, "pair op: mask outside %d qubits", n);
    const u64 top = 1ull << (63 - __builtin_clzll(x));
    FH_REQUIRE(fixmask & top, "pair op: fixmask must contain the top bit of x");
    FH_REQUIRE((fixval & ~fixmask) == 0, "pair op: fixval has bits outside fixmask");
    FH_REQUIRE((fixval & top) == 0, "pair op: fixval must have the top bit of x clear");
    FH_REQUIRE(popcnt(fixmask) <= FH_MAX_FIX, "pair op: more than %d fixed bits", FH_MAX_FIX);
    memset(op, 0, sizeof(*op));
    op->x = x;
    op->fixmask = fixmask;
    op->fixval = fixval;
    op->zeta = zeta;
    if (m) memcpy(op->m, m, sizeof(double) * 8);
    op->param = -1;
    int k = 0;
    for (int b = 0; b < n; ++b)
        if (fixmask >> b & 1) op->pos[k++] = (unsigned char)b;
    op->npos = (unsigned char)k;
    return FH_OK;
}

extern "C" int fh_apply_pair(fh_state *st, uint64_t x, uint64_t fixmask, uint64_t fixval, uint64_t zeta,
                             const double m[8]) {
    FH_REQUIRE(st && m, "fh_apply_pair: NULL argument");
    PairOp op;
    FH_TRY(fh_fill_pair(&op, st->n, x, fixmask, fixval, zeta, m));
    fh_ctx *ctx = st->ctx;
    PairOp *d_op = nullptr;
    FH_CUDA(cudaMallocAsync(&d_op, sizeof(PairOp), ctx->stream));
    FH_CUDA(cudaMemcpyAsync(d_op, &op, sizeof(PairOp), cudaMemcpyHostToDevice, ctx->stream));
    launch_pair(ctx->stream, ctx->sm_count, st->d, d_op, st->n, op.npos, 0);
    FH_CUDA(cudaGetLastError());
    FH_CUDA(cudaFreeAsync(d_op, ctx->stream));
    FH_CUDA(cudaStreamSynchronize(ctx->stream));   // `op` is a pageable stack source
    return FH_OK;
}

extern "C" int fh_state_swap_bits(fh_state *dst, const fh_state *src, int n_pairs, const int32_t *a, const int32_t *b) {
    FH_REQUIRE(dst && src && (n_pairs == 0 || (a && b)), "fh_state_swap_bits: NULL argument");
    FH_REQUIRE(dst->n == src->n && dst->d != src->d, "fh_state_swap_bits: states must have equal size and distinct buffers");
    FH_REQUIRE(n_pairs >= 0 && n_pairs <= 8, "fh_state_swap_bits: at most 8 bit pairs");
    int aa[8], bb[8];
    u64 seen = 0;
    for (int k = 0; k < n_pairs; ++k) {
        FH_REQUIRE(a[k] >= 0 && a[k] < src->n && b[k] >= 0 && b[k] < src->n && a[k] != b[k],
                   "fh_state_swap_bits: bit positions out of range");
        FH_REQUIRE(!(seen >> a[k] & 1ull) && !(seen >> b[k] & 1ull), "fh_state_swap_bits: bit pairs must be disjoint");
        seen |= (1ull << a[k]) | (1ull << b[k]);
        aa[k] = a[k];
        bb[k] = b[k];
    }
    fh_ctx *ctx = src->ctx;
    FH_CUDA(cudaSetDevice(ctx->device));
    launch_swap_bits(ctx->stream, ctx->sm_count, src->d, dst->d, src->n, n_pairs, aa, bb);
    FH_CUDA(cudaGetLastError());
    return FH_OK;
}

extern "C" int fh_apply_diag(fh_state *st, int n_terms, const uint64_t *z, const double *angle) {
    FH_REQUIRE(st && (n_terms == 0 || (z && angle)), "fh_apply_diag: NULL argument");
    FH_REQUIRE(n_terms >= 0, "fh_apply_diag: negative term count");
    if (n_terms == 0) return FH_OK;
    std::vector<DiagTerm> terms(n_terms);
    for (int m = 0; m < n_terms; ++m) {
        terms[m].z = z[m];
        terms[m].angle = angle[m];
        terms[m].c = cos(angle[m]);
        terms[m].s = sin(angle[m]);
        terms[m].coef = 0.0;
    }
    fh_ctx *ctx = st->ctx;
    DiagTerm *d = nullptr;
    FH_CUDA(cudaMallocAsync(&d, sizeof(DiagTerm) * n_terms, ctx->stream));
    FH_CUDA(cudaMemcpyAsync(d, terms.data(), sizeof(DiagTerm) * n_terms, cudaMemcpyHostToDevice, ctx->stream));
    launch_diag(ctx->stream, ctx->sm_count, st->d, d, n_terms, st->n, 0, ctx->d_diag);
    FH_CUDA(cudaGetLastError());
    FH_CUDA(cudaFreeAsync(d, ctx->stream));
    FH_CUDA(cudaStreamSynchronize(ctx->stream));
    return FH_OK;
}

extern "C" int fh_apply_pauli_rot_batch(fh_state *st, int m, const uint64_t *x, const uint64_t *z,
                                        const double *half_angles) {
    FH_REQUIRE(st && (m == 0 || (x && z && half_angles)), "fh_apply_pauli_rot_batch: NULL argument");
    FH_REQUIRE(m >= 0, "fh_apply_pauli_rot_batch: negative count");
    if (m == 0) return FH_OK;
    fh_ctx *ctx = st->ctx;
    // descriptors for every string up front, one upload, then one launch per string (in order)
    std::vector<PairOp> pairs;
    std::vector<DiagTerm> diags;
    std::vector<int> order;   // >=0: pair index; <0: ~diag index
    for (int t = 0; t < m; ++t) {
        const double a = half_angles[t], c = cos(a), s = sin(a);
        if (x[t] == 0) {
            if (z[t] == 0) continue;   // identity string: global phase, dropped (reference adapt_vqe.py:92-93)
            DiagTerm d;
            d.z = z[t];
            d.angle = a;
            d.c = c;
            d.s = s;
            d.coef = 0.0;
            order.push_back(~(int)diags.size());
            diags.push_back(d);
            continue;
        }
        // w(i) = B (-1)^popcount(i&z),  B = i^k (-1)^popcount(x&z),  k = popcount(x&z)
        const int k = popcnt(x[t] & z[t]);
        static const double ipow[4][2] = {{1, 0}, {0, 1}, {-1, 0}, {0, -1}};
        double br = ipow[k & 3][0], bi = ipow[k & 3][1];
        if (k & 1) { br = -br; bi = -bi; }
        // exp(-i a P): m00 = m11 = c, m01 = -i s B, m10 = -i s conj(B)
        const double mm[8] = {c, 0, s * bi, -s * br, -s * bi, -s * br, c, 0};
        const u64 top = 1ull << (63 - __builtin_clzll(x[t]));
        PairOp op;
        FH_TRY(fh_fill_pair(&op, st->n, x[t], top, 0, z[t], mm));
        order.push_back((int)pairs.size());
        pairs.push_back(op);
    }
    PairOp *d_pairs = nullptr;
    DiagTerm *d_diags = nullptr;
    if (!pairs.empty()) {
        FH_CUDA(cudaMallocAsync(&d_pairs, sizeof(PairOp) * pairs.size(), ctx->stream));
        FH_CUDA(cudaMemcpyAsync(d_pairs, pairs.data(), sizeof(PairOp) * pairs.size(), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (!diags.empty()) {
        FH_CUDA(cudaMallocAsync(&d_diags, sizeof(DiagTerm) * diags.size(), ctx->stream));
        FH_CUDA(cudaMemcpyAsync(d_diags, diags.data(), sizeof(DiagTerm) * diags.size(), cudaMemcpyHostToDevice, ctx->stream));
    }
    for (int o : order) {
        if (o >= 0)
            launch_pair(ctx->stream, ctx->sm_count, st->d, d_pairs + o, st->n, 1, 0);
        else
            launch_diag(ctx->stream, ctx->sm_count, st->d, d_diags + (~o), 1, st->n, 0);
    }
    FH_CUDA(cudaGetLastError());
    if (d_pairs) FH_CUDA(cudaFreeAsync(d_pairs, ctx->stream));
    if (d_diags) FH_CUDA(cudaFreeAsync(d_diags, ctx->stream));
    FH_CUDA(cudaStreamSynchronize(ctx->stream));
    return FH_OK;
}

// ----------------------------------------------------------------------------------------------
// observables
// ----------------------------------------------------------------------------------------------
extern "C" int fh_table_upload(fh_ctx *ctx, int n_qubits, int n_terms, const uint64_t *x, const uint64_t *z,
                               const double *coeff_re, const double *coeff_im, fh_table **out) {
    FH_REQUIRE(ctx && out, "fh_table_upload: NULL argument");
    FH_REQUIRE(n_terms >= 0 && (n_terms == 0 || (x && z && coeff_re)), "fh_table_upload: NULL term arrays");
    FH_REQUIRE(n_qubits >= 1 && n_qubits <= 33, "fh_table_upload: n_qubits=%d outside [1, 33]", n_qubits);
    const u64 full = (1ull << n_qubits) - 1ull;
    fh_table *tab = new (std::nothrow) fh_table();
    if (!tab) return FH_ENOMEM;
    tab->ctx = ctx;
    tab->n = n_qubits;
    tab->n_terms = n_terms;
    tab->all_real = true;
    tab->d_groups = nullptr;
    tab->d_classes = nullptr;
    tab->d_vals = nullptr;
    // group by x-mask in first-seen order; term order inside a group = table order
    std::map<u64, int> group_of;
    std::vector<std::vector<TabTerm>> buckets;
    std::vector<u64> group_x;
    for (int t = 0; t < n_terms; ++t) {
        if ((x[t] | z[t]) & ~full) {
            delete tab;
            fh_set_error("fh_table_upload: term %d has bits outside %d qubits", t, n_qubits);
            return FH_EINVAL;
        }
        auto it = group_of.find(x[t]);
        int g;
        if (it == group_of.end()) {
            g = (int)buckets.size();
            group_of[x[t]] = g;
            buckets.emplace_back();
            group_x.push_back(x[t]);
        } else {
            g = it->second;
        }
        const int k = popcnt(x[t] & z[t]) & 3;
        const double cr = coeff_re[t], ci = coeff_im ? coeff_im[t] : 0.0;
        TabTerm tt;
        tt.z = z[t];
        switch (k) {   // (cr + i ci) * i^k
            case 0: tt.dr = cr;  tt.di = ci;  break;
            case 1: tt.dr = -ci; tt.di = cr;  break;
            case 2: tt.dr = -cr; tt.di = -ci; break;
            default: tt.dr = ci; tt.di = -cr; break;
        }
        if (tt.di != 0.0) tab->all_real = false;
        buckets[g].push_back(tt);
    }
    std::vector<TabTerm> diag_tab_terms;
    for (size_t g = 0; g < buckets.size(); ++g) {
        const u64 gx = group_x[g];
        const std::vector<TabTerm> &bt = buckets[g];
        TabGroup grp;
        memset(&grp, 0, sizeof(grp));
        grp.x = gx;
        grp.first_class = (int)tab->classes.size();
        for (int q = 0; q < 4; ++q) grp.pos[q] = 63;
        const int kx = popcnt(gx);
        if (kx >= 1 && kx <= 4) {
            // classes by zeta = z & ~x (first-seen order); per class a 2^kx table over the x-bit pattern of j
            grp.kbits = kx;
            int q = 0;
            for (int b = 0; b < n_qubits; ++b)
                if (gx >> b & 1ull) grp.pos[q++] = (unsigned char)b;
            std::vector<u64> zetas;
            for (const TabTerm &tt : bt) {
                const u64 zeta = tt.z & ~gx;
                bool seen = false;
                for (u64 v : zetas) seen = seen || v == zeta;
                if (!seen) zetas.push_back(zeta);
            }
            for (u64 zeta : zetas) {
                std::vector<double2> tbl((size_t)1 << kx, make_double2(0.0, 0.0));
                bool any = false;
                for (int pat = 0; pat < (1 << kx); ++pat) {
                    u64 dep = 0;
                    for (int b = 0; b < kx; ++b)
                        if (pat >> b & 1) dep |= 1ull << grp.pos[b];
                    double vr = 0.0, vi = 0.0;
                    for (const TabTerm &tt : bt) {
                        if ((tt.z & ~gx) != zeta) continue;
                        const double sg = (popcnt(dep & tt.z) & 1) ? -1.0 : 1.0;
                        vr += sg * tt.dr;
                        vi += sg * tt.di;
                    }
                    tbl[pat] = make_double2(vr, vi);
                    any = any || vr != 0.0 || vi != 0.0;
                }
                if (!any) continue;
                TabClass cl;
                cl.zeta = zeta;
                cl.vofs = (int)tab->vals.size();
                cl.pad = 0;
                tab->classes.push_back(cl);
                tab->vals.insert(tab->vals.end(), tbl.begin(), tbl.end());
            }
        } else {
            grp.kbits = 0;
            for (const TabTerm &tt : bt) {
                // diagonal terms whose z-mask lies inside one 12-bit chunk of the index go into three additive
                // factor tables (built below); only chunk-straddling ones stay per-term classes
                if (gx == 0 && ((tt.z & ~0xfffull) == 0 || (tt.z & ~(0xfffull << 12)) == 0 || (tt.z & 0xffffffull) == 0)) {
                    diag_tab_terms.push_back(tt);
                    continue;
                }
                TabClass cl;
                cl.zeta = tt.z;
                cl.vofs = (int)tab->vals.size();
                cl.pad = 0;
                tab->classes.push_back(cl);
                tab->vals.push_back(make_double2(tt.dr, tt.di));
            }
        }
        grp.n_class = (int)tab->classes.size() - grp.first_class;
        grp.live = 0;
        for (int c = grp.first_class; c < grp.first_class + grp.n_class; ++c)
            for (int pat = 0; pat < (1 << grp.kbits); ++pat) {
                const double2 v = tab->vals[tab->classes[c].vofs + pat];
                if (v.x != 0.0 || v.y != 0.0) grp.live |= 1u << pat;
            }
        for (int r = 0; r < 8; ++r) {                 // how bits 8, 9, 10 of the index move the x-bit pattern
            unsigned d = 0;
            for (int q = 0; q < grp.kbits; ++q)
                if ((grp.pos[q] == 8 && (r & 1)) || (grp.pos[q] == 9 && (r & 2)) || (grp.pos[q] == 10 && (r & 4)))
                    d |= 1u << q;
            grp.rpat |= d << (4 * r);
        }
        tab->groups.push_back(grp);
        tab->terms.insert(tab->terms.end(), bt.begin(), bt.end());
        if (gx == 0) tab->diag_terms = bt;
    }
    tab->n_groups = (int)tab->groups.size();
    FH_CUDA(cudaSetDevice(ctx->device));
    if (!diag_tab_terms.empty()) {
        std::vector<double2> dt(4096 + 4096 + 1024, make_double2(0.0, 0.0));
        for (const TabTerm &tt : diag_tab_terms) {
            const int chunk = (tt.z & ~0xfffull) == 0 ? 0 : ((tt.z & ~(0xfffull << 12)) == 0 ? 1 : 2);
            const int base = chunk == 0 ? 0 : (chunk == 1 ? 4096 : 8192), len = chunk == 2 ? 1024 : 4096;
            for (int v = 0; v < len; ++v) {
                const u64 gl = (u64)v << (12 * chunk);
                const double sg = (popcnt(gl & tt.z) & 1) ? -1.0 : 1.0;
                dt[base + v].x += sg * tt.dr;
                dt[base + v].y += sg * tt.di;
            }
        }
        FH_CUDA(cudaMalloc(&tab->d_diag, sizeof(double2) * dt.size()));
        FH_CUDA(cudaMemcpy(tab->d_diag, dt.data(), sizeof(double2) * dt.size(), cudaMemcpyHostToDevice));
    }
    if (tab->n_groups) {
        FH_CUDA(cudaMalloc(&tab->d_groups, sizeof(TabGroup) * tab->groups.size()));
        FH_CUDA(cudaMalloc(&tab->d_classes, sizeof(TabClass) * (tab->classes.size() + 1)));
        FH_CUDA(cudaMalloc(&tab->d_vals, sizeof(double2) * (tab->vals.size() + 1)));
        FH_CUDA(cudaMemcpyAsync(tab->d_groups, tab->groups.data(), sizeof(TabGroup) * tab->groups.size(),
                                cudaMemcpyHostToDevice, ctx->stream));
        if (!tab->classes.empty())
            FH_CUDA(cudaMemcpyAsync(tab->d_classes, tab->classes.data(), sizeof(TabClass) * tab->classes.size(),
                                    cudaMemcpyHostToDevice, ctx->stream));
        if (!tab->vals.empty())
            FH_CUDA(cudaMemcpyAsync(tab->d_vals, tab->vals.data(), sizeof(double2) * tab->vals.size(),
                                    cudaMemcpyHostToDevice, ctx->stream));
        FH_CUDA(cudaStreamSynchronize(ctx->stream));
        FH_TRY(fh_table_plan_tiles(tab));
    }
    *out = tab;
    return FH_OK;
}

extern "C" int fh_table_free(fh_table *tab) {
    if (!tab) return FH_OK;
    cudaSetDevice(tab->ctx->device);
    cudaStreamSynchronize(tab->ctx->stream);
    cudaFree(tab->d_groups);
    cudaFree(tab->d_classes);
    cudaFree(tab->d_vals);
    cudaFree(tab->d_diag);
    fh_table_tiles_free(tab->tiles);
    fh_sector_forget_table(tab->uid);
    fh_sector_forget_table_plan(tab->uid);
    delete tab;
    return FH_OK;
}

extern "C" int fh_table_info(const fh_table *tab, int *n_terms, int *n_groups) {
    FH_REQUIRE(tab, "fh_table_info: table is NULL");
    if (n_terms) *n_terms = tab->n_terms;
    if (n_groups) *n_groups = tab->n_groups;
    return FH_OK;
}

// enqueue K2 without synchronising; result lands in ctx->d_result[slot*2 .. slot*2+1]
int fh_enqueue_apply_table(const fh_table *tab, const double2 *in, double2 *out, int result_slot) {
    fh_ctx *ctx = tab->ctx;
    launch_apply_table(ctx->stream, ctx->sm_count, tab, in, out, out ? 1 : 0, ctx->d_partials,
                       ctx->d_result + 2 * result_slot);
    FH_CUDA(cudaGetLastError());
    return FH_OK;
}

extern "C" int fh_apply_table(const fh_table *tab, const fh_state *in, fh_state *out, double *e_re, double *e_im) {
    FH_REQUIRE(tab && in, "fh_apply_table: NULL argument");
    FH_REQUIRE(in->n == tab->n, "fh_apply_table: table is for %d qubits, state has %d", tab->n, in->n);
    FH_REQUIRE(!out || out->n == in->n, "fh_apply_table: output qubit count differs");
    FH_REQUIRE(!out || out->d != in->d, "fh_apply_table: in and out must be different buffers");
    fh_ctx *ctx = tab->ctx;
    FH_TRY(fh_enqueue_apply_table(tab, in->d, out ? out->d : nullptr, 0));
    FH_CUDA(cudaMemcpyAsync(ctx->h_result, ctx->d_result, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FH_CUDA(cudaStreamSynchronize(ctx->stream));
    if (e_re) *e_re = ctx->h_result[0];
    if (e_im) *e_im = ctx->h_result[1];
    return FH_OK;
}

extern "C" int fh_apply_table_accumulate(const fh_table *tab, const fh_state *in, fh_state *out, double *e_re,
                                         double *e_im) {
    FH_REQUIRE(tab && in && out, "fh_apply_table_accumulate: NULL argument");
    FH_REQUIRE(in->n == tab->n && out->n == in->n, "fh_apply_table_accumulate: qubit counts differ");
    FH_REQUIRE(out->d != in->d, "fh_apply_table_accumulate: in and out must be different buffers");
    fh_ctx *ctx = tab->ctx;
    launch_apply_table(ctx->stream, ctx->sm_count, tab, in->d, out->d, 2, ctx->d_partials, ctx->d_result);
    FH_CUDA(cudaGetLastError());
    FH_CUDA(cudaMemcpyAsync(ctx->h_result, ctx->d_result, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FH_CUDA(cudaStreamSynchronize(ctx->stream));
    if (e_re) *e_re = ctx->h_result[0];
    if (e_im) *e_im = ctx->h_result[1];
    return FH_OK;
}

// ----------------------------------------------------------------------------------------------
// pool
// ----------------------------------------------------------------------------------------------
// Greedy cover of the pool's x-masks by sets of T index bits (one pass each).  A pass is only worth a full read of
// psi and lambda (32 * 2^n B) if it serves enough entries (each costs 4 * 2^n B through k_pool): small tails stay
// on the per-entry kernel.
static void build_pool_passes(fh_pool *pool) {
    const int n = pool->n;
    // measured (tools/run_k3.py, profiles/): 12-bit tiles win 13-16 % once the states exceed L2 (n >= 22); below that
    // the per-entry kernel reading straight from L2 is faster, so tiles are off unless FHSIM_POOL_TILE_BITS asks
    int T = n >= 22 ? 12 : 0;
    if (const char *env = getenv("FHSIM_POOL_TILE_BITS")) T = atoi(env);
    pool->tile_bits = 0;
    pool->tile_grid = 0;
    const int n_entries = (int)pool->entries.size();
    std::vector<int> todo;
    if (T >= 6 && T <= 12 && n >= T + 4 && n <= 32)
        for (int e = 0; e < n_entries; ++e)
            if (popcnt(pool->entries[e].x) <= T) todo.push_back(e);
    std::vector<char> covered(n_entries, 0);
    const int min_pass = 24;      // measured: a pass costs about as much as 20-25 entries on k_pool32 (profiles/r01_s6_ncu_full_24q.md)
    while ((int)todo.size() >= min_pass) {
        u64 bits = 0;
        while (popcnt(bits) < T) {
            int best = -1;
            double best_score = -1.0;
            for (int b = 0; b < n; ++b) {
                if (bits >> b & 1ull) continue;
                const u64 nb = bits | (1ull << b);
                const int room = T - popcnt(nb);
                double score = 0.0;
                for (int e : todo) {
                    const u64 x = pool->entries[e].x;
                    if (popcnt(x & ~nb) <= room) score += (double)(1u << (2 * popcnt(x & nb)));
                }
                if (score > best_score) {
                    best_score = score;
                    best = b;
                }
            }
            bits |= 1ull << best;
        }
        std::vector<int> got;
        for (int e : todo) {
            const PoolEntry &pe = pool->entries[e];
            if ((pe.x & ~bits) == 0 && popcnt(pe.fixmask & bits) <= 4) got.push_back(e);
        }
        if ((int)got.size() < min_pass) break;
        int tb[16], nb = 0;
        for (int b = 0; b < n; ++b)
            if (bits >> b & 1ull) tb[nb++] = b;
        for (size_t off = 0; off < got.size(); off += FH_POOL_PASS_MAX_RECS) {
            PoolPass ps;
            memset(&ps, 0, sizeof(ps));
            ps.nbits = T;
            ps.first_rec = (int)pool->tile_recs.size();
            for (int q = 0; q < T; ++q) ps.bits[q] = (unsigned char)tb[q];
            const size_t end = off + FH_POOL_PASS_MAX_RECS < got.size() ? off + FH_POOL_PASS_MAX_RECS : got.size();
            for (size_t k = off; k < end; ++k) {
                const PoolEntry &pe = pool->entries[got[k]];
                PoolTileRec r;
                memset(&r, 0, sizeof(r));
                r.fixmask_out = (unsigned)(pe.fixmask & ~bits);
                r.fixval_out = (unsigned)(pe.fixval & ~bits);
                r.zeta = (unsigned)pe.zeta;
                r.entry = got[k];
                r.br = pe.br;
                r.bi = pe.bi;
                for (int q = 0; q < 4; ++q) r.lowmask[q] = 0xffffffffu;
                int nl = 0;
                for (int q = 0; q < T; ++q) {
                    const u64 g = 1ull << tb[q];
                    if (pe.x & g) r.xlocal |= 1u << q;
                    if (pe.fixmask & g) {
                        r.lowmask[nl++] = (1u << q) - 1u;
                        if (pe.fixval & g) r.lfixval |= 1u << q;
                    }
                }
                r.nlfix = nl;
                pool->tile_recs.push_back(r);
                covered[got[k]] = 1;
            }
            ps.nrec = (int)pool->tile_recs.size() - ps.first_rec;
            pool->passes.push_back(ps);
        }
        std::vector<int> left;
        for (int e : todo)
            if (!covered[e]) left.push_back(e);
        todo.swap(left);
    }
    if (!pool->passes.empty()) {
        pool->tile_bits = T;
        const u64 ntiles = 1ull << (n - T);
        pool->tile_grid = (int)(ntiles < 1024 ? ntiles : 1024);
        for (int e = 0; e < n_entries; ++e)
            if (!covered[e]) pool->rest.push_back(e);
    }
}

extern "C" int fh_pool_upload(fh_ctx *ctx, int n_qubits, int n_entries, const uint64_t *x, const uint64_t *fixmask,
                              const uint64_t *fixval, const uint64_t *zeta, const double *b_re, const double *b_im,
                              const int32_t *out_index, int n_out, fh_pool **out) {
    FH_REQUIRE(ctx && out, "fh_pool_upload: NULL argument");
    FH_REQUIRE(n_entries >= 0 && n_out >= 0, "fh_pool_upload: negative size");
    FH_REQUIRE(n_entries == 0 || (x && fixmask && fixval && zeta && b_re && b_im && out_index),
               "fh_pool_upload: NULL entry arrays");
    fh_pool *pool = new (std::nothrow) fh_pool();
    if (!pool) return FH_ENOMEM;
    pool->ctx = ctx;
    pool->n = n_qubits;
    pool->n_entries = n_entries;
    pool->n_out = n_out;
    pool->d_entries = nullptr;
    pool->d_out_first = nullptr;
    pool->d_partials = nullptr;
    pool->d_out = nullptr;
    pool->h_out = nullptr;
    int min_fix = 64;
    int prev_out = 0;
    pool->narrow = 1;
    pool->out_first.assign(n_out + 1, 0);
    for (int e = 0; e < n_entries; ++e) {
        PairOp tmp;
        int rc = fh_fill_pair(&tmp, n_qubits, x[e], fixmask[e], fixval[e], zeta[e], nullptr);
        if (rc == FH_OK && (out_index[e] < prev_out || out_index[e] >= n_out)) {
            fh_set_error("fh_pool_upload: out_index must be non-decreasing and < n_out (entry %d)", e);
            rc = FH_EINVAL;
        }
        if (rc != FH_OK) {
            delete pool;
            return rc;
        }
        prev_out = out_index[e];
        PoolEntry pe;
        memset(&pe, 0, sizeof(pe));
        pe.x = tmp.x;
        pe.fixmask = tmp.fixmask;
        pe.fixval = tmp.fixval;
        pe.zeta = tmp.zeta;
        pe.br = b_re[e];
        pe.bi = b_im[e];
        pe.out = out_index[e];
        pe.npos = tmp.npos;
        memcpy(pe.pos, tmp.pos, FH_MAX_FIX);
        pool->entries.push_back(pe);
        pool->out_first[out_index[e] + 1] = e + 1;
        if (tmp.npos < min_fix) min_fix = tmp.npos;
        if (tmp.npos > 4) pool->narrow = 0;
    }
    for (int o = 1; o <= n_out; ++o)
        if (pool->out_first[o] < pool->out_first[o - 1]) pool->out_first[o] = pool->out_first[o - 1];
    if (n_entries == 0) min_fix = 1;
    const u64 max_pairs = 1ull << (n_qubits - min_fix);
    u64 chunks = (max_pairs + 128 * 8 - 1) / (128 * 8);
    if (chunks < 1) chunks = 1;
    if (chunks > 1024) chunks = 1024;
    pool->kchunks = (int)chunks;
    build_pool_passes(pool);
    pool->chunks = pool->kchunks > pool->tile_grid ? pool->kchunks : pool->tile_grid;
    FH_CUDA(cudaSetDevice(ctx->device));
    if (n_entries) {
        FH_CUDA(cudaMalloc(&pool->d_entries, sizeof(PoolEntry) * n_entries));
        FH_CUDA(cudaMemcpyAsync(pool->d_entries, pool->entries.data(), sizeof(PoolEntry) * n_entries,
                                cudaMemcpyHostToDevice, ctx->stream));
        FH_CUDA(cudaMalloc(&pool->d_partials, sizeof(double) * (size_t)n_entries * pool->chunks));
        // slots a kernel never writes (k_pool: [kchunks, chunks), k_pool_tile: [tile_grid, chunks)) stay zero
        FH_CUDA(cudaMemsetAsync(pool->d_partials, 0, sizeof(double) * (size_t)n_entries * pool->chunks, ctx->stream));
    }
    if (!pool->passes.empty()) {
        FH_CUDA(cudaMalloc(&pool->d_passes, sizeof(PoolPass) * pool->passes.size()));
        FH_CUDA(cudaMemcpyAsync(pool->d_passes, pool->passes.data(), sizeof(PoolPass) * pool->passes.size(),
                                cudaMemcpyHostToDevice, ctx->stream));
        FH_CUDA(cudaMalloc(&pool->d_tile_recs, sizeof(PoolTileRec) * pool->tile_recs.size()));
        FH_CUDA(cudaMemcpyAsync(pool->d_tile_recs, pool->tile_recs.data(), sizeof(PoolTileRec) * pool->tile_recs.size(),
                                cudaMemcpyHostToDevice, ctx->stream));
    }
    if (!pool->rest.empty()) {
        FH_CUDA(cudaMalloc(&pool->d_rest, sizeof(int) * pool->rest.size()));
        FH_CUDA(cudaMemcpyAsync(pool->d_rest, pool->rest.data(), sizeof(int) * pool->rest.size(), cudaMemcpyHostToDevice,
                                ctx->stream));
    }
    FH_CUDA(cudaMalloc(&pool->d_out_first, sizeof(int) * (n_out + 1)));
    FH_CUDA(cudaMemcpyAsync(pool->d_out_first, pool->out_first.data(), sizeof(int) * (n_out + 1), cudaMemcpyHostToDevice,
                            ctx->stream));
    FH_CUDA(cudaMalloc(&pool->d_out, sizeof(double) * (n_out > 0 ? n_out : 1)));
    FH_CUDA(cudaMallocHost(&pool->h_out, sizeof(double) * (n_out > 0 ? n_out : 1)));
    FH_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = pool;
    return FH_OK;
}

extern "C" int fh_pool_free(fh_pool *pool) {
    if (!pool) return FH_OK;
    cudaSetDevice(pool->ctx->device);
    cudaStreamSynchronize(pool->ctx->stream);
    cudaFree(pool->d_entries);
    cudaFree(pool->d_passes);
    cudaFree(pool->d_tile_recs);
    cudaFree(pool->d_rest);
    cudaFree(pool->d_out_first);
    cudaFree(pool->d_partials);
    cudaFree(pool->d_out);
    cudaFreeHost(pool->h_out);
    fh_sector_forget_pool(pool->uid);
    fh_sector_forget_pool_plan(pool->uid);
    delete pool;
    return FH_OK;
}

// enqueue K3 for outputs [first, first+count) without synchronising; results in pool->d_out[first..]
int fh_enqueue_pool(const fh_pool *pool, const double2 *psi, const double2 *lam, int first, int count,
                    double *d_out_override) {
    fh_ctx *ctx = pool->ctx;
    const int e0 = pool->out_first[first], e1 = pool->out_first[first + count];
    if (pool->passes.empty()) {
        launch_pool(ctx->stream, pool->d_entries, e0, e1 - e0, pool->kchunks, pool->n, psi, lam, pool->d_partials, nullptr,
                    e0, e1, pool->chunks, pool->narrow);
    } else {
        launch_pool_tiles(ctx->stream, pool->d_passes, (int)pool->passes.size(), pool->d_tile_recs, pool->tile_bits,
                          pool->tile_grid, pool->chunks, pool->n, psi, lam, pool->d_partials, e0, e1);
        launch_pool(ctx->stream, pool->d_entries, 0, (int)pool->rest.size(), pool->kchunks, pool->n, psi, lam,
                    pool->d_partials, pool->d_rest, e0, e1, pool->chunks, pool->narrow);
    }
    launch_pool_finalize(ctx->stream, pool->d_partials, pool->d_out_first, pool->chunks, first, count,
                         d_out_override ? d_out_override : pool->d_out);
    FH_CUDA(cudaGetLastError());
    return FH_OK;
}

extern "C" int fh_pool_gradients(const fh_pool *pool, const fh_state *psi, const fh_state *lambda, int first, int count,
                                 double *out) {
    FH_REQUIRE(pool && psi && lambda, "fh_pool_gradients: NULL argument");
    FH_REQUIRE(psi->n == pool->n && lambda->n == pool->n, "fh_pool_gradients: qubit count mismatch");
    FH_REQUIRE(first >= 0 && count >= 0 && first + count <= pool->n_out, "fh_pool_gradients: range [%d, %d) outside pool of %d",
               first, first + count, pool->n_out);
    if (count == 0) return FH_OK;
    fh_ctx *ctx = pool->ctx;
    FH_TRY(fh_enqueue_pool(pool, psi->d, lambda->d, first, count, nullptr));
    if (!out) return FH_OK;     // enqueue only (results stay on the device); used to time the kernel alone
    FH_CUDA(cudaMemcpyAsync(pool->h_out, pool->d_out + first, sizeof(double) * count, cudaMemcpyDeviceToHost, ctx->stream));
    FH_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(out, pool->h_out, sizeof(double) * count);
    return FH_OK;
}

int fh_alloc_check(void *p, const char *what) {
    if (!p) {
        fh_set_error("allocation failed: %s", what);
        return FH_ENOMEM;
    }
    return FH_OK;
}
