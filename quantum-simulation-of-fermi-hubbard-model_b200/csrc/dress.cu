// iQCC Hamiltonian dressing on packed Pauli tables, on the device.
//
// Reference models/iqcc_hubbard.py:184-189 replaces H by  H + sin(tau)(-i/2)[H, P] + (1/2)(1 - cos tau)(P H P - H)
// = exp(i tau P/2) H exp(-i tau P/2) for every entangler P, by symbolic QubitOperator products; the term table grows
// geometrically (100 -> 11 588 terms in one epoch at 3x3).  On packed (x, z, coeff) tables this is integer work:
//   a term that commutes with P keeps its coefficient;
//   a term c_t P_t that anticommutes becomes cos(tau) c_t P_t  +  (-i sin(tau) c_t) P_t P,  with
//   P_t P = i^(k_t + k_P - k_3) (-1)^popcount(z_t & x_P) P_3,  P_3 = (x_t ^ x_P, z_t ^ z_P),  k = popcount(x & z).
// t -> t ^ P is a bijection on strings, so every output string receives at most its own (kept) term plus ONE generated
// term: merging is a hash lookup of the partner, not a sort, and the two-term sum is order independent -- the result is
// bit-identical to the host restatement (fhsim.tables.PauliTable.dressed).  Output order = surviving input terms in
// input order, then the new strings in the order of the terms that generated them (the first-seen order of the
// reference's dict-based +=); terms with |c| <= tol are dropped.
//
//   k_hash_build     open-addressing table (x, z) -> term index           (atomicCAS on the slot owner)
//   k_dress_mark     per term: new coefficient, "creates a new string" flag, "survives" flag
//   k_scan_*         exclusive prefix sums of the flags (three-pass block scan)
//   k_dress_emit     compaction into the output arrays
#include <math.h>
#include <string.h>

#include <new>

#include "common.cuh"

struct fh_ptable {
    u64 uid = fh_next_uid();
    fh_ctx *ctx = nullptr;
    int n = 0;
    int count = 0, cap = 0;
    u64 *d_x = nullptr, *d_z = nullptr;
    double2 *d_c = nullptr;
    // scratch of the dressing pass (grown on demand)
    u64 *d_x2 = nullptr, *d_z2 = nullptr;
    double2 *d_c2 = nullptr, *d_cnew = nullptr;
    int *d_hash = nullptr, *d_flag_new = nullptr, *d_flag_keep = nullptr, *d_pos_new = nullptr, *d_pos_keep = nullptr,
        *d_block = nullptr, *d_total = nullptr;
    int hash_cap = 0, scratch_cap = 0;
    int *h_total = nullptr;
};

#define SCAN_BLOCK 1024

__device__ __forceinline__ unsigned hash_key(u64 x, u64 z) {
    u64 h = x * 0x9e3779b97f4a7c15ull ^ (z + 0x7f4a7c159e3779b9ull) * 0xc2b2ae3d27d4eb4full;
    h ^= h >> 29;
    h *= 0xbf58476d1ce4e5b9ull;
    h ^= h >> 32;
    return (unsigned)h;
}

__global__ void __launch_bounds__(256) k_hash_build(const u64 *__restrict__ x, const u64 *__restrict__ z, int n,
                                                    int *__restrict__ table, unsigned mask) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    unsigned s = hash_key(x[t], z[t]) & mask;
    for (;;) {
        const int prev = atomicCAS(table + s, -1, t);
        if (prev == -1) return;
        s = (s + 1) & mask;         // the table holds distinct strings (canonical tables): no equality test needed on insert
    }
}

__device__ __forceinline__ int hash_find(const u64 *x, const u64 *z, const int *table, unsigned mask, u64 kx, u64 kz) {
    unsigned s = hash_key(kx, kz) & mask;
    for (;;) {
        const int t = table[s];
        if (t < 0) return -1;
        if (x[t] == kx && z[t] == kz) return t;
        s = (s + 1) & mask;
    }
}

// (-i sin(tau) c_u) * i^(k_u + k_P - k_3) * (-1)^popcount(z_u & x_P): the term generated FROM term u
__device__ __forceinline__ double2 generated(double2 cu, u64 xu, u64 zu, u64 xp, u64 zp, double sn) {
    const int k1 = __popcll(xu & zu), k2 = __popcll(xp & zp), k3 = __popcll((xu ^ xp) & (zu ^ zp));
    int e = (k1 + k2 - k3) & 3;
    if (__popcll(zu & xp) & 1) e = (e + 2) & 3;
    // -i sin * c = (sin * c.im, -sin * c.re): two roundings, as numpy's ((-1j * sn) * c)
    const double pr = sn * cu.y, pi = -(sn * cu.x);
    switch (e) {            // times i^e, exact
        case 0: return make_double2(pr, pi);
        case 1: return make_double2(-pi, pr);
        case 2: return make_double2(-pr, -pi);
        default: return make_double2(pi, -pr);
    }
}

__global__ void __launch_bounds__(256) k_dress_mark(const u64 *__restrict__ x, const u64 *__restrict__ z,
                                                    const double2 *__restrict__ c, int n, const int *__restrict__ table,
                                                    unsigned mask, u64 xp, u64 zp, double cs, double sn, double tol,
                                                    double2 *__restrict__ cnew, int *__restrict__ flag_new,
                                                    int *__restrict__ flag_keep) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const u64 xt = x[t], zt = z[t];
    const double2 ct = c[t];
    const bool anti = ((__popcll(xt & zp) + __popcll(zt & xp)) & 1) != 0;
    double2 out = ct;
    int creates = 0;
    if (anti) {
        out = make_double2(cs * ct.x, cs * ct.y);
        const int u = hash_find(x, z, table, mask, xt ^ xp, zt ^ zp);
        if (u >= 0) {
            const double2 g = generated(c[u], x[u], z[u], xp, zp, sn);
            out.x += g.x;
            out.y += g.y;
        } else {
            creates = 1;        // this term generates a string that is not in the table yet
        }
    }
    cnew[t] = out;
    flag_keep[t] = (hypot(out.x, out.y) > tol) ? 1 : 0;
    // the generated coefficient is tested against tol when it is emitted
    if (creates) {
        const double2 g = generated(ct, xt, zt, xp, zp, sn);
        creates = (hypot(g.x, g.y) > tol) ? 1 : 0;
    }
    flag_new[t] = creates;
}

// ---- exclusive scan of int flags: per-block scan, scan of block sums (one block), add back ----
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_blocks(const int *__restrict__ in, int n, int *__restrict__ out,
                                                            int *__restrict__ block_sums) {
    __shared__ int sh[SCAN_BLOCK];
    const int t = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    const int v = t < n ? in[t] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < SCAN_BLOCK; off <<= 1) {
        const int add = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
        __syncthreads();
        sh[threadIdx.x] += add;
        __syncthreads();
    }
    if (t < n) out[t] = sh[threadIdx.x] - v;
    if (threadIdx.x == SCAN_BLOCK - 1) block_sums[blockIdx.x] = sh[threadIdx.x];
}

__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_sums(int *__restrict__ block_sums, int nblocks, int *__restrict__ total) {
    __shared__ int sh[SCAN_BLOCK];
    int carry = 0;
    for (int base = 0; base < nblocks; base += SCAN_BLOCK) {
        const int t = base + threadIdx.x;
        const int v = t < nblocks ? block_sums[t] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < SCAN_BLOCK; off <<= 1) {
            const int add = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
            __syncthreads();
            sh[threadIdx.x] += add;
            __syncthreads();
        }
        if (t < nblocks) block_sums[t] = carry + sh[threadIdx.x] - v;
        const int chunk_total = sh[SCAN_BLOCK - 1];
        __syncthreads();
        carry += chunk_total;
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(256) k_dress_emit(const u64 *__restrict__ x, const u64 *__restrict__ z,
                                                    const double2 *__restrict__ c, const double2 *__restrict__ cnew, int n,
                                                    const int *__restrict__ flag_keep, const int *__restrict__ pos_keep,
                                                    const int *__restrict__ blk_keep, const int *__restrict__ flag_new,
                                                    const int *__restrict__ pos_new, const int *__restrict__ blk_new,
                                                    const int *__restrict__ totals, u64 xp, u64 zp, double sn,
                                                    u64 *__restrict__ xo, u64 *__restrict__ zo, double2 *__restrict__ co) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int kept_total = totals[0];
    if (flag_keep[t]) {
        const int p = pos_keep[t] + blk_keep[t / SCAN_BLOCK];
        xo[p] = x[t];
        zo[p] = z[t];
        co[p] = cnew[t];
    }
    if (flag_new[t]) {
        const int p = kept_total + pos_new[t] + blk_new[t / SCAN_BLOCK];
        xo[p] = x[t] ^ xp;
        zo[p] = z[t] ^ zp;
        co[p] = generated(c[t], x[t], z[t], xp, zp, sn);
    }
}

// ----------------------------------------------------------------------------------------------
static int ptable_reserve(fh_ptable *pt, int cap) {
    if (cap <= pt->cap) return FH_OK;
    int ncap = pt->cap ? pt->cap : 1024;
    while (ncap < cap) ncap *= 2;
    u64 *nx = nullptr, *nz = nullptr;
    double2 *nc = nullptr;
    FH_CUDA(cudaMalloc(&nx, sizeof(u64) * ncap));
    FH_CUDA(cudaMalloc(&nz, sizeof(u64) * ncap));
    FH_CUDA(cudaMalloc(&nc, sizeof(double2) * ncap));
    if (pt->count) {
        cudaStream_t s = pt->ctx->stream;
        FH_CUDA(cudaMemcpyAsync(nx, pt->d_x, sizeof(u64) * pt->count, cudaMemcpyDeviceToDevice, s));
        FH_CUDA(cudaMemcpyAsync(nz, pt->d_z, sizeof(u64) * pt->count, cudaMemcpyDeviceToDevice, s));
        FH_CUDA(cudaMemcpyAsync(nc, pt->d_c, sizeof(double2) * pt->count, cudaMemcpyDeviceToDevice, s));
        FH_CUDA(cudaStreamSynchronize(s));
    }
    cudaFree(pt->d_x);
    cudaFree(pt->d_z);
    cudaFree(pt->d_c);
    pt->d_x = nx;
    pt->d_z = nz;
    pt->d_c = nc;
    pt->cap = ncap;
    return FH_OK;
}

static int ptable_scratch(fh_ptable *pt, int n) {
    if (n > pt->scratch_cap) {
        int cap = pt->scratch_cap ? pt->scratch_cap : 1024;
        while (cap < n) cap *= 2;
        cudaFree(pt->d_cnew); cudaFree(pt->d_flag_new); cudaFree(pt->d_flag_keep); cudaFree(pt->d_pos_new);
        cudaFree(pt->d_pos_keep); cudaFree(pt->d_block); cudaFree(pt->d_x2); cudaFree(pt->d_z2); cudaFree(pt->d_c2);
        FH_CUDA(cudaMalloc(&pt->d_cnew, sizeof(double2) * cap));
        FH_CUDA(cudaMalloc(&pt->d_flag_new, sizeof(int) * cap));
        FH_CUDA(cudaMalloc(&pt->d_flag_keep, sizeof(int) * cap));
        FH_CUDA(cudaMalloc(&pt->d_pos_new, sizeof(int) * cap));
        FH_CUDA(cudaMalloc(&pt->d_pos_keep, sizeof(int) * cap));
        FH_CUDA(cudaMalloc(&pt->d_block, sizeof(int) * 2 * ((cap + SCAN_BLOCK - 1) / SCAN_BLOCK + 1)));
        FH_CUDA(cudaMalloc(&pt->d_x2, sizeof(u64) * 2 * cap));
        FH_CUDA(cudaMalloc(&pt->d_z2, sizeof(u64) * 2 * cap));
        FH_CUDA(cudaMalloc(&pt->d_c2, sizeof(double2) * 2 * cap));
        pt->scratch_cap = cap;
    }
    int hcap = 1024;
    while (hcap < 4 * n) hcap *= 2;
    if (hcap > pt->hash_cap) {
        cudaFree(pt->d_hash);
        FH_CUDA(cudaMalloc(&pt->d_hash, sizeof(int) * hcap));
        pt->hash_cap = hcap;
    }
    if (!pt->d_total) {
        FH_CUDA(cudaMalloc(&pt->d_total, sizeof(int) * 2));
        FH_CUDA(cudaMallocHost(&pt->h_total, sizeof(int) * 2));
    }
    return FH_OK;
}

extern "C" int fh_ptable_upload(fh_ctx *ctx, int n_qubits, int n_terms, const uint64_t *x, const uint64_t *z,
                                const double *coeff_re, const double *coeff_im, fh_ptable **out) {
    FH_REQUIRE(ctx && out, "fh_ptable_upload: NULL argument");
    FH_REQUIRE(n_qubits >= 1 && n_qubits <= 64, "fh_ptable_upload: n_qubits=%d outside [1, 64]", n_qubits);
    FH_REQUIRE(n_terms >= 0 && (n_terms == 0 || (x && z && coeff_re)), "fh_ptable_upload: NULL term arrays");
    FH_CUDA(cudaSetDevice(ctx->device));
    fh_ptable *pt = new (std::nothrow) fh_ptable();
    if (!pt) return FH_ENOMEM;
    pt->ctx = ctx;
    pt->n = n_qubits;
    int rc = ptable_reserve(pt, n_terms > 0 ? n_terms : 1);
    if (rc != FH_OK) {
        delete pt;
        return rc;
    }
    if (n_terms) {
        std::vector<double2> c((size_t)n_terms);
        for (int t = 0; t < n_terms; ++t) c[t] = make_double2(coeff_re[t], coeff_im ? coeff_im[t] : 0.0);
        FH_CUDA(cudaMemcpy(pt->d_x, x, sizeof(u64) * n_terms, cudaMemcpyHostToDevice));
        FH_CUDA(cudaMemcpy(pt->d_z, z, sizeof(u64) * n_terms, cudaMemcpyHostToDevice));
        FH_CUDA(cudaMemcpy(pt->d_c, c.data(), sizeof(double2) * n_terms, cudaMemcpyHostToDevice));
    }
    pt->count = n_terms;
    *out = pt;
    return FH_OK;
}

extern "C" int fh_ptable_free(fh_ptable *pt) {
    if (!pt) return FH_OK;
    cudaSetDevice(pt->ctx->device);
    cudaStreamSynchronize(pt->ctx->stream);
    cudaFree(pt->d_x); cudaFree(pt->d_z); cudaFree(pt->d_c);
    cudaFree(pt->d_x2); cudaFree(pt->d_z2); cudaFree(pt->d_c2); cudaFree(pt->d_cnew);
    cudaFree(pt->d_hash); cudaFree(pt->d_flag_new); cudaFree(pt->d_flag_keep); cudaFree(pt->d_pos_new);
    cudaFree(pt->d_pos_keep); cudaFree(pt->d_block); cudaFree(pt->d_total);
    cudaFreeHost(pt->h_total);
    delete pt;
    return FH_OK;
}

extern "C" int fh_ptable_size(const fh_ptable *pt, int *n_terms) {
    FH_REQUIRE(pt && n_terms, "fh_ptable_size: NULL argument");
    *n_terms = pt->count;
    return FH_OK;
}

extern "C" int fh_ptable_download(const fh_ptable *pt, uint64_t *x, uint64_t *z, double *coeff_re, double *coeff_im) {
    FH_REQUIRE(pt && x && z && coeff_re && coeff_im, "fh_ptable_download: NULL argument");
    if (pt->count == 0) return FH_OK;
    FH_CUDA(cudaSetDevice(pt->ctx->device));
    FH_CUDA(cudaStreamSynchronize(pt->ctx->stream));
    std::vector<double2> c((size_t)pt->count);
    FH_CUDA(cudaMemcpy(x, pt->d_x, sizeof(u64) * pt->count, cudaMemcpyDeviceToHost));
    FH_CUDA(cudaMemcpy(z, pt->d_z, sizeof(u64) * pt->count, cudaMemcpyDeviceToHost));
    FH_CUDA(cudaMemcpy(c.data(), pt->d_c, sizeof(double2) * pt->count, cudaMemcpyDeviceToHost));
    for (int t = 0; t < pt->count; ++t) {
        coeff_re[t] = c[t].x;
        coeff_im[t] = c[t].y;
    }
    return FH_OK;
}

// table <- exp(i tau P/2) table exp(-i tau P/2), P = (xp, zp); one host read (the new term count) per call
extern "C" int fh_ptable_dress(fh_ptable *pt, uint64_t xp, uint64_t zp, double tau, double tol) {
    FH_REQUIRE(pt, "fh_ptable_dress: table is NULL");
    FH_REQUIRE(tol >= 0.0, "fh_ptable_dress: negative tolerance");
    const int n = pt->count;
    if (n == 0) return FH_OK;
    fh_ctx *ctx = pt->ctx;
    FH_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    FH_TRY(ptable_scratch(pt, n));
    const unsigned mask = (unsigned)pt->hash_cap - 1u;
    const int g = (n + 255) / 256, nb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    const double cs = cos(tau), sn = sin(tau);
    FH_CUDA(cudaMemsetAsync(pt->d_hash, 0xff, sizeof(int) * pt->hash_cap, s));
    k_hash_build<<<g, 256, 0, s>>>(pt->d_x, pt->d_z, n, pt->d_hash, mask);
    k_dress_mark<<<g, 256, 0, s>>>(pt->d_x, pt->d_z, pt->d_c, n, pt->d_hash, mask, xp, zp, cs, sn, tol, pt->d_cnew,
                                   pt->d_flag_new, pt->d_flag_keep);
    int *blk_keep = pt->d_block, *blk_new = pt->d_block + nb + 1;
    k_scan_blocks<<<nb, SCAN_BLOCK, 0, s>>>(pt->d_flag_keep, n, pt->d_pos_keep, blk_keep);
    k_scan_sums<<<1, SCAN_BLOCK, 0, s>>>(blk_keep, nb, pt->d_total);
    k_scan_blocks<<<nb, SCAN_BLOCK, 0, s>>>(pt->d_flag_new, n, pt->d_pos_new, blk_new);
    k_scan_sums<<<1, SCAN_BLOCK, 0, s>>>(blk_new, nb, pt->d_total + 1);
    k_dress_emit<<<g, 256, 0, s>>>(pt->d_x, pt->d_z, pt->d_c, pt->d_cnew, n, pt->d_flag_keep, pt->d_pos_keep, blk_keep,
                                   pt->d_flag_new, pt->d_pos_new, blk_new, pt->d_total, xp, zp, sn, pt->d_x2, pt->d_z2, pt->d_c2);
    FH_CUDA(cudaGetLastError());
    FH_CUDA(cudaMemcpyAsync(pt->h_total, pt->d_total, sizeof(int) * 2, cudaMemcpyDeviceToHost, s));
    FH_CUDA(cudaStreamSynchronize(s));
    const int out_n = pt->h_total[0] + pt->h_total[1];
    FH_TRY(ptable_reserve(pt, out_n > 0 ? out_n : 1));
    if (out_n) {
        FH_CUDA(cudaMemcpyAsync(pt->d_x, pt->d_x2, sizeof(u64) * out_n, cudaMemcpyDeviceToDevice, s));
        FH_CUDA(cudaMemcpyAsync(pt->d_z, pt->d_z2, sizeof(u64) * out_n, cudaMemcpyDeviceToDevice, s));
        FH_CUDA(cudaMemcpyAsync(pt->d_c, pt->d_c2, sizeof(double2) * out_n, cudaMemcpyDeviceToDevice, s));
    }
    pt->count = out_n;
    return FH_OK;
}
