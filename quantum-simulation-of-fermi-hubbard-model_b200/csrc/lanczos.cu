// K4: matrix-free Lanczos on the fused H|v> kernel (fh_apply_table) -- the replacement for
// scipy.sparse.linalg.eigsh(which='SA') in the reference's linalg/exact_diagonalization.py:34-51,
// 181-229 and for openfermion.get_ground_state (models/iqcc_hubbard.py:57).
//
// Full re-orthogonalisation against the Krylov basis (one batched multi-dot + one multi-axpy launch per
// iteration), deflation against already converged eigenvectors (resolves the 4-fold degenerate 3x3
// ground level), restart from the current Ritz vector when the basis buffer is full.  The start
// vector lives in the (N_up, N_dn) sector; H maps the sector to itself exactly because each x-mask
// group's weight is summed before it multiplies an amplitude (+t/2 - t/2 = 0 exactly).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "common.cuh"

#define FULL 0xffffffffu
#define MD_CHUNKS 32

static long long g_fh_launch_count_lz = 0;
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }

__device__ __forceinline__ double lz_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// partials[(r*MD_CHUNKS + c)*2 + {0,1}] = chunk c of <V_r | w>
__global__ void __launch_bounds__(256) k_multi_dot(const double2 *__restrict__ V, u64 dim, const double2 *__restrict__ w,
                                                   double *__restrict__ partials) {
    __shared__ double red[16];
    const double2 *v = V + (size_t)blockIdx.y * dim;
    const u64 per = (dim + MD_CHUNKS - 1) / MD_CHUNKS;
    const u64 lo = (u64)blockIdx.x * per;
    const u64 hi = lo + per < dim ? lo + per : dim;
    double re = 0.0, im = 0.0;
    for (u64 i = lo + threadIdx.x; i < hi; i += 256) {
        const double2 a = v[i], b = w[i];
        re += a.x * b.x + a.y * b.y;
        im += a.x * b.y - a.y * b.x;
    }
    re = lz_warp_sum(re);
    im = lz_warp_sum(im);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
        red[2 * wid] = re;
        red[2 * wid + 1] = im;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sr = 0.0, si = 0.0;
        for (int k = 0; k < 8; ++k) {
            sr += red[2 * k];
            si += red[2 * k + 1];
        }
        partials[((size_t)blockIdx.y * MD_CHUNKS + blockIdx.x) * 2] = sr;
        partials[((size_t)blockIdx.y * MD_CHUNKS + blockIdx.x) * 2 + 1] = si;
    }
}

__global__ void __launch_bounds__(32) k_multi_dot_finalize(const double *__restrict__ partials, double *__restrict__ coef) {
    const double *p = partials + (size_t)blockIdx.x * MD_CHUNKS * 2;
    double re = threadIdx.x < MD_CHUNKS ? p[2 * threadIdx.x] : 0.0;
    double im = threadIdx.x < MD_CHUNKS ? p[2 * threadIdx.x + 1] : 0.0;
    re = lz_warp_sum(re);
    im = lz_warp_sum(im);
    if (threadIdx.x == 0) {
        coef[2 * blockIdx.x] = re;
        coef[2 * blockIdx.x + 1] = im;
    }
}

// w[i] += sign * sum_r coef[r] * V_r[i]
__global__ void __launch_bounds__(256) k_multi_axpy(double2 *__restrict__ w, const double2 *__restrict__ V, u64 dim,
                                                    const double *__restrict__ coef, int nvec, double sign) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += stride) {
        double re = 0.0, im = 0.0;
        for (int r = 0; r < nvec; ++r) {
            const double cr = __ldg(coef + 2 * r), ci = __ldg(coef + 2 * r + 1);
            const double2 v = V[(size_t)r * dim + i];
            re += cr * v.x - ci * v.y;
            im += cr * v.y + ci * v.x;
        }
        double2 o = w[i];
        o.x += sign * re;
        o.y += sign * im;
        w[i] = o;
    }
}

// ---- host-side tridiagonal helpers ------------------------------------------------------------
// number of eigenvalues of T(alpha, beta) below x (Sturm count via LDL^T pivots)
static int sturm_count(const std::vector<double> &a, const std::vector<double> &b, int m, double x) {
    int count = 0;
    double d = 1.0;
    for (int i = 0; i < m; ++i) {
        const double off = i > 0 ? b[i - 1] * b[i - 1] : 0.0;
        d = a[i] - x - (i > 0 ? off / d : 0.0);
        if (fabs(d) < 1e-300) d = -1e-300;
        if (d < 0) ++count;
    }
    return count;
}

static double lowest_eigenvalue(const std::vector<double> &a, const std::vector<double> &b, int m) {
    double lo = a[0], hi = a[0];
    for (int i = 0; i < m; ++i) {
        const double r = (i > 0 ? fabs(b[i - 1]) : 0.0) + (i < m - 1 ? fabs(b[i]) : 0.0);
        lo = fmin(lo, a[i] - r);
        hi = fmax(hi, a[i] + r);
    }
    for (int it = 0; it < 200 && hi - lo > 1e-15 * fmax(1.0, fabs(lo) + fabs(hi)); ++it) {
        const double mid = 0.5 * (lo + hi);
        if (sturm_count(a, b, m, mid) >= 1) hi = mid; else lo = mid;
    }
    return 0.5 * (lo + hi);
}

// eigenvector of the lowest eigenvalue by inverse iteration (T - shift positive definite -> Thomas is stable)
static void lowest_eigenvector(const std::vector<double> &a, const std::vector<double> &b, int m, double theta,
                               std::vector<double> &y) {
    const double shift = theta - 1e-9 * fmax(1.0, fabs(theta));
    y.assign(m, 1.0 / sqrt((double)m));
    std::vector<double> c(m), d(m);
    for (int it = 0; it < 4; ++it) {
        // forward elimination
        double piv = a[0] - shift;
        c[0] = (m > 1 ? b[0] : 0.0) / piv;
        d[0] = y[0] / piv;
        for (int i = 1; i < m; ++i) {
            piv = a[i] - shift - b[i - 1] * c[i - 1];
            if (fabs(piv) < 1e-300) piv = 1e-300;
            c[i] = (i < m - 1 ? b[i] : 0.0) / piv;
            d[i] = (y[i] - b[i - 1] * d[i - 1]) / piv;
        }
        y[m - 1] = d[m - 1];
        for (int i = m - 2; i >= 0; --i) y[i] = d[i] - c[i] * y[i + 1];
        double nrm = 0.0;
        for (int i = 0; i < m; ++i) nrm += y[i] * y[i];
        nrm = sqrt(nrm);
        for (int i = 0; i < m; ++i) y[i] /= nrm;
    }
}

// ---- driver -------------------------------------------------------------------------------------
struct LzBuffers {
    double2 *V = nullptr, *F = nullptr, *w = nullptr, *x = nullptr;
    double *partials = nullptr, *coef = nullptr, *h_coef = nullptr;
    ~LzBuffers() {
        cudaFree(V);
        cudaFree(F);
        cudaFree(w);
        cudaFree(x);
        cudaFree(partials);
        cudaFree(coef);
        cudaFreeHost(h_coef);
    }
};

static int project_out(fh_ctx *ctx, double2 *w, const double2 *basis, int nvec, u64 dim, LzBuffers &B) {
    if (nvec <= 0) return FH_OK;
    for (int off = 0; off < nvec; off += 32768) {
        const int cnt = nvec - off < 32768 ? nvec - off : 32768;
        dim3 grid(MD_CHUNKS, cnt);
        k_multi_dot<<<grid, 256, 0, ctx->stream>>>(basis + (size_t)off * dim, dim, w, B.partials);
        k_multi_dot_finalize<<<cnt, 32, 0, ctx->stream>>>(B.partials, B.coef);
        int g = (int)((dim + 255) / 256);
        if (g > ctx->sm_count * 8) g = ctx->sm_count * 8;
        k_multi_axpy<<<g, 256, 0, ctx->stream>>>(w, basis + (size_t)off * dim, dim, B.coef, cnt, -1.0);
    }
    FH_CUDA(cudaGetLastError());
    return FH_OK;
}

static int device_norm(fh_ctx *ctx, const double2 *v, u64 dim, double *out) {
    launch_inner(ctx->stream, ctx->sm_count, v, v, dim, ctx->d_partials, ctx->d_result);
    FH_CUDA(cudaMemcpyAsync(ctx->h_result, ctx->d_result, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FH_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = sqrt(fmax(ctx->h_result[0], 0.0));
    return FH_OK;
}

// =================================================================================================
// Sector-compressed, device-resident Lanczos
// =================================================================================================
// The drivers only ever ask for eigenpairs inside one (N_up, N_dn) sector (linalg/exact_diagonalization.py:34-51 restricts
// the sparse matrix to it), and the Hubbard Hamiltonian conserves both numbers.  So the Krylov vectors are stored
// COMPRESSED: one amplitude per sector state, rank r = rank_up * D_dn + rank_dn, full index i(r) = up_bits[rank_up] |
// dn_bits[rank_dn] (up orbitals = even wires = odd index bits for even n).  3x3: 15 876 instead of 262 144 amplitudes;
// 4x4: 165 636 900 amplitudes = 2.65 GB instead of 64 GiB, i.e. the 32-qubit ground state fits ONE B200.
//   k_sector_matvec   w = H v on compressed vectors (partner rank through two 2^(n/2)-entry rank tables), fused <v|H|v>
//   CGS2 re-orthogonalisation, norm and normalisation all read their scalars from device memory, so a whole batch of
//   iterations is enqueued without a host round trip; alpha / beta come back once per batch for the Ritz test.
struct SectorGeom {
    int n, half;                 // qubits, orbitals per spin
    unsigned d_up, d_dn;         // sector dimensions per spin
    u64 dim;                     // d_up * d_dn
    const unsigned *up_bits, *dn_bits;      // [d_up], [d_dn]: deposited index bits of each pattern (ascending)
    const unsigned *rank_up, *rank_dn;      // [2^half]: rank of a pattern, 0xffffffff if its popcount is wrong
};

__device__ __forceinline__ unsigned compress_even_bits(unsigned x) {      // bits 0,2,4,.. -> bits 0,1,2,..
    x &= 0x55555555u;
    x = (x | (x >> 1)) & 0x33333333u;
    x = (x | (x >> 2)) & 0x0f0f0f0fu;
    x = (x | (x >> 4)) & 0x00ff00ffu;
    x = (x | (x >> 8)) & 0x0000ffffu;
    return x;
}

template <bool REAL>
__global__ void __launch_bounds__(256) k_sector_matvec(const SectorGeom sg, const TabGroup *__restrict__ groups, int ngroups,
                                                       const TabClass *__restrict__ classes, int nclasses,
                                                       const double2 *__restrict__ vals, int nvals, int use_smem,
                                                       const double2 *__restrict__ dtab, const double2 *__restrict__ in,
                                                       double2 *__restrict__ out, double *__restrict__ partials,
                                                       unsigned *__restrict__ counter, double *__restrict__ alpha_out,
                                                       unsigned *__restrict__ err_flag) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[8];
    const TabGroup *G = groups;
    const TabClass *Cl = classes;
    const double2 *V = vals;
    if (use_smem) {
        uint4 *dst = reinterpret_cast<uint4 *>(smem_raw);
        const int ng4 = ngroups * (int)(sizeof(TabGroup) / 16), nc4 = nclasses, nv4 = nvals;
        const uint4 *s_g = reinterpret_cast<const uint4 *>(groups), *s_c = reinterpret_cast<const uint4 *>(classes),
                    *s_v = reinterpret_cast<const uint4 *>(vals);
        for (int t = threadIdx.x; t < ng4; t += blockDim.x) dst[t] = __ldg(s_g + t);
        for (int t = threadIdx.x; t < nc4; t += blockDim.x) dst[ng4 + t] = __ldg(s_c + t);
        for (int t = threadIdx.x; t < nv4; t += blockDim.x) dst[ng4 + nc4 + t] = __ldg(s_v + t);
        __syncthreads();
        G = reinterpret_cast<const TabGroup *>(dst);
        Cl = reinterpret_cast<const TabClass *>(dst + ng4);
        V = reinterpret_cast<const double2 *>(dst + ng4 + nc4);
    }
    double er = 0.0;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < sg.dim; r += stride) {
        const unsigned ru = (unsigned)(r / sg.d_dn), rd = (unsigned)(r - (u64)ru * sg.d_dn);
        const unsigned i = __ldg(sg.up_bits + ru) | __ldg(sg.dn_bits + rd);
        const double2 self = in[r];
        double ar = 0.0, ai = 0.0;
        if (dtab) {
            const double2 d = cadd(cadd(__ldg(dtab + (i & 0xfffu)), __ldg(dtab + 4096 + ((i >> 12) & 0xfffu))),
                                   __ldg(dtab + 8192 + (i >> 24)));
            ar = d.x * self.x - d.y * self.y;
            ai = d.x * self.y + d.y * self.x;
        }
        for (int g = 0; g < ngroups; ++g) {
            const uint4 g0 = reinterpret_cast<const uint4 *>(G + g)[0];       // x lo, x hi, first_class, n_class
            const unsigned p = reinterpret_cast<const uint2 *>(G + g)[2].x;  // pos[4]
            const unsigned live = reinterpret_cast<const uint4 *>(G + g)[1].w;
            const unsigned j = i ^ g0.x;
            const unsigned pat = ((j >> (p & 63u)) & 1u) | (((j >> ((p >> 8) & 63u)) & 1u) << 1) |
                                 (((j >> ((p >> 16) & 63u)) & 1u) << 2) | (((j >> ((p >> 24) & 63u)) & 1u) << 3);
            if (!((live >> pat) & 1u)) continue;
            double wr = 0.0, wi = 0.0;
            const int c0 = (int)g0.z, c1 = c0 + (int)g0.w;
            for (int c = c0; c < c1; ++c) {
                const TabClass cl = Cl[c];
                const double sgn = (__popc(j & (unsigned)cl.zeta) & 1) ? -1.0 : 1.0;
                const double2 v = V[cl.vofs + pat];
                wr += sgn * v.x;
                if (!REAL) wi += sgn * v.y;
            }
            if (wr == 0.0 && (REAL || wi == 0.0)) continue;
            double2 pv = self;
            if (g0.x != 0u) {
                const unsigned qu = __ldg(sg.rank_up + compress_even_bits(j >> 1)), qd = __ldg(sg.rank_dn + compress_even_bits(j));
                if ((qu | qd) == 0xffffffffu) {          // a term that leaves the sector: not a sector-conserving table
                    *err_flag = 1u;
                    continue;
                }
                pv = in[(u64)qu * sg.d_dn + qd];
            }
            if (REAL) {
                ar += wr * pv.x;
                ai += wr * pv.y;
            } else {
                ar += wr * pv.x - wi * pv.y;
                ai += wr * pv.y + wi * pv.x;
            }
        }
        out[r] = make_double2(ar, ai);
        er += self.x * ar + self.y * ai;
    }
    // <v|H|v>: per-CTA partial, folded in slot order by the last CTA (deterministic)
    er = lz_warp_sum(er);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) red[wid] = er;
    __syncthreads();
    __shared__ unsigned is_last;
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int k = 0; k < 8; ++k) s += red[k];
        partials[blockIdx.x] = s;
        __threadfence();
        is_last = (atomicAdd(counter, 1u) == gridDim.x - 1u) ? 1u : 0u;
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double s = 0.0;
        for (int t = threadIdx.x; t < (int)gridDim.x; t += blockDim.x) s += __ldcg(partials + t);
        s = lz_warp_sum(s);
        if (lane == 0) red[wid] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0.0;
            for (int k = 0; k < 8; ++k) tot += red[k];
            *alpha_out = tot;
            *counter = 0u;
        }
    }
}

__device__ __forceinline__ u64 lz_mix64(u64 z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256) k_sector_gauss(double2 *v, u64 dim, u64 seed) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < dim; r += stride) {
        const u64 h1 = lz_mix64(r * 2 + seed * 0x100000001b3ull), h2 = lz_mix64(h1 + 1);
        const double u1 = ((h1 >> 11) + 1.0) * (1.0 / 9007199254740993.0), u2 = (h2 >> 11) * (1.0 / 9007199254740992.0);
        v[r] = make_double2(sqrt(-2.0 * log(u1)) * cospi(2.0 * u2), 0.0);
    }
}

// dst = src / sqrt(*norm2) (zeros if the norm vanished: invariant subspace reached)
__global__ void __launch_bounds__(256) k_scale_by_dev_norm(double2 *__restrict__ dst, const double2 *__restrict__ src, u64 dim,
                                                           const double *__restrict__ norm2) {
    const double n2 = *norm2;
    const double f = n2 > 1e-280 ? rsqrt(n2) : 0.0;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < dim; r += stride) {
        const double2 a = src[r];
        dst[r] = make_double2(a.x * f, a.y * f);
    }
}

// full[i(r)] = x[r]  (full must be zeroed before)
__global__ void __launch_bounds__(256) k_sector_scatter(const SectorGeom sg, const double2 *__restrict__ x, double2 *__restrict__ full) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < sg.dim; r += stride) {
        const unsigned ru = (unsigned)(r / sg.d_dn), rd = (unsigned)(r - (u64)ru * sg.d_dn);
        full[__ldg(sg.up_bits + ru) | __ldg(sg.dn_bits + rd)] = x[r];
    }
}

// does every x != 0 group only connect states of equal (N_up, N_dn)?  (tabulated groups only: <= 4 x bits)
static bool table_conserves_sector(const fh_table *tab, u64 upmask, u64 dnmask) {
    for (const TabGroup &g : tab->groups) {
        if (g.x == 0) continue;
        if (g.kbits == 0) return false;
        for (int pat = 0; pat < (1 << g.kbits); ++pat) {
            if (!((g.live >> pat) & 1u)) continue;
            int dup = 0, ddn = 0;      // change of N_up / N_dn going from the partner j (pattern) to i = j ^ x
            for (int b = 0; b < g.kbits; ++b) {
                const u64 bit = 1ull << g.pos[b];
                const int was = (pat >> b) & 1, now = was ^ 1;
                if (bit & upmask) dup += now - was;
                if (bit & dnmask) ddn += now - was;
            }
            if (dup != 0 || ddn != 0) return false;
        }
    }
    return true;
}

struct SectorBuffers {
    unsigned *up_bits = nullptr, *dn_bits = nullptr, *rank_up = nullptr, *rank_dn = nullptr, *err = nullptr, *counter = nullptr;
    double2 *V = nullptr, *F = nullptr, *w = nullptr, *x = nullptr;
    double *partials = nullptr, *coef = nullptr, *h_coef = nullptr, *mv_partials = nullptr;
    double *d_alpha = nullptr, *d_beta2 = nullptr, *h_scal = nullptr;
    ~SectorBuffers() {
        cudaFree(up_bits); cudaFree(dn_bits); cudaFree(rank_up); cudaFree(rank_dn); cudaFree(err); cudaFree(counter);
        cudaFree(V); cudaFree(F); cudaFree(w); cudaFree(x);
        cudaFree(partials); cudaFree(coef); cudaFreeHost(h_coef); cudaFree(mv_partials);
        cudaFree(d_alpha); cudaFree(d_beta2); cudaFreeHost(h_scal);
    }
};

static void sector_patterns(int half, int count, int shift, std::vector<unsigned> &bits, std::vector<unsigned> &rank) {
    rank.assign((size_t)1 << half, 0xffffffffu);
    bits.clear();
    for (unsigned pat = 0; pat < (1u << half); ++pat) {
        if (__builtin_popcount(pat) != count) continue;
        unsigned dep = 0;
        for (int b = 0; b < half; ++b)
            if (pat >> b & 1u) dep |= 1u << (2 * b + shift);
        rank[pat] = (unsigned)bits.size();
        bits.push_back(dep);
    }
}

static int sector_grid(u64 dim, int sm) {
    u64 g = (dim + 255) / 256;
    if (g > (u64)sm * 8) g = (u64)sm * 8;
    return (int)(g < 1 ? 1 : g);
}

// CGS against nvec compressed vectors with every coefficient kept on the device
static int sector_project_out(fh_ctx *ctx, double2 *w, const double2 *basis, int nvec, u64 dim, SectorBuffers &B) {
    for (int off = 0; off < nvec; off += 16384) {
        const int cnt = nvec - off < 16384 ? nvec - off : 16384;
        dim3 grid(MD_CHUNKS, cnt);
        k_multi_dot<<<grid, 256, 0, ctx->stream>>>(basis + (size_t)off * dim, dim, w, B.partials);
        k_multi_dot_finalize<<<cnt, 32, 0, ctx->stream>>>(B.partials, B.coef);
        k_multi_axpy<<<sector_grid(dim, ctx->sm_count), 256, 0, ctx->stream>>>(w, basis + (size_t)off * dim, dim, B.coef, cnt, -1.0);
    }
    FH_CUDA(cudaGetLastError());
    return FH_OK;
}

// stats (may be NULL): [0] sector dimension, [1] seconds inside the iteration loops, [2] matvecs, [3] host syncs
extern "C" int fh_lanczos_sector(const fh_table *tab, int n_up, int n_dn, int k, double tol, int max_iter, uint64_t seed,
                                 double *evals, fh_state *const *evecs, double *compressed_out, int *iterations,
                                 double *stats) {
    FH_REQUIRE(tab && evals, "fh_lanczos_sector: NULL argument");
    FH_REQUIRE(k >= 1 && k <= 64, "fh_lanczos_sector: k=%d outside [1, 64]", k);
    FH_REQUIRE(max_iter >= 2 && tol > 0, "fh_lanczos_sector: bad max_iter / tol");
    const int n = tab->n, half = n / 2;
    FH_REQUIRE(n % 2 == 0 && n >= 2 && n <= 32, "fh_lanczos_sector: needs an even number of qubits <= 32 (got %d)", n);
    FH_REQUIRE(n_up >= 0 && n_dn >= 0 && n_up <= half && n_dn <= half, "fh_lanczos_sector: bad sector (%d up, %d down)", n_up, n_dn);
    u64 upmask = 0, dnmask = 0;
    for (int q = 0; q < n; ++q) ((q % 2 == 0) ? upmask : dnmask) |= 1ull << (n - 1 - q);
    FH_REQUIRE(table_conserves_sector(tab, upmask, dnmask), "fh_lanczos_sector: the table does not conserve (N_up, N_dn)");
    if (evecs)
        for (int e = 0; e < k; ++e) FH_REQUIRE(evecs[e] && evecs[e]->n == n, "fh_lanczos_sector: evecs[%d] missing or wrong size", e);
    fh_ctx *ctx = tab->ctx;
    FH_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;

    // even wires (up) sit at odd index bits, odd wires (down) at even index bits (n even, wire q <-> bit n-1-q)
    std::vector<unsigned> h_up_bits, h_dn_bits, h_rank_up, h_rank_dn;
    sector_patterns(half, n_up, 1, h_up_bits, h_rank_up);
    sector_patterns(half, n_dn, 0, h_dn_bits, h_rank_dn);
    SectorBuffers B;
    SectorGeom sg;
    sg.n = n;
    sg.half = half;
    sg.d_up = (unsigned)h_up_bits.size();
    sg.d_dn = (unsigned)h_dn_bits.size();
    sg.dim = (u64)sg.d_up * sg.d_dn;
    FH_REQUIRE(sg.dim >= 1, "fh_lanczos_sector: empty sector");
    const u64 dim = sg.dim;
    const size_t vbytes = dim * sizeof(double2);
    FH_CUDA(cudaMalloc(&B.up_bits, sizeof(unsigned) * h_up_bits.size()));
    FH_CUDA(cudaMalloc(&B.dn_bits, sizeof(unsigned) * h_dn_bits.size()));
    FH_CUDA(cudaMalloc(&B.rank_up, sizeof(unsigned) * h_rank_up.size()));
    FH_CUDA(cudaMalloc(&B.rank_dn, sizeof(unsigned) * h_rank_dn.size()));
    FH_CUDA(cudaMemcpy(B.up_bits, h_up_bits.data(), sizeof(unsigned) * h_up_bits.size(), cudaMemcpyHostToDevice));
    FH_CUDA(cudaMemcpy(B.dn_bits, h_dn_bits.data(), sizeof(unsigned) * h_dn_bits.size(), cudaMemcpyHostToDevice));
    FH_CUDA(cudaMemcpy(B.rank_up, h_rank_up.data(), sizeof(unsigned) * h_rank_up.size(), cudaMemcpyHostToDevice));
    FH_CUDA(cudaMemcpy(B.rank_dn, h_rank_dn.data(), sizeof(unsigned) * h_rank_dn.size(), cudaMemcpyHostToDevice));
    sg.up_bits = B.up_bits;
    sg.dn_bits = B.dn_bits;
    sg.rank_up = B.rank_up;
    sg.rank_dn = B.rank_dn;

    // Krylov basis: as long as memory allows up to 160 vectors (no restart needed for the small lattices), at least 8
    size_t free_b = 0, total_b = 0;
    FH_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const size_t fixed = vbytes * (size_t)(k + 3);
    FH_REQUIRE(free_b > fixed + 8 * vbytes, "fh_lanczos_sector: not enough device memory for sector vectors of %llu amplitudes",
               (unsigned long long)dim);
    size_t cap = (size_t)((free_b - fixed) * 0.6 / vbytes);
    size_t want = dim <= (1ull << 22) ? 160 : 24;
    if (const char *env = getenv("FHSIM_LANCZOS_BASIS")) want = (size_t)atoi(env);
    if (want < 8) want = 8;
    if (cap > want) cap = want;
    if (cap > (size_t)max_iter) cap = max_iter;
    if (cap > dim) cap = dim;
    FH_REQUIRE(cap >= 2, "fh_lanczos_sector: room for only %zu Krylov vectors", cap);
    const int mcap = (int)cap;
    const int check_every = dim <= (1ull << 22) ? 8 : 1;

    FH_CUDA(cudaMalloc(&B.V, vbytes * mcap));
    FH_CUDA(cudaMalloc(&B.F, vbytes * k));
    FH_CUDA(cudaMalloc(&B.w, vbytes));
    FH_CUDA(cudaMalloc(&B.x, vbytes));
    const int maxvec = mcap > k ? mcap : k;
    FH_CUDA(cudaMalloc(&B.partials, sizeof(double) * 2 * MD_CHUNKS * (size_t)maxvec));
    FH_CUDA(cudaMalloc(&B.coef, sizeof(double) * 2 * (maxvec + 1)));
    FH_CUDA(cudaMallocHost(&B.h_coef, sizeof(double) * 2 * (maxvec + 1)));
    FH_CUDA(cudaMalloc(&B.d_alpha, sizeof(double) * (mcap + 1)));
    FH_CUDA(cudaMalloc(&B.d_beta2, sizeof(double) * 2 * (mcap + 1)));
    FH_CUDA(cudaMallocHost(&B.h_scal, sizeof(double) * 3 * (mcap + 2)));
    FH_CUDA(cudaMalloc(&B.err, sizeof(unsigned)));
    FH_CUDA(cudaMalloc(&B.counter, sizeof(unsigned)));
    FH_CUDA(cudaMemset(B.err, 0, sizeof(unsigned)));
    FH_CUDA(cudaMemset(B.counter, 0, sizeof(unsigned)));
    const int mv_grid = sector_grid(dim, ctx->sm_count);
    FH_CUDA(cudaMalloc(&B.mv_partials, sizeof(double) * mv_grid));

    const int ngroups = (int)tab->groups.size(), nclasses = (int)tab->classes.size(), nvals = (int)tab->vals.size();
    const size_t need = (size_t)ngroups * sizeof(TabGroup) + (size_t)nclasses * sizeof(TabClass) + (size_t)nvals * sizeof(double2);
    const int use_smem = need <= 40 * 1024;
    auto matvec = [&](const double2 *v, double2 *w, double *alpha_dev) {
        ++g_fh_launch_count_lz;
        if (tab->all_real)
            k_sector_matvec<true><<<mv_grid, 256, use_smem ? need : 0, s>>>(sg, tab->d_groups, ngroups, tab->d_classes, nclasses,
                                                                            tab->d_vals, nvals, use_smem, tab->d_diag, v, w,
                                                                            B.mv_partials, B.counter, alpha_dev, B.err);
        else
            k_sector_matvec<false><<<mv_grid, 256, use_smem ? need : 0, s>>>(sg, tab->d_groups, ngroups, tab->d_classes, nclasses,
                                                                             tab->d_vals, nvals, use_smem, tab->d_diag, v, w,
                                                                             B.mv_partials, B.counter, alpha_dev, B.err);
    };
    // <a|a> into dst[0] (dst[1] = imaginary part, zero) through the batched-dot kernels
    auto norm2_to = [&](const double2 *a, double *dst) {
        dim3 grid(MD_CHUNKS, 1);
        k_multi_dot<<<grid, 256, 0, s>>>(a, dim, a, B.partials);
        k_multi_dot_finalize<<<1, 32, 0, s>>>(B.partials, dst);
    };
    const int vgrid = sector_grid(dim, ctx->sm_count);

    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    FH_CUDA(cudaEventCreate(&ev0));
    FH_CUDA(cudaEventCreate(&ev1));
    double loop_ms = 0.0;
    long long matvecs = 0, syncs = 0;
    int total_iters = 0;
    std::vector<double> alpha, beta, y;
    int rc = FH_OK;
    for (int e = 0; e < k && rc == FH_OK; ++e) {
        k_sector_gauss<<<vgrid, 256, 0, s>>>(B.x, dim, seed + 0x9e37ull * (u64)e);
        for (int pass = 0; pass < 2; ++pass) rc = rc == FH_OK ? sector_project_out(ctx, B.x, B.F, e, dim, B) : rc;
        norm2_to(B.x, B.d_beta2);
        k_scale_by_dev_norm<<<vgrid, 256, 0, s>>>(B.V, B.x, dim, B.d_beta2);
        double theta = 0.0;
        bool converged = false;
        int iters_this = 0;
        while (!converged && iters_this < max_iter && rc == FH_OK) {
            // one Lanczos run from V[0] (already normalised), enqueued in batches of check_every iterations
            alpha.clear();
            beta.clear();
            int m_done = 0, m_conv = -1;
            cudaEventRecord(ev0, s);
            while (m_done < mcap && iters_this < max_iter && m_conv < 0) {
                const int batch_end = m_done + check_every < mcap ? m_done + check_every : mcap;
                for (int m = m_done; m < batch_end; ++m) {
                    const double2 *v = B.V + (size_t)m * dim;
                    matvec(v, B.w, B.d_alpha + m);
                    ++matvecs;
                    for (int pass = 0; pass < 2; ++pass) {          // full re-orthogonalisation (CGS2) + deflation
                        rc = rc == FH_OK ? sector_project_out(ctx, B.w, B.V, m + 1, dim, B) : rc;
                        rc = rc == FH_OK ? sector_project_out(ctx, B.w, B.F, e, dim, B) : rc;
                    }
                    norm2_to(B.w, B.d_beta2 + 2 * m);
                    if (m + 1 < mcap)
                        k_scale_by_dev_norm<<<vgrid, 256, 0, s>>>(B.V + (size_t)(m + 1) * dim, B.w, dim, B.d_beta2 + 2 * m);
                }
                if (rc != FH_OK) break;
                cudaMemcpyAsync(B.h_scal, B.d_alpha, sizeof(double) * batch_end, cudaMemcpyDeviceToHost, s);
                cudaMemcpyAsync(B.h_scal + (mcap + 2), B.d_beta2, sizeof(double) * 2 * batch_end, cudaMemcpyDeviceToHost, s);
                if (cudaStreamSynchronize(s) != cudaSuccess) { rc = FH_ECUDA; fh_set_error("fh_lanczos_sector: %s", cudaGetErrorString(cudaGetLastError())); break; }
                ++syncs;
                for (int m = m_done; m < batch_end && m_conv < 0; ++m) {
                    alpha.push_back(B.h_scal[m]);
                    const double bnorm = sqrt(fmax(B.h_scal[(mcap + 2) + 2 * m], 0.0));
                    const int mm = m + 1;
                    theta = lowest_eigenvalue(alpha, beta, mm);
                    lowest_eigenvector(alpha, beta, mm, theta, y);
                    ++iters_this;
                    if (fabs(bnorm * y[mm - 1]) < tol * fmax(1.0, fabs(theta)) || bnorm < 1e-13) {
                        m_conv = mm;
                        break;
                    }
                    beta.push_back(bnorm);
                }
                m_done = batch_end;
            }
            cudaEventRecord(ev1, s);
            if (rc != FH_OK) break;
            const int mv = m_conv > 0 ? m_conv : (int)alpha.size();
            if (m_conv > 0) converged = true;
            else {                                       // basis full (or iteration budget spent): Ritz pair of everything seen
                beta.resize(mv > 0 ? mv - 1 : 0);
                theta = lowest_eigenvalue(alpha, beta, mv);
                lowest_eigenvector(alpha, beta, mv, theta, y);
            }
            // Ritz vector x = sum_r y_r V_r, cleaned against the converged eigenvectors, normalised, restart vector V[0]
            for (int r = 0; r < mv; ++r) {
                B.h_coef[2 * r] = y[r];
                B.h_coef[2 * r + 1] = 0.0;
            }
            cudaMemcpyAsync(B.coef, B.h_coef, sizeof(double) * 2 * mv, cudaMemcpyHostToDevice, s);
            cudaMemsetAsync(B.x, 0, vbytes, s);
            k_multi_axpy<<<vgrid, 256, 0, s>>>(B.x, B.V, dim, B.coef, mv, 1.0);
            rc = sector_project_out(ctx, B.x, B.F, e, dim, B);
            norm2_to(B.x, B.d_beta2);
            k_scale_by_dev_norm<<<vgrid, 256, 0, s>>>(B.V, B.x, dim, B.d_beta2);
            if (cudaStreamSynchronize(s) != cudaSuccess) { rc = FH_ECUDA; fh_set_error("fh_lanczos_sector: %s", cudaGetErrorString(cudaGetLastError())); break; }
            ++syncs;
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev0, ev1);
            loop_ms += ms;
        }
        if (rc != FH_OK) break;
        if (!converged) {
            fh_set_error("fh_lanczos_sector: eigenpair %d not converged to %.1e in %d iterations", e, tol, max_iter);
            rc = FH_EINVAL;
            break;
        }
        total_iters += iters_this;
        evals[e] = theta;
        cudaMemcpyAsync(B.F + (size_t)e * dim, B.V, vbytes, cudaMemcpyDeviceToDevice, s);     // V[0] holds the normalised Ritz vector
        if (evecs) {
            cudaMemsetAsync(evecs[e]->d, 0, sizeof(double2) << n, s);
            k_sector_scatter<<<vgrid, 256, 0, s>>>(sg, B.V, evecs[e]->d);
        }
        if (compressed_out)
            cudaMemcpyAsync(compressed_out + 2 * (size_t)e * dim, B.V, vbytes, cudaMemcpyDeviceToHost, s);
    }
    unsigned h_err = 0;
    cudaMemcpyAsync(&h_err, B.err, sizeof(unsigned), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    if (rc == FH_OK && cudaGetLastError() != cudaSuccess) rc = FH_ECUDA;
    if (rc == FH_OK && h_err) {
        fh_set_error("fh_lanczos_sector: the table connects the sector to states outside it");
        rc = FH_EINVAL;
    }
    if (iterations) *iterations = total_iters;
    if (stats) {
        stats[0] = (double)dim;
        stats[1] = loop_ms * 1e-3;
        stats[2] = (double)matvecs;
        stats[3] = (double)syncs;
    }
    return rc;
}

extern "C" int fh_lanczos(const fh_table *tab, int n_up, int n_dn, int k, double tol, int max_iter, uint64_t seed,
                          double *evals, fh_state *const *evecs, int *iterations) {
    FH_REQUIRE(tab && evals, "fh_lanczos: NULL argument");
    FH_REQUIRE(k >= 1 && k <= 64, "fh_lanczos: k=%d outside [1, 64]", k);
    FH_REQUIRE(max_iter >= 2, "fh_lanczos: max_iter must be >= 2");
    FH_REQUIRE(tol > 0, "fh_lanczos: tol must be positive");
    const int n = tab->n;
    FH_REQUIRE((n_up < 0 && n_dn < 0) || (n_up >= 0 && n_dn >= 0 && n_up <= (n + 1) / 2 && n_dn <= n / 2),
               "fh_lanczos: bad sector (%d up, %d down) for %d qubits", n_up, n_dn, n);
    if (evecs)
        for (int e = 0; e < k; ++e) FH_REQUIRE(evecs[e] && evecs[e]->n == n, "fh_lanczos: evecs[%d] missing or wrong size", e);
    // sector requests on number-conserving tables (every ED call of the drivers) run on compressed vectors
    if (n_up >= 0 && n % 2 == 0 && n <= 32 && !getenv("FHSIM_LANCZOS_FULL")) {
        u64 upmask = 0, dnmask = 0;
        for (int q = 0; q < n; ++q) ((q % 2 == 0) ? upmask : dnmask) |= 1ull << (n - 1 - q);
        if (table_conserves_sector(tab, upmask, dnmask))
            return fh_lanczos_sector(tab, n_up, n_dn, k, tol, max_iter, seed, evals, evecs, nullptr, iterations, nullptr);
    }
    fh_ctx *ctx = tab->ctx;
    FH_CUDA(cudaSetDevice(ctx->device));
    const u64 dim = 1ull << n;
    const size_t vbytes = dim * sizeof(double2);

    // Krylov basis capacity from free memory
    size_t free_b = 0, total_b = 0;
    FH_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const size_t fixed = vbytes * (size_t)(k + 2);
    FH_REQUIRE(free_b > fixed + 3 * vbytes, "fh_lanczos: not enough device memory for %d-qubit vectors", n);
    size_t cap = (size_t)((free_b - fixed) * 0.7 / vbytes);
    if (cap > (size_t)max_iter) cap = max_iter;
    // Full re-orthogonalisation costs O(m) vector passes per iteration, so the basis is kept short and Lanczos is
    // restarted from the current Ritz vector when it is full (measured on 3x3: 32 is 2-4x faster than an unbounded
    // basis for the 4-fold degenerate level; tests/perf_lanczos.py)
    size_t basis_cap = 32;
    if (const char *env = getenv("FHSIM_LANCZOS_BASIS")) basis_cap = (size_t)atoi(env);
    if (basis_cap < 8) basis_cap = 8;
    if (cap > basis_cap) cap = basis_cap;
    FH_REQUIRE(cap >= 3, "fh_lanczos: room for only %zu Krylov vectors", cap);
    const int mcap = (int)cap;

    LzBuffers B;
    FH_CUDA(cudaMalloc(&B.V, vbytes * mcap));
    FH_CUDA(cudaMalloc(&B.F, vbytes * k));
    FH_CUDA(cudaMalloc(&B.w, vbytes));
    FH_CUDA(cudaMalloc(&B.x, vbytes));
    const int maxvec = mcap > k ? mcap : k;
    FH_CUDA(cudaMalloc(&B.partials, sizeof(double) * 2 * MD_CHUNKS * (size_t)(maxvec < 32768 ? maxvec : 32768)));
    FH_CUDA(cudaMalloc(&B.coef, sizeof(double) * 2 * maxvec));
    FH_CUDA(cudaMallocHost(&B.h_coef, sizeof(double) * 2 * maxvec));

    int total_iters = 0;
    std::vector<double> alpha, beta, y;
    for (int e = 0; e < k; ++e) {
        // start vector: random in the sector, orthogonal to the converged eigenvectors
        launch_sector_random(ctx->stream, ctx->sm_count, B.x, n, n_up, n_dn, seed + 0x9e37ull * (u64)e);
        FH_TRY(project_out(ctx, B.x, B.F, e, dim, B));
        double nrm = 0.0;
        FH_TRY(device_norm(ctx, B.x, dim, &nrm));
        FH_REQUIRE(nrm > 1e-12, "fh_lanczos: sector is empty or exhausted (eigenpair %d)", e);
        launch_scale(ctx->stream, ctx->sm_count, B.x, 1.0 / nrm, dim);

        double theta = 0.0, prev_theta = 1e300;
        bool converged = false;
        int iters_this = 0;
        while (!converged && iters_this < max_iter) {
            // (re)start Lanczos from B.x
            alpha.clear();
            beta.clear();
            FH_CUDA(cudaMemcpyAsync(B.V, B.x, vbytes, cudaMemcpyDeviceToDevice, ctx->stream));
            int m = 0;
            for (; m < mcap && iters_this < max_iter; ++iters_this) {
                double2 *v = B.V + (size_t)m * dim;
                // w = H v ; alpha = <v|H|v>
                launch_apply_table(ctx->stream, ctx->sm_count, tab, v, B.w, 1, ctx->d_partials, ctx->d_result);
                // full re-orthogonalisation (covers the three-term recurrence) + deflation, applied twice
                for (int pass = 0; pass < 2; ++pass) {
                    FH_TRY(project_out(ctx, B.w, B.V, m + 1, dim, B));
                    FH_TRY(project_out(ctx, B.w, B.F, e, dim, B));
                }
                launch_inner(ctx->stream, ctx->sm_count, B.w, B.w, dim, ctx->d_partials, ctx->d_result + 2);
                FH_CUDA(cudaMemcpyAsync(ctx->h_result, ctx->d_result, 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
                FH_CUDA(cudaStreamSynchronize(ctx->stream));
                alpha.push_back(ctx->h_result[0]);
                const double bnorm = sqrt(fmax(ctx->h_result[2], 0.0));
                ++m;
                theta = lowest_eigenvalue(alpha, beta, m);
                lowest_eigenvector(alpha, beta, m, theta, y);
                const double resid = fabs(bnorm * y[m - 1]);
                if (resid < tol * fmax(1.0, fabs(theta)) || bnorm < 1e-13) {
                    converged = true;
                    ++iters_this;
                    break;
                }
                prev_theta = theta;
                if (m < mcap) {
                    beta.push_back(bnorm);
                    FH_CUDA(cudaMemcpyAsync(B.V + (size_t)m * dim, B.w, vbytes, cudaMemcpyDeviceToDevice, ctx->stream));
                    launch_scale(ctx->stream, ctx->sm_count, B.V + (size_t)m * dim, 1.0 / bnorm, dim);
                }
            }
            // Ritz vector x = sum_r y_r V_r
            const int mv = (int)alpha.size();
            for (int r = 0; r < mv; ++r) {
                B.h_coef[2 * r] = y[r];
                B.h_coef[2 * r + 1] = 0.0;
            }
            FH_CUDA(cudaMemcpyAsync(B.coef, B.h_coef, sizeof(double) * 2 * mv, cudaMemcpyHostToDevice, ctx->stream));
            FH_CUDA(cudaMemsetAsync(B.x, 0, vbytes, ctx->stream));
            int g = (int)((dim + 255) / 256);
            if (g > ctx->sm_count * 8) g = ctx->sm_count * 8;
            k_multi_axpy<<<g, 256, 0, ctx->stream>>>(B.x, B.V, dim, B.coef, mv, 1.0);
            FH_TRY(project_out(ctx, B.x, B.F, e, dim, B));
            FH_TRY(device_norm(ctx, B.x, dim, &nrm));
            launch_scale(ctx->stream, ctx->sm_count, B.x, 1.0 / nrm, dim);
            FH_CUDA(cudaStreamSynchronize(ctx->stream));   // h_coef reused by the next restart
        }
        (void)prev_theta;
        FH_REQUIRE(converged, "fh_lanczos: eigenpair %d not converged to %.1e in %d iterations", e, tol, max_iter);
        total_iters += iters_this;
        evals[e] = theta;
        FH_CUDA(cudaMemcpyAsync(B.F + (size_t)e * dim, B.x, vbytes, cudaMemcpyDeviceToDevice, ctx->stream));
        if (evecs) FH_CUDA(cudaMemcpyAsync(evecs[e]->d, B.x, vbytes, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    FH_CUDA(cudaStreamSynchronize(ctx->stream));
    if (iterations) *iterations = total_iters;
    return FH_OK;
}
