// Hand-written sm_100a kernels of the complex128 statevector hot path.
//
// Everything here is HBM/L2-bandwidth work on 16-byte amplitudes: one complex128 == one 128-bit
// LDG/STG (double2), warps walk consecutive indices so every request is a full 32-byte sector and,
// whenever the fixed bits of an op are not the lowest ones, full 128-byte lines.  No tensor cores:
// nothing on this path is a dense contraction.
//
//   k_pair / k_pair_adjoint   K1: 2x2 rotations on index pairs (Pauli / fermionic / Givens / CNOT)
//   k_diag / k_diag_adjoint   K1: diagonal phase ops (RZ, Z-string rotations)
//   k_tile                    K1: run of ops fused in a shared-memory tile of 2^T amplitudes
//   k_apply_table             K2: out = H in fused with <in|H|in>
//   k_pool / k_pool_finalize  K3: batched pool gradients 2 Im <lambda|G_k|psi>
//   k_inner, k_axpby, ...     K4: Lanczos vector kernels
#include <cooperative_groups.h>

#include "common.cuh"
#include "devmath.cuh"

// number of kernel launches issued through the wrappers below (measurement; read via fh_program_last_stats)
long long g_fh_launch_count = 0;
thread_local int g_fh_tile_pdl_scope = 0;     // > 0 while fh_program_evaluate enqueues its kernels (see launch_tile)

// ----------------------------------------------------------------------------------------------
// K1: pair op on the whole state
// ----------------------------------------------------------------------------------------------
template <int UNROLL>
__global__ void __launch_bounds__(128) k_pair(double2 *__restrict__ psi, const PairOp *__restrict__ opp, u64 npairs,
                                              int dagger) {
    __shared__ PairOp op;
    load_pair_op(&op, opp);
    const Mat2 M = op_matrix(op, dagger);
    const u64 x = op.x, fixval = op.fixval, zeta = op.zeta;
    const int npos = op.npos;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 base = (u64)blockIdx.x * blockDim.x + threadIdx.x; base < npairs; base += stride * UNROLL) {
        u64 ii[UNROLL];
        double2 a[UNROLL], b[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const u64 idx = base + (u64)u * stride;
            if (idx < npairs) {
                ii[u] = deposit_zeros(idx, op.pos, npos) | fixval;
                a[u] = psi[ii[u]];
                b[u] = psi[ii[u] ^ x];
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const u64 idx = base + (u64)u * stride;
            if (idx < npairs) {
                rot2(M, sign_of(ii[u] & zeta), a[u], b[u]);
                psi[ii[u]] = a[u];
                psi[ii[u] ^ x] = b[u];
            }
        }
    }
}

// adjoint step of a rotation op U = exp(-i a Ghat): partial of Im<lam|Ghat|psi> at the current
// (post-op) states, then psi <- U^dagger psi, lam <- U^dagger lam.  One partial per block.
__global__ void __launch_bounds__(128) k_pair_adjoint(double2 *__restrict__ psi, double2 *__restrict__ lam,
                                                      const PairOp *__restrict__ opp, u64 npairs,
                                                      double *__restrict__ partials) {
    __shared__ PairOp op;
    __shared__ double red[4];
    load_pair_op(&op, opp);
    const Mat2 M = op_matrix(op, 1);
    const double2 bh = make_double2(op.bhat[0], op.bhat[1]);
    const u64 x = op.x, fixval = op.fixval, zeta = op.zeta;
    const int npos = op.npos;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    double acc = 0.0;
    for (u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x; idx < npairs; idx += stride) {
        const u64 i = deposit_zeros(idx, op.pos, npos) | fixval, j = i ^ x;
        double2 a = psi[i], b = psi[j], la = lam[i], lb = lam[j];
        const double sgn = sign_of(i & zeta);
        const double2 gi = cscale(cmul(bh, b), sgn);           // (Ghat psi)_i
        const double2 gj = cscale(cmul(cconj(bh), a), sgn);    // (Ghat psi)_j
        acc += im_conj_mul(la, gi) + im_conj_mul(lb, gj);
        rot2(M, sgn, a, b);
        rot2(M, sgn, la, lb);
        psi[i] = a;
        psi[j] = b;
        lam[i] = la;
        lam[j] = lb;
    }
    const double r = block_sum<128>(acc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = r;
}

// ----------------------------------------------------------------------------------------------
// K1: diagonal op on the whole state
// ----------------------------------------------------------------------------------------------
#define DIAG_SMEM_TERMS 512

__device__ __forceinline__ double2 diag_phase(const DiagTerm *t, int nterms, u64 i, int dagger) {
    double2 ph;
    if (nterms <= 4) {
        ph = make_double2(1.0, 0.0);
        for (int m = 0; m < nterms; ++m) {
            const double sg = sign_of(i & t[m].z);
            ph = cmul(ph, make_double2(t[m].c, -sg * t[m].s));
        }
    } else {
        double tot = 0.0;
        for (int m = 0; m < nterms; ++m) tot += sign_of(i & t[m].z) * t[m].angle;
        double s, c;
        sincos(tot, &s, &c);
        ph = make_double2(c, -s);
    }
    if (dagger) ph.y = -ph.y;
    return ph;
}

__global__ void __launch_bounds__(256) k_diag(double2 *__restrict__ psi, const DiagTerm *__restrict__ terms, int nterms,
                                              u64 dim, int dagger) {
    __shared__ DiagTerm st[DIAG_SMEM_TERMS];
    for (int t = threadIdx.x; t < nterms; t += blockDim.x) st[t] = terms[t];
    __syncthreads();
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += stride) {
        const double2 ph = diag_phase(st, nterms, i, dagger);
        psi[i] = cmul(ph, psi[i]);
    }
}

// adjoint step of exp(-i theta D), D(i) = sum coef_m sgn_m: partial of Im<lam|D|psi>, then undo on both
__global__ void __launch_bounds__(256) k_diag_adjoint(double2 *__restrict__ psi, double2 *__restrict__ lam,
                                                      const DiagTerm *__restrict__ terms, int nterms, u64 dim,
                                                      double *__restrict__ partials) {
    __shared__ DiagTerm st[DIAG_SMEM_TERMS];
    __shared__ double red[8];
    for (int t = threadIdx.x; t < nterms; t += blockDim.x) st[t] = terms[t];
    __syncthreads();
    const u64 stride = (u64)gridDim.x * blockDim.x;
    double acc = 0.0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += stride) {
        double d = 0.0;
        for (int m = 0; m < nterms; ++m) d += sign_of(i & st[m].z) * st[m].coef;
        const double2 a = psi[i], l = lam[i];
        acc += d * im_conj_mul(l, a);
        const double2 ph = diag_phase(st, nterms, i, 1);
        psi[i] = cmul(ph, a);
        lam[i] = cmul(ph, l);
    }
    const double r = block_sum<256>(acc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = r;
}

// ----------------------------------------------------------------------------------------------
// K1: fused run of ops inside a shared-memory tile
// ----------------------------------------------------------------------------------------------
// A tile = all 2^T amplitudes that differ only in the T tile bits.  Each CTA stages one tile in shared
// memory (plus the global index of every slot), applies the whole run of ops to it with one
// __syncthreads per op, and writes it back: one global read + one write for the whole run.
// The op descriptors of the run are gathered into shared memory by all threads in parallel first, so
// the per-op loop never waits on global memory.
#define TILE_MAX_SUB FH_TILE_MAX_SUB
#define TILE_MAX_TERMS FH_TILE_MAX_TERMS

__device__ __forceinline__ double2 tile_diag_phase(const TileTerm *t, int nterms, unsigned gi) {
    double2 ph;
    if (nterms <= 4) {
        ph = make_double2(1.0, 0.0);
        for (int m = 0; m < nterms; ++m) {
            const double sg = (__popc(gi & (unsigned)t[m].z) & 1) ? -1.0 : 1.0;
            ph = cmul(ph, make_double2(t[m].c, -sg * t[m].s));
        }
    } else {
        double tot = 0.0;
        for (int m = 0; m < nterms; ++m) tot += ((__popc(gi & (unsigned)t[m].z) & 1) ? -1.0 : 1.0) * t[m].angle;
        double sn, cs;
        sincos(tot, &sn, &cs);
        ph = make_double2(cs, -sn);
    }
    return ph;
}

__device__ __forceinline__ int n_terms_of(const TileRec &r) { return r.nterms; }

// Shared-memory slot of tile-local index l.  The three low bits (the 128-bit bank group) are XOR-folded with every
// higher 3-bit group, so eight lanes that differ in ANY three bit positions of distinct residue mod 3 hit eight
// different bank groups: consecutive indices (tile load/store) and the strided accesses of the register runs alike.
__device__ __forceinline__ unsigned tile_slot(unsigned l) {
    return l ^ ((l >> 3) & 7u) ^ ((l >> 6) & 7u) ^ ((l >> 9) & 7u) ^ ((l >> 12) & 7u);
}

// Optional in-kernel timeline (make TIMELINE=1 -> -DFH_TILE_TIMELINE; tools/probe_timeline.py): thread 0 of CTA 0
// stamps clock64() at the phase boundaries of tile_run into g_fh_tile_timeline, read back by fh_debug_tile_timeline.
// Slots: 0 entry, 1 prologue done (records, tables[, first batch] + barrier), 2 tile in shared memory, 4+k start of op k
// (k < 40), 3 ops done, 63 tile stored.  Compiled out by default (the macro expands to nothing).
#ifdef FH_TILE_TIMELINE
__device__ long long g_fh_tile_timeline[64];
#define FH_TLMARK(k)                                                              \
    do {                                                                          \
        if (blockIdx.x == 0 && threadIdx.x == 0) g_fh_tile_timeline[k] = clock64(); \
    } while (0)
extern "C" int fh_debug_tile_timeline(long long *out64) {
    return (int)cudaMemcpyFromSymbol(out64, g_fh_tile_timeline, sizeof(long long) * 64);
}
#else
#define FH_TLMARK(k) do { } while (0)
#endif

// PDL = launched with programmatic stream serialization (launch_tile): the op records and the scatter tables are set up
// while the previous kernel of the stream / graph is still draining, and only then griddepcontrol.wait orders the tile
// loads after that kernel's writes.  Without PDL the first tile batch is issued before the record copy instead.
template <bool PDL>
__device__ __forceinline__ void tile_run(double2 *__restrict__ psi, const TileLaunch &tl,
                                         const TileRec *__restrict__ recs, const TileTerm *__restrict__ terms, int n,
                                         unsigned char *smem_raw) {
    __shared__ TileRec rec[TILE_MAX_SUB];
    __shared__ TileTerm tterm[TILE_MAX_TERMS];
    __shared__ unsigned slo[64], shi[128];      // scatter tables: local index bits -> global bit positions
    __shared__ double2 ph[192];                 // diagonal ops: phase factor tables over local bits 0..5 / 6..12
    FH_TLMARK(0);
    const int T = tl.nbits, nsub = tl.nsub;
    unsigned lomask_g = 0, himask_g = 0;        // global masks of the tile's local bits 0..5 / 6..
    lomask_g = tile_mask(tl, T, 0, 6);
    himask_g = tile_mask(tl, T, 6, TILE_BITS_CAP);
    const unsigned L = 1u << T;
    double2 *buf = reinterpret_cast<double2 *>(smem_raw);
    unsigned int *gidx = reinterpret_cast<unsigned int *>(buf + L);

    // ---- prologue ----
    // The first batch of the CTA's first tile is addressed straight from tl.bits (no shared-memory table yet) and its
    // four 128-bit loads are issued BEFORE the record copy and the table construction, so the two global-memory
    // latencies of a launch (op records, tile) overlap instead of adding up and one barrier covers both.  Measured
    // with clock64 at 18 qubits: prologue 2 100-2 500 + tile load 2 900 cycles before, see DESIGN 7.
    const u64 ntiles = 1ull << (n - T);
    const bool have_tile = !PDL && (u64)blockIdx.x < ntiles;
    unsigned gg0[4];
    double2 vv0[4];
    if (PDL) asm volatile("griddepcontrol.launch_dependents;");     // the next launch may start its own prologue
    if (have_tile) {
        const unsigned base0 = (unsigned)tile_base(tl, T, (u64)blockIdx.x);
        // scatter of this thread's batch-0 local index
        const unsigned gt = tile_scatter(tl, T, threadIdx.x, 0, TILE_BITS_CAP);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned l = threadIdx.x + u * blockDim.x;
            if (l < L) {
                // blockDim is a power of two > threadIdx.x: the bits of u * blockDim are disjoint from the thread's
                const unsigned gu = tile_scatter(tl, T, u * blockDim.x, 0, TILE_BITS_CAP);
                gg0[u] = base0 | gt | gu;
                vv0[u] = psi[gg0[u]];
            }
        }
    }
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(recs + tl.first_rec);
        uint4 *dst = reinterpret_cast<uint4 *>(rec);
        const int chunks = nsub * (int)(sizeof(TileRec) / 16);
        for (int c = threadIdx.x; c < chunks; c += blockDim.x) dst[c] = __ldg(src + c);
        const uint4 *tsrc = reinterpret_cast<const uint4 *>(terms + tl.first_term);
        uint4 *tdst = reinterpret_cast<uint4 *>(tterm);
        const int tchunks = tl.nterms * (int)(sizeof(TileTerm) / 16);
        for (int c = threadIdx.x; c < tchunks; c += blockDim.x) tdst[c] = __ldg(tsrc + c);
    }
    for (unsigned v = threadIdx.x; v < 64u; v += blockDim.x) slo[v] = tile_scatter(tl, T, v, 0, 6);
    for (unsigned v = threadIdx.x; v < 128u; v += blockDim.x) shi[v] = tile_scatter(tl, T, v, 6, TILE_BITS_CAP);
    if (have_tile) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned l = threadIdx.x + u * blockDim.x;
            if (l < L) {
                const unsigned sl = tile_slot(l);
                gidx[sl] = gg0[u];
                buf[sl] = vv0[u];
            }
        }
    }
    __syncthreads();

    FH_TLMARK(1);
    bool first = true;
    if (PDL) {
        asm volatile("griddepcontrol.wait;" ::: "memory");          // the previous kernel's amplitudes are complete
        first = false;                                                // nothing was pre-loaded
    }
    for (u64 t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const unsigned base = (unsigned)tile_base(tl, T, t);
        // four independent 128-bit loads in flight per thread before anything is written to shared memory
        // (batch 0 of the first tile is already in place)
        for (unsigned l0 = threadIdx.x + (first ? 4u * blockDim.x : 0u); l0 < L; l0 += 4 * blockDim.x) {
            unsigned gg[4];
            double2 vv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned l = l0 + u * blockDim.x;
                if (l < L) {
                    gg[u] = base | slo[l & 63u] | shi[l >> 6];
                    vv[u] = psi[gg[u]];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned l = l0 + u * blockDim.x;
                if (l < L) {
                    const unsigned sl = tile_slot(l);
                    gidx[sl] = gg[u];
                    buf[sl] = vv[u];
                }
            }
        }
        if (!first || L > 4u * blockDim.x) __syncthreads();     // CTA-uniform condition
        first = false;
        FH_TLMARK(2);
        int sidx = 0;
        while (sidx < nsub) {
            if (sidx < 40) FH_TLMARK(4 + sidx);
            const uint4 *rp = reinterpret_cast<const uint4 *>(&rec[sidx]);
            const uint4 q0 = rp[0], q1 = rp[1], q2 = rp[2];
            // q0 = {fixmask_out, fixval_out, zeta, xlocal}; q1 = {lfixval, type, nlfix, term_off}; q2 = lowmask[4]
            const int type = (int)q1.y;
            if (type != 2) {
                if ((base & q0.x) == q0.y) {
                    const double2 *mp = reinterpret_cast<const double2 *>(rec[sidx].m);
                    const double2 ma = mp[0], mb = mp[1], mc = mp[2], md = mp[3];
                    const unsigned npairs = L >> q1.z;
                    // warp-uniform per op: sign of the out-of-tile bits, slot-space image of the x-mask (the slot map
                    // is XOR-linear), in-tile part of zeta in local coordinates
                    const unsigned sbase = (unsigned)__popc(base & q0.z);
                    const unsigned zloc = rp[3].z;
                    const unsigned xs = tile_slot(q0.w);
                    const int nlfix = (int)q1.z;
                    for (unsigned k = threadIdx.x; k < npairs; k += blockDim.x) {
                        unsigned il = k;
                        il = ((il & ~q2.x) << 1) | (il & q2.x);
                        il = ((il & ~q2.y) << 1) | (il & q2.y);
                        if (nlfix > 2) {
                            il = ((il & ~q2.z) << 1) | (il & q2.z);
                            il = ((il & ~q2.w) << 1) | (il & q2.w);
                        }
                        il |= q1.x;
                        const double sg = ((sbase + (unsigned)__popc(il & zloc)) & 1u) ? -1.0 : 1.0;
                        il = tile_slot(il);
                        const unsigned jl = il ^ xs;
                        double2 a = buf[il], b = buf[jl];
                        if (type == 3) {
                            const double s01 = sg * mb.x, s10 = sg * mc.x;
                            const double2 ra = make_double2(ma.x * a.x + s01 * b.x, ma.x * a.y + s01 * b.y);
                            const double2 rb = make_double2(s10 * a.x + md.x * b.x, s10 * a.y + md.x * b.y);
                            a = ra;
                            b = rb;
                        } else if (type == 6) {
                            // real diagonal (every rotation): ra = c a + s m01 b, rb = s m10 a + c' b
                            const double2 sb = cscale(b, sg), sa = cscale(a, sg);
                            const double2 ra = make_double2(ma.x * a.x + (mb.x * sb.x - mb.y * sb.y),
                                                            ma.x * a.y + (mb.x * sb.y + mb.y * sb.x));
                            const double2 rb = make_double2(md.x * b.x + (mc.x * sa.x - mc.y * sa.y),
                                                            md.x * b.y + (mc.x * sa.y + mc.y * sa.x));
                            a = ra;
                            b = rb;
                        } else {
                            Mat2 M;
                            M.m00 = ma; M.m01 = mb; M.m10 = mc; M.m11 = md;
                            rot2(M, sg, a, b);
                        }
                        buf[il] = a;
                        buf[jl] = b;
                    }
                }
                sidx += 1;
            } else {
                // diagonal op: exp(-i sum_m angle_m sgn_m(index)).  Terms whose in-tile z bits sit entirely in local
                // bits 0..5 (or entirely in 6..) are folded into two small phase tables built once per tile; only
                // terms straddling both halves are evaluated per amplitude.
                const TileTerm *dt = tterm + (int)q1.w;
                const int cnt = (int)n_terms_of(rec[sidx]);
                for (unsigned v = threadIdx.x; v < 192u; v += blockDim.x) {
                    const bool lo = v < 64u;
                    const unsigned gl = base | (lo ? slo[v] : shi[v - 64u]);
                    double tot = 0.0;
                    for (int m = 0; m < cnt; ++m) {
                        const unsigned z = (unsigned)dt[m].z;
                        const bool in_lo = (z & himask_g) == 0u;
                        const bool in_hi = (z & lomask_g) == 0u && !in_lo;
                        if (lo ? in_lo : in_hi) tot += ((__popc(gl & z) & 1) ? -1.0 : 1.0) * dt[m].angle;
                    }
                    double sn, cs;
                    sincos(tot, &sn, &cs);
                    ph[v] = make_double2(cs, -sn);
                }
                __syncthreads();
                for (unsigned sl = threadIdx.x; sl < L; sl += blockDim.x) {
                    const unsigned l = tile_slot(sl);
                    double2 f = cmul(ph[l & 63u], ph[64u + (l >> 6)]);
                    const unsigned gi = gidx[sl];
                    for (int m = 0; m < cnt; ++m) {
                        const unsigned z = (unsigned)dt[m].z;
                        if ((z & himask_g) != 0u && (z & lomask_g) != 0u) {
                            const double sg = (__popc(gi & z) & 1) ? -1.0 : 1.0;
                            f = cmul(f, make_double2(dt[m].c, -sg * dt[m].s));
                        }
                    }
                    buf[sl] = cmul(f, buf[sl]);
                }
                sidx += 1;
            }
            __syncthreads();
        }
        FH_TLMARK(3);
        for (unsigned l = threadIdx.x; l < L; l += blockDim.x) {
            const unsigned sl = tile_slot(l);
            psi[gidx[sl]] = buf[sl];
        }
        __syncthreads();
        FH_TLMARK(63);
    }
}

template <bool PDL>
__global__ void __launch_bounds__(512, 2) k_tile(double2 *__restrict__ psi, const TileLaunch tl,
                                                 const TileRec *__restrict__ recs, const TileTerm *__restrict__ terms,
                                                 int n) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    tile_run<PDL>(psi, tl, recs, terms, n, smem_raw);
}

// Several consecutive tile runs in ONE cooperative launch: a grid-wide barrier replaces the kernel boundary between
// runs (the 18-qubit regime is launch-latency bound: 13 runs for the forward pass of the benchmark circuit).
__global__ void __launch_bounds__(512, 2) k_tile_multi(double2 *__restrict__ psi, const TileLaunch *__restrict__ tls,
                                                       int nl, const TileRec *__restrict__ recs,
                                                       const TileTerm *__restrict__ terms, int n) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ TileLaunch tl;
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    for (int l = 0; l < nl; ++l) {
        if (threadIdx.x < sizeof(TileLaunch) / 4)
            reinterpret_cast<unsigned *>(&tl)[threadIdx.x] = reinterpret_cast<const unsigned *>(tls + l)[threadIdx.x];
        __syncthreads();
        tile_run<false>(psi, tl, recs, terms, n, smem_raw);
        if (l + 1 < nl) grid.sync();
    }
}

// Fused adjoint sweep of one tile run.  Both psi and lambda tiles are staged in shared memory; the run's records
// (already inverted: dagger order, dagger matrices) are walked once.  For every parametrised op the partial
// Im <lambda| Ghat |psi> over the tile is taken at the op's output side (before undoing it), then the inverse op is
// applied to both states.  Partials: warp sums go to shared memory without extra barriers, are folded per tile in a
// fixed order, accumulated over the CTA's tiles, and written once per (op, CTA): deterministic.
#define TILE_ADJ_WARPS 16
__global__ void __launch_bounds__(512, 1) k_tile_adjoint(double2 *__restrict__ psi, double2 *__restrict__ lam,
                                                         const TileLaunch tl, const TileRec *__restrict__ recs,
                                                         const TileTerm *__restrict__ terms, int n,
                                                         double *__restrict__ gpart, int seg_base) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ TileRec rec[TILE_MAX_SUB];
    __shared__ TileTerm tterm[TILE_MAX_TERMS];
    __shared__ unsigned slo[64], shi[128];
    __shared__ double2 ph[192];
    __shared__ double dsum[192];                          // diagonal generator value tables (sum of coef * sign)
    __shared__ double wsum[TILE_MAX_SUB][TILE_ADJ_WARPS];  // per-op warp partials of the current tile
    __shared__ double cacc[TILE_MAX_SUB];                 // per-op partial of this CTA over all its tiles
    const int T = tl.nbits, nsub = tl.nsub;
    const unsigned L = 1u << T;
    double2 *bufp = reinterpret_cast<double2 *>(smem_raw);
    double2 *bufl = bufp + L;
    unsigned int *gidx = reinterpret_cast<unsigned int *>(bufl + L);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    unsigned lomask_g = 0, himask_g = 0;
    lomask_g = tile_mask(tl, T, 0, 6);
    himask_g = tile_mask(tl, T, 6, TILE_BITS_CAP);
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(recs + tl.first_rec);
        uint4 *dst = reinterpret_cast<uint4 *>(rec);
        const int chunks = nsub * (int)(sizeof(TileRec) / 16);
        for (int c = threadIdx.x; c < chunks; c += blockDim.x) dst[c] = __ldg(src + c);
        const uint4 *tsrc = reinterpret_cast<const uint4 *>(terms + tl.first_term);
        uint4 *tdst = reinterpret_cast<uint4 *>(tterm);
        const int tchunks = tl.nterms * (int)(sizeof(TileTerm) / 16);
        for (int c = threadIdx.x; c < tchunks; c += blockDim.x) tdst[c] = __ldg(tsrc + c);
    }
    for (unsigned v = threadIdx.x; v < 64u; v += blockDim.x) slo[v] = tile_scatter(tl, T, v, 0, 6);
    for (unsigned v = threadIdx.x; v < 128u; v += blockDim.x) shi[v] = tile_scatter(tl, T, v, 6, TILE_BITS_CAP);
    for (int o = threadIdx.x; o < TILE_MAX_SUB; o += blockDim.x) cacc[o] = 0.0;
    __syncthreads();

    const u64 ntiles = 1ull << (n - T);
    for (u64 t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const unsigned base = (unsigned)tile_base(tl, T, t);
        for (unsigned l = threadIdx.x; l < L; l += blockDim.x) {
            const unsigned g = base | slo[l & 63u] | shi[l >> 6];
            const unsigned sl = tile_slot(l);
            const double2 a = psi[g], b = lam[g];
            gidx[sl] = g;
            bufp[sl] = a;
            bufl[sl] = b;
        }
        __syncthreads();
        for (int sidx = 0; sidx < nsub; ++sidx) {
            const uint4 *rp = reinterpret_cast<const uint4 *>(&rec[sidx]);
            const uint4 q0 = rp[0], q1 = rp[1], q2 = rp[2];
            const int type = (int)q1.y;
            const int seg = (int)rp[3].y;
            double acc = 0.0;
            if (type != 2) {
                if ((base & q0.x) == q0.y) {
                    const double2 *mp = reinterpret_cast<const double2 *>(rec[sidx].m);
                    Mat2 M;
                    M.m00 = mp[0]; M.m01 = mp[1]; M.m10 = mp[2]; M.m11 = mp[3];
                    const double2 bh = mp[4];
                    const double2 bhc = cconj(bh);
                    const unsigned npairs = L >> q1.z;
                    for (unsigned k = threadIdx.x; k < npairs; k += blockDim.x) {
                        unsigned il = k;
                        il = ((il & ~q2.x) << 1) | (il & q2.x);
                        il = ((il & ~q2.y) << 1) | (il & q2.y);
                        il = ((il & ~q2.z) << 1) | (il & q2.z);
                        il = ((il & ~q2.w) << 1) | (il & q2.w);
                        il |= q1.x;
                        const unsigned jl = tile_slot(il ^ q0.w);
                        il = tile_slot(il);
                        double2 a = bufp[il], b = bufp[jl], la = bufl[il], lb = bufl[jl];
                        const double sg = (__popc(gidx[il] & q0.z) & 1) ? -1.0 : 1.0;
                        if (seg >= 0) {
                            const double2 gi = cscale(cmul(bh, b), sg);      // (Ghat psi)_i
                            const double2 gj = cscale(cmul(bhc, a), sg);     // (Ghat psi)_j
                            acc += im_conj_mul(la, gi) + im_conj_mul(lb, gj);
                        }
                        rot2(M, sg, a, b);
                        rot2(M, sg, la, lb);
                        bufp[il] = a;
                        bufp[jl] = b;
                        bufl[il] = la;
                        bufl[jl] = lb;
                    }
                }
            } else {
                const TileTerm *dt = tterm + (int)q1.w;
                const int cnt = (int)n_terms_of(rec[sidx]);
                for (unsigned v = threadIdx.x; v < 192u; v += blockDim.x) {
                    const bool lo = v < 64u;
                    const unsigned gl = base | (lo ? slo[v] : shi[v - 64u]);
                    double tot = 0.0, dv = 0.0;
                    for (int m = 0; m < cnt; ++m) {
                        const unsigned z = (unsigned)dt[m].z;
                        const bool in_lo = (z & himask_g) == 0u;
                        const bool in_hi = (z & lomask_g) == 0u && !in_lo;
                        if (lo ? in_lo : in_hi) {
                            const double sg = (__popc(gl & z) & 1) ? -1.0 : 1.0;
                            tot += sg * dt[m].angle;
                            dv += sg * dt[m].coef;
                        }
                    }
                    double sn, cs;
                    sincos(tot, &sn, &cs);
                    ph[v] = make_double2(cs, -sn);
                    dsum[v] = dv;
                }
                __syncthreads();
                for (unsigned sl = threadIdx.x; sl < L; sl += blockDim.x) {
                    const unsigned l = tile_slot(sl);
                    double2 f = cmul(ph[l & 63u], ph[64u + (l >> 6)]);
                    double d = dsum[l & 63u] + dsum[64u + (l >> 6)];
                    const unsigned gi = gidx[sl];
                    for (int m = 0; m < cnt; ++m) {
                        const unsigned z = (unsigned)dt[m].z;
                        if ((z & himask_g) != 0u && (z & lomask_g) != 0u) {
                            const double sg = (__popc(gi & z) & 1) ? -1.0 : 1.0;
                            f = cmul(f, make_double2(dt[m].c, -sg * dt[m].s));
                            d += sg * dt[m].coef;
                        }
                    }
                    const double2 a = bufp[sl], la = bufl[sl];
                    if (seg >= 0) acc += d * im_conj_mul(la, a);
                    bufp[sl] = cmul(f, a);
                    bufl[sl] = cmul(f, la);
                }
            }
            if (seg >= 0) {
                acc = warp_sum(acc);
                if (lane == 0) wsum[sidx][warp] = acc;
            }
            __syncthreads();
        }
        for (unsigned l = threadIdx.x; l < L; l += blockDim.x) {
            const unsigned sl = tile_slot(l);
            const unsigned g = gidx[sl];
            psi[g] = bufp[sl];
            lam[g] = bufl[sl];
        }
        // fold this tile's warp partials (fixed order) into the CTA accumulators
        for (int o = threadIdx.x; o < nsub; o += blockDim.x) {
            if (reinterpret_cast<const uint4 *>(&rec[o])[3].y != 0xffffffffu) {
                double v = 0.0;
                for (int w = 0; w < nwarps; ++w) v += wsum[o][w];
                cacc[o] += v;
            }
        }
        __syncthreads();
    }
    for (int o = threadIdx.x; o < nsub; o += blockDim.x) {
        const int seg = (int)reinterpret_cast<const uint4 *>(&rec[o])[3].y;
        if (seg >= 0) gpart[(size_t)(seg_base + seg) * FH_GRAD_BLOCKS + blockIdx.x] = cacc[o];
    }
}

// ----------------------------------------------------------------------------------------------
// K2: out = H in, e = <in|H|in>
// ----------------------------------------------------------------------------------------------
// One thread owns output index i and walks the x-mask groups.  The weight of in[i^x_g] is a sum over
// the group's classes of (-1)^popcount(j & zeta_c) * V_c[pattern(j)] (see TabGroup in common.cuh):
// for a hopping group that is one parity + one 4-entry table lookup, and half of all (i, group)
// combinations have weight 0 and skip the gather.  Tables live in shared memory.  Memory traffic is
// one (mostly L1/L2-resident) gather per contributing group plus one store -- independent of the
// term count.
template <bool REAL, int MODE>
__global__ void __launch_bounds__(256) k_apply_table(const TabGroup *__restrict__ groups, int ngroups,
                                                     const TabClass *__restrict__ classes, int nclasses,
                                                     const double2 *__restrict__ vals, int nvals, int use_smem,
                                                     const double2 *__restrict__ in, double2 *__restrict__ out, u64 dim,
                                                     double *__restrict__ partials, const double2 *__restrict__ dtab) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[8];
    const TabGroup *G = groups;
    const TabClass *Cl = classes;
    const double2 *V = vals;
    if (use_smem) {
        uint4 *dst = reinterpret_cast<uint4 *>(smem_raw);
        const int ng4 = ngroups * (int)(sizeof(TabGroup) / 16), nc4 = nclasses, nv4 = nvals;
        const uint4 *sg = reinterpret_cast<const uint4 *>(groups);
        const uint4 *sc = reinterpret_cast<const uint4 *>(classes);
        const uint4 *sv = reinterpret_cast<const uint4 *>(vals);
        for (int t = threadIdx.x; t < ng4; t += blockDim.x) dst[t] = __ldg(sg + t);
        for (int t = threadIdx.x; t < nc4; t += blockDim.x) dst[ng4 + t] = __ldg(sc + t);
        for (int t = threadIdx.x; t < nv4; t += blockDim.x) dst[ng4 + nc4 + t] = __ldg(sv + t);
        __syncthreads();
        G = reinterpret_cast<const TabGroup *>(dst);
        Cl = reinterpret_cast<const TabClass *>(dst + ng4);
        V = reinterpret_cast<const double2 *>(dst + ng4 + nc4);
    }
    const u64 stride = (u64)gridDim.x * blockDim.x;
    double er = 0.0, ei = 0.0;
    constexpr int B = 8;          // groups per batch: weights first, then B independent gathers in flight, then the FMAs
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += stride) {
        double2 acc = make_double2(0.0, 0.0);
        for (int gb = 0; gb < ngroups; gb += B) {
            u64 jj[B];
            double wr[B], wi[B];
#pragma unroll
            for (int u = 0; u < B; ++u) {
                wr[u] = 0.0;
                wi[u] = 0.0;
                jj[u] = i;
                if (gb + u < ngroups) {
                    const uint4 g0 = reinterpret_cast<const uint4 *>(G + gb + u)[0];   // x (lo, hi), first_class, n_class
                    const unsigned p = reinterpret_cast<const uint2 *>(G + gb + u)[2].x;   // pos[4]
                    const u64 j = i ^ ((u64)g0.x | ((u64)g0.y << 32));
                    const unsigned pat = (unsigned)((j >> (p & 63u)) & 1ull) |
                                         ((unsigned)((j >> ((p >> 8) & 63u)) & 1ull) << 1) |
                                         ((unsigned)((j >> ((p >> 16) & 63u)) & 1ull) << 2) |
                                         ((unsigned)((j >> ((p >> 24) & 63u)) & 1ull) << 3);
                    const int c0 = (int)g0.z, c1 = c0 + (int)g0.w;
                    for (int c = c0; c < c1; ++c) {
                        const TabClass cl = Cl[c];
                        const double sg = sign_of(j & cl.zeta);
                        if (REAL) {
                            wr[u] += sg * V[cl.vofs + pat].x;
                        } else {
                            const double2 v = V[cl.vofs + pat];
                            wr[u] += sg * v.x;
                            wi[u] += sg * v.y;
                        }
                    }
                    jj[u] = j;
                }
            }
            double2 vv[B];
#pragma unroll
            for (int u = 0; u < B; ++u) {
                const bool live = REAL ? (wr[u] != 0.0) : (wr[u] != 0.0 || wi[u] != 0.0);
                vv[u] = live ? in[jj[u]] : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int u = 0; u < B; ++u) {
                if (REAL) {
                    acc.x += wr[u] * vv[u].x;
                    acc.y += wr[u] * vv[u].y;
                } else {
                    acc = cadd(acc, cmul(make_double2(wr[u], wi[u]), vv[u]));
                }
            }
        }
        const double2 self = in[i];
        if (dtab) {     // tabulated diagonal part
            const double2 d = cadd(cadd(__ldg(dtab + (i & 0xfffull)), __ldg(dtab + 4096 + ((i >> 12) & 0xfffull))),
                                   __ldg(dtab + 8192 + (i >> 24)));
            acc = cadd(acc, cmul(d, self));
        }
        er += self.x * acc.x + self.y * acc.y;     // conj(self) * acc
        ei += self.x * acc.y - self.y * acc.x;
        if (MODE == 1) out[i] = acc;
        if (MODE == 2) out[i] = cadd(out[i], acc);
    }
    const double sr = block_sum<256>(er, red);
    const double si = block_sum<256>(ei, red);
    if (threadIdx.x == 0) {
        partials[2 * blockIdx.x] = sr;
        partials[2 * blockIdx.x + 1] = si;
    }
}

// K2, main variant (n >= 10): each thread owns the FOUR outputs i0 | (r << 8), r = 0..3 (i0 has bits 8, 9 clear),
// so everything that does not depend on bits 8/9 of the index -- group decode, partner index, x-bit pattern,
// the parity over zeta -- is computed once per four amplitudes; per amplitude only a sign flip, the table
// entry (shared unless the group's x-mask contains bit 8 or 9) and the gather remain.  Consecutive lanes own
// consecutive indices, so every gather of a warp is one contiguous 512-byte run.
template <typename IDX, bool REAL, int MODE, int RL>
__global__ void __launch_bounds__(256) k_apply_table4(const TabGroup *__restrict__ groups, int ngroups,
                                                      const TabClass *__restrict__ classes, int nclasses,
                                                      const double2 *__restrict__ vals, int nvals,
                                                      const double2 *__restrict__ in, double2 *__restrict__ out,
                                                      unsigned nblk, double *__restrict__ partials,
                                                      const double2 *__restrict__ dtab, unsigned *__restrict__ counter,
                                                      double *__restrict__ result, int slow_bits) {
    constexpr int R = 1 << RL;                    // outputs per thread: i0 | (r << 8), r < R
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[8];
    uint4 *dst = reinterpret_cast<uint4 *>(smem_raw);
    const int ng4 = ngroups * 2, nc4 = nclasses, nv4 = nvals;
    {
        const uint4 *sg = reinterpret_cast<const uint4 *>(groups);
        const uint4 *sc = reinterpret_cast<const uint4 *>(classes);
        const uint4 *sv = reinterpret_cast<const uint4 *>(vals);
        for (int t = threadIdx.x; t < ng4; t += blockDim.x) dst[t] = __ldg(sg + t);
        for (int t = threadIdx.x; t < nc4; t += blockDim.x) dst[ng4 + t] = __ldg(sc + t);
        for (int t = threadIdx.x; t < nv4; t += blockDim.x) dst[ng4 + nc4 + t] = __ldg(sv + t);
        __syncthreads();
    }
    const uint4 *G = dst;
    const uint4 *Cl = dst + ng4;
    const double2 *V = reinterpret_cast<const double2 *>(dst + ng4 + nc4);
    double er = 0.0, ei = 0.0;
    // Traversal order (states larger than L2): the CTAs that are resident at the same time work on blocks that differ in
    // the TOP index bits, and the few remaining (slow) bits just above the block advance in Gray-code order.  The window of
    // amplitudes being read at any moment then spans all high bits -- the partners i ^ x of every group whose x-mask
    // avoids the slow bits are inside it (L2 hits instead of DRAM gathers) -- and consecutive windows differ in ONE slow
    // bit, so the partner window of that bit is the one that was just streamed.  slow_bits = 0: linear order.
    const unsigned fast_mask = slow_bits ? ((nblk >> slow_bits) - 1u) : 0xffffffffu;
    for (unsigned q = blockIdx.x; q < nblk; q += gridDim.x) {
        unsigned blk = q;
        if (slow_bits) {
            const unsigned hi = q >> (31 - __clz(nblk >> slow_bits));       // q / (nblk >> slow_bits)
            blk = ((q & fast_mask) << slow_bits) | (hi ^ (hi >> 1));
        }
        const IDX i0 = ((IDX)blk << (8 + RL)) | (IDX)threadIdx.x;
        double ar[R], ai[R];
#pragma unroll
        for (int r = 0; r < R; ++r) ar[r] = ai[r] = 0.0;
        for (int g = 0; g < ngroups; ++g) {
            const uint4 g0 = G[2 * g];            // x lo, x hi, first_class, n_class
            const uint4 g1 = G[2 * g + 1];        // pos[4], kbits, rpat (8 x 4 bits), -
            const IDX j0 = i0 ^ (sizeof(IDX) == 4 ? (IDX)g0.x : (IDX)((u64)g0.x | ((u64)g0.y << 32)));
            const unsigned p = g1.x;
            const unsigned pat0 = (unsigned)((j0 >> (p & 63u)) & 1u) | ((unsigned)((j0 >> ((p >> 8) & 63u)) & 1u) << 1) |
                                  ((unsigned)((j0 >> ((p >> 16) & 63u)) & 1u) << 2) |
                                  ((unsigned)((j0 >> ((p >> 24) & 63u)) & 1u) << 3);
            const unsigned rpat = g1.z;
            // all R outputs see the same x-bit pattern unless the mask contains bit 8..10: a pattern no class has a
            // weight for (half of all cases for a hopping group) skips the class loop and the gathers
            if (rpat == 0u && !((g1.w >> pat0) & 1u)) continue;
            if (REAL && rpat == 0u && g0.w == 1u) {
                // one class, one shared table entry (every hopping group): the R weights are +-v, so only the sign
                // bits are formed and applied to the gathered amplitudes -- no weight array, no liveness tests
                const uint4 cl = Cl[g0.z];
                unsigned sb;
                if (sizeof(IDX) == 4) sb = __popc((unsigned)j0 & cl.x);
                else sb = __popcll((u64)j0 & ((u64)cl.x | ((u64)cl.y << 32)));
                const unsigned zr = cl.x >> 8;
                const double v = V[cl.z + pat0].x;
                double2 vv[R];
#pragma unroll
                for (int r = 0; r < R; ++r) vv[r] = in[j0 ^ ((IDX)r << 8)];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int f = (int)((sb + __popc(zr & (unsigned)r)) << 31);
                    const double w = __hiloint2double(__double2hiint(v) ^ f, __double2loint(v));
                    ar[r] += w * vv[r].x;
                    ai[r] += w * vv[r].y;
                }
                continue;
            }
            double wr[R], wi[R];
#pragma unroll
            for (int r = 0; r < R; ++r) wr[r] = wi[r] = 0.0;
            const int c0 = (int)g0.z, c1 = c0 + (int)g0.w;
            for (int c = c0; c < c1; ++c) {
                const uint4 cl = Cl[c];           // zeta lo, zeta hi, vofs, -
                unsigned sb;
                if (sizeof(IDX) == 4) sb = __popc((unsigned)j0 & cl.x);
                else sb = __popcll((u64)j0 & ((u64)cl.x | ((u64)cl.y << 32)));
                const unsigned zr = cl.x >> 8;    // zeta bits 8.. decide the sign flips between the R outputs
                if (rpat == 0u) {
                    const double2 v = V[cl.z + pat0];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int f = (int)((sb + __popc(zr & (unsigned)r)) << 31);
                        wr[r] += __hiloint2double(__double2hiint(v.x) ^ f, __double2loint(v.x));
                        if (!REAL) wi[r] += __hiloint2double(__double2hiint(v.y) ^ f, __double2loint(v.y));
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const double2 v = V[cl.z + (pat0 ^ ((rpat >> (4 * r)) & 15u))];
                        const int f = (int)((sb + __popc(zr & (unsigned)r)) << 31);
                        wr[r] += __hiloint2double(__double2hiint(v.x) ^ f, __double2loint(v.x));
                        if (!REAL) wi[r] += __hiloint2double(__double2hiint(v.y) ^ f, __double2loint(v.y));
                    }
                }
            }
            double2 vv[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const bool live = REAL ? (wr[r] != 0.0) : (wr[r] != 0.0 || wi[r] != 0.0);
                vv[r] = live ? in[j0 ^ ((IDX)r << 8)] : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (REAL) {
                    ar[r] += wr[r] * vv[r].x;
                    ai[r] += wr[r] * vv[r].y;
                } else {
                    ar[r] += wr[r] * vv[r].x - wi[r] * vv[r].y;
                    ai[r] += wr[r] * vv[r].y + wi[r] * vv[r].x;
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const IDX i = i0 | ((IDX)r << 8);
            const double2 self = in[i];
            if (dtab) {     // tabulated diagonal part: three additive factor tables over 12-bit chunks of the index
                const double2 d = cadd(cadd(__ldg(dtab + (i & (IDX)0xfff)), __ldg(dtab + 4096 + ((i >> 12) & (IDX)0xfff))),
                                       __ldg(dtab + 8192 + (unsigned)((u64)i >> 24)));
                if (REAL) {
                    ar[r] += d.x * self.x;
                    ai[r] += d.x * self.y;
                } else {
                    ar[r] += d.x * self.x - d.y * self.y;
                    ai[r] += d.x * self.y + d.y * self.x;
                }
            }
            er += self.x * ar[r] + self.y * ai[r];     // conj(self) * acc
            ei += self.x * ai[r] - self.y * ar[r];
            if (MODE == 1) out[i] = make_double2(ar[r], ai[r]);
            if (MODE == 2) {
                const double2 o = out[i];
                out[i] = make_double2(o.x + ar[r], o.y + ai[r]);
            }
        }
    }
    const double sr = block_sum<256>(er, red);
    const double si = block_sum<256>(ei, red);
    // the last CTA to finish folds all per-CTA partials in slot order (deterministic) -- no separate finalize launch
    __shared__ unsigned is_last;
    if (threadIdx.x == 0) {
        partials[2 * blockIdx.x] = sr;
        partials[2 * blockIdx.x + 1] = si;
        __threadfence();
        is_last = (atomicAdd(counter, 1u) == gridDim.x - 1u) ? 1u : 0u;
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double r = 0.0, im = 0.0;
        for (int t = threadIdx.x; t < (int)gridDim.x; t += blockDim.x) {
            r += __ldcg(partials + 2 * t);
            im += __ldcg(partials + 2 * t + 1);
        }
        const double fr = block_sum<256>(r, red);
        const double fi = block_sum<256>(im, red);
        if (threadIdx.x == 0) {
            result[0] = fr;
            result[1] = fi;
            *counter = 0u;
        }
    }
}

// sum `count` (re,im) partial pairs in a fixed order -> result[0..1]; single block
__global__ void __launch_bounds__(256) k_finalize_pairs(const double *__restrict__ partials, int count,
                                                        double *__restrict__ result) {
    __shared__ double red[8];
    double r = 0.0, im = 0.0;
    for (int t = threadIdx.x; t < count; t += blockDim.x) {
        r += partials[2 * t];
        im += partials[2 * t + 1];
    }
    const double sr = block_sum<256>(r, red);
    const double si = block_sum<256>(im, red);
    if (threadIdx.x == 0) {
        result[0] = sr;
        result[1] = si;
    }
}

// ----------------------------------------------------------------------------------------------
// K3: pool gradients
// ----------------------------------------------------------------------------------------------
// grid = (chunks, entries).  Only the 2^(n-|fixmask|) connected index pairs of an entry are enumerated
// (bit-deposit of the free bits around the fixed pattern), never all 2^n.
__global__ void __launch_bounds__(128) k_pool(const PoolEntry *__restrict__ entries, int first_entry, int n,
                                              const double2 *__restrict__ psi, const double2 *__restrict__ lam,
                                              double *__restrict__ partials, const int *__restrict__ entry_ids, int e0,
                                              int e1, int row) {
    __shared__ PoolEntry e;
    __shared__ double red[4];
    const int eid = entry_ids ? entry_ids[first_entry + blockIdx.y] : first_entry + (int)blockIdx.y;
    if (eid < e0 || eid >= e1) return;          // outside the requested output range (uniform per block)
    {
        const int words = sizeof(PoolEntry) / 8;
        const u64 *src = reinterpret_cast<const u64 *>(entries + eid);
        for (int t = threadIdx.x; t < words; t += blockDim.x) reinterpret_cast<u64 *>(&e)[t] = __ldg(src + t);
        __syncthreads();
    }
    const u64 npairs = 1ull << (n - e.npos);
    const u64 stride = (u64)gridDim.x * blockDim.x;
    const double2 B = make_double2(e.br, e.bi);
    const double2 Bc = make_double2(e.br, -e.bi);
    double acc = 0.0;
    u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    // two pairs in flight per iteration
    for (; idx + stride < npairs; idx += 2 * stride) {
        const u64 i0 = deposit_zeros(idx, e.pos, e.npos) | e.fixval, j0 = i0 ^ e.x;
        const u64 i1 = deposit_zeros(idx + stride, e.pos, e.npos) | e.fixval, j1 = i1 ^ e.x;
        const double2 a0 = psi[i0], b0 = psi[j0], la0 = lam[i0], lb0 = lam[j0];
        const double2 a1 = psi[i1], b1 = psi[j1], la1 = lam[i1], lb1 = lam[j1];
        acc += sign_of(i0 & e.zeta) * (im_conj_mul(la0, cmul(B, b0)) + im_conj_mul(lb0, cmul(Bc, a0)));
        acc += sign_of(i1 & e.zeta) * (im_conj_mul(la1, cmul(B, b1)) + im_conj_mul(lb1, cmul(Bc, a1)));
    }
    for (; idx < npairs; idx += stride) {
        const u64 i0 = deposit_zeros(idx, e.pos, e.npos) | e.fixval, j0 = i0 ^ e.x;
        const double2 a0 = psi[i0], b0 = psi[j0], la0 = lam[i0], lb0 = lam[j0];
        acc += sign_of(i0 & e.zeta) * (im_conj_mul(la0, cmul(B, b0)) + im_conj_mul(lb0, cmul(Bc, a0)));
    }
    const double r = block_sum<128>(acc, red);
    if (threadIdx.x == 0) partials[(size_t)eid * row + blockIdx.x] = 2.0 * r;
}

// K3, 32-bit index variant for n <= 31 and patterns of at most 4 pinned bits (every pool the drivers build): the
// bit-deposit is four register mask insertions instead of a loop over a position array, and all index math is 32-bit.
__global__ void __launch_bounds__(128) k_pool32(const PoolEntry *__restrict__ entries, int first_entry, int n,
                                                const double2 *__restrict__ psi, const double2 *__restrict__ lam,
                                                double *__restrict__ partials, const int *__restrict__ entry_ids, int e0,
                                                int e1, int row) {
    __shared__ PoolEntry e;
    __shared__ double red[4];
    const int eid = entry_ids ? entry_ids[first_entry + blockIdx.y] : first_entry + (int)blockIdx.y;
    if (eid < e0 || eid >= e1) return;
    {
        const int words = sizeof(PoolEntry) / 8;
        const u64 *src = reinterpret_cast<const u64 *>(entries + eid);
        for (int t = threadIdx.x; t < words; t += blockDim.x) reinterpret_cast<u64 *>(&e)[t] = __ldg(src + t);
        __syncthreads();
    }
    const int npos = e.npos;
    const unsigned m0 = npos > 0 ? (1u << e.pos[0]) - 1u : 0xffffffffu, m1 = npos > 1 ? (1u << e.pos[1]) - 1u : 0xffffffffu,
                   m2 = npos > 2 ? (1u << e.pos[2]) - 1u : 0xffffffffu, m3 = npos > 3 ? (1u << e.pos[3]) - 1u : 0xffffffffu;
    const unsigned fixval = (unsigned)e.fixval, x = (unsigned)e.x, zeta = (unsigned)e.zeta;
    const unsigned npairs = 1u << (n - npos);
    const unsigned stride = gridDim.x * blockDim.x;
    const double2 B = make_double2(e.br, e.bi);
    const double2 Bc = make_double2(e.br, -e.bi);
    double acc = 0.0;
    auto place = [&](unsigned v) {
        v = ((v & ~m0) << 1) | (v & m0);
        v = ((v & ~m1) << 1) | (v & m1);
        v = ((v & ~m2) << 1) | (v & m2);
        v = ((v & ~m3) << 1) | (v & m3);
        return v | fixval;
    };
    unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    for (; idx + stride < npairs; idx += 2 * stride) {
        const unsigned i0 = place(idx), j0 = i0 ^ x, i1 = place(idx + stride), j1 = i1 ^ x;
        const double2 a0 = psi[i0], b0 = psi[j0], la0 = lam[i0], lb0 = lam[j0];
        const double2 a1 = psi[i1], b1 = psi[j1], la1 = lam[i1], lb1 = lam[j1];
        const double s0 = (__popc(i0 & zeta) & 1) ? -1.0 : 1.0, s1 = (__popc(i1 & zeta) & 1) ? -1.0 : 1.0;
        acc += s0 * (im_conj_mul(la0, cmul(B, b0)) + im_conj_mul(lb0, cmul(Bc, a0)));
        acc += s1 * (im_conj_mul(la1, cmul(B, b1)) + im_conj_mul(lb1, cmul(Bc, a1)));
    }
    for (; idx < npairs; idx += stride) {
        const unsigned i0 = place(idx), j0 = i0 ^ x;
        const double2 a0 = psi[i0], b0 = psi[j0], la0 = lam[i0], lb0 = lam[j0];
        const double s0 = (__popc(i0 & zeta) & 1) ? -1.0 : 1.0;
        acc += s0 * (im_conj_mul(la0, cmul(B, b0)) + im_conj_mul(lb0, cmul(Bc, a0)));
    }
    const double r = block_sum<128>(acc, red);
    if (threadIdx.x == 0) partials[(size_t)eid * row + blockIdx.x] = 2.0 * r;
}

// K3 on shared-memory tiles: blockIdx.y = pass (a set of T index bits), blockIdx.x strides over the 2^(n-T) tiles.
// psi and lambda tiles are loaded once and every entry of the pass is evaluated from them, one warp per entry
// (its lanes cover the entry's in-tile pairs), so a gradient costs shared-memory instead of L2/HBM traffic.
// Per-entry partials: lane sums -> warp shuffle -> shared accumulator owned by that warp -> one store per (entry, CTA).
__global__ void __launch_bounds__(512) k_pool_tile(const PoolPass *__restrict__ passes, const PoolTileRec *__restrict__ recs,
                                                   int n, int chunks, const double2 *__restrict__ psi,
                                                   const double2 *__restrict__ lam, double *__restrict__ partials, int e0,
                                                   int e1) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ PoolTileRec rec[FH_POOL_PASS_MAX_RECS];
    __shared__ double accE[FH_POOL_PASS_MAX_RECS];
    __shared__ unsigned slo[64], shi[128];
    __shared__ PoolPass pass;
    if (threadIdx.x < sizeof(PoolPass) / 4)
        reinterpret_cast<unsigned *>(&pass)[threadIdx.x] = reinterpret_cast<const unsigned *>(passes + blockIdx.y)[threadIdx.x];
    __syncthreads();
    const int T = pass.nbits, nrec = pass.nrec;
    const unsigned L = 1u << T;
    double2 *bufp = reinterpret_cast<double2 *>(smem_raw);
    double2 *bufl = bufp + L;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(recs + pass.first_rec);
        uint4 *dst = reinterpret_cast<uint4 *>(rec);
        const int chunks16 = nrec * (int)(sizeof(PoolTileRec) / 16);
        for (int c = threadIdx.x; c < chunks16; c += blockDim.x) dst[c] = __ldg(src + c);
    }
    for (unsigned v = threadIdx.x; v < 64u; v += blockDim.x) {
        unsigned g = 0;
        for (int b = 0; b < 6 && b < T; ++b) g |= ((v >> b) & 1u) << pass.bits[b];
        slo[v] = g;
    }
    for (unsigned v = threadIdx.x; v < 128u; v += blockDim.x) {
        unsigned g = 0;
        for (int b = 6; b < T; ++b) g |= ((v >> (b - 6)) & 1u) << pass.bits[b];
        shi[v] = g;
    }
    for (int o = threadIdx.x; o < nrec; o += blockDim.x) accE[o] = 0.0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const u64 ntiles = 1ull << (n - T);
    for (u64 t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const unsigned base = (unsigned)deposit_zeros(t, pass.bits, T);
        for (unsigned l0 = threadIdx.x; l0 < L; l0 += 2 * blockDim.x) {
            unsigned gg[2];
            double2 va[2], vb[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const unsigned l = l0 + u * blockDim.x;
                if (l < L) {
                    gg[u] = base | slo[l & 63u] | shi[l >> 6];
                    va[u] = psi[gg[u]];
                    vb[u] = lam[gg[u]];
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const unsigned l = l0 + u * blockDim.x;
                if (l < L) {
                    const unsigned sl = tile_slot(l);
                    bufp[sl] = va[u];
                    bufl[sl] = vb[u];
                }
            }
        }
        __syncthreads();
        for (int o = warp; o < nrec; o += nwarps) {
            const uint4 *rp = reinterpret_cast<const uint4 *>(&rec[o]);
            const uint4 q0 = rp[0], q1 = rp[1], q2 = rp[2];
            // q0 = {fixmask_out, fixval_out, zeta, xlocal}; q1 = {lfixval, nlfix, entry, -}; q2 = lowmask[4]
            if ((int)q1.z < e0 || (int)q1.z >= e1) continue;
            if ((base & q0.x) != q0.y) continue;
            const double2 B = reinterpret_cast<const double2 *>(&rec[o])[3];
            const double2 Bc = make_double2(B.x, -B.y);
            const unsigned npairs = L >> q1.y;
            double acc = 0.0;
            for (unsigned k = lane; k < npairs; k += 32) {
                unsigned il = k;
                il = ((il & ~q2.x) << 1) | (il & q2.x);
                il = ((il & ~q2.y) << 1) | (il & q2.y);
                il = ((il & ~q2.z) << 1) | (il & q2.z);
                il = ((il & ~q2.w) << 1) | (il & q2.w);
                il |= q1.x;
                const unsigned gi = base | slo[il & 63u] | shi[il >> 6];
                const unsigned si = tile_slot(il), sj = tile_slot(il ^ q0.w);
                const double2 a = bufp[si], b = bufp[sj], la = bufl[si], lb = bufl[sj];
                const double sg = (__popc(gi & q0.z) & 1) ? -1.0 : 1.0;
                acc += sg * (im_conj_mul(la, cmul(B, b)) + im_conj_mul(lb, cmul(Bc, a)));
            }
            acc = warp_sum(acc);
            if (lane == 0) accE[o] += acc;
        }
        __syncthreads();
    }
    for (int o = threadIdx.x; o < nrec; o += blockDim.x) {
        const int entry = rec[o].entry;
        if (entry >= e0 && entry < e1) partials[(size_t)entry * chunks + blockIdx.x] = 2.0 * accE[o];
    }
}

// one warp per output: fixed-order sum over its entries' chunk partials
__global__ void __launch_bounds__(32) k_pool_finalize(const double *__restrict__ partials,
                                                      const int *__restrict__ out_first, int chunks, int first_out,
                                                      double *__restrict__ out) {
    const int o = first_out + blockIdx.x;
    const int lo = out_first[o] * chunks, hi = out_first[o + 1] * chunks;
    double acc = 0.0;
    for (int t = lo + threadIdx.x; t < hi; t += 32) acc += partials[t];
    acc = warp_sum(acc);
    if (threadIdx.x == 0) out[o] = acc;
}

// generic segmented sum: out[s] = sum partials[first[s] .. first[s+1])
__global__ void __launch_bounds__(32) k_sum_segments(const double *__restrict__ partials, const int *__restrict__ first,
                                                     double *__restrict__ out) {
    const int s = blockIdx.x;
    double acc = 0.0;
    for (int t = first[s] + threadIdx.x; t < first[s + 1]; t += 32) acc += partials[t];
    acc = warp_sum(acc);
    if (threadIdx.x == 0) out[s] = acc;
}

// ----------------------------------------------------------------------------------------------
// K4 / misc vector kernels
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_inner(const double2 *__restrict__ a, const double2 *__restrict__ b, u64 dim,
                                               double *__restrict__ partials) {
    __shared__ double red[8];
    const u64 stride = (u64)gridDim.x * blockDim.x;
    double r = 0.0, im = 0.0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += stride) {
        const double2 x = a[i], y = b[i];
        r += x.x * y.x + x.y * y.y;
        im += x.x * y.y - x.y * y.x;
    }
    const double sr = block_sum<256>(r, red);
    const double si = block_sum<256>(im, red);
    if (threadIdx.x == 0) {
        partials[2 * blockIdx.x] = sr;
        partials[2 * blockIdx.x + 1] = si;
    }
}

__global__ void __launch_bounds__(256) k_set_basis(double2 *psi, u64 dim, u64 index) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += stride)
        psi[i] = make_double2(i == index ? 1.0 : 0.0, 0.0);
}

__global__ void __launch_bounds__(256) k_flush(double4 *buf, size_t count) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
        buf[i] = make_double4(0.0, 0.0, 0.0, 0.0);
}

__global__ void __launch_bounds__(256) k_axpby(double2 *__restrict__ y, double a, const double2 *__restrict__ x,
                                               double b, u64 dim) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += stride) {
        const double2 xv = x[i], yv = y[i];
        y[i] = make_double2(a * xv.x + b * yv.x, a * xv.y + b * yv.y);
    }
}

__global__ void __launch_bounds__(256) k_caxpy(double2 *__restrict__ y, double ar, double ai,
                                               const double2 *__restrict__ x, u64 dim) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    const double2 a = make_double2(ar, ai);
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += stride) y[i] = cadd(y[i], cmul(a, x[i]));
}

__global__ void __launch_bounds__(256) k_lanczos_update(double2 *__restrict__ w, const double2 *__restrict__ v,
                                                        const double2 *__restrict__ vprev, double alpha, double beta,
                                                        u64 dim) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += stride) {
        const double2 wv = w[i], a = v[i], b = vprev[i];
        w[i] = make_double2(wv.x - alpha * a.x - beta * b.x, wv.y - alpha * a.y - beta * b.y);
    }
}

__global__ void __launch_bounds__(256) k_scale(double2 *__restrict__ y, double a, u64 dim) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += stride) y[i] = cscale(y[i], a);
}

__device__ __forceinline__ u64 mix64(u64 z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

// Gaussian start vector supported on the (n_up, n_dn) sector (even wires = up); n_up < 0: full space
__global__ void __launch_bounds__(256) k_sector_random(double2 *v, int n, int n_up, int n_dn, u64 seed) {
    const u64 dim = 1ull << n;
    u64 upmask = 0;
    for (int q = 0; q < n; q += 2) upmask |= 1ull << (n - 1 - q);
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += stride) {
        double val = 0.0;
        const bool in_sector = n_up < 0 || (__popcll(i & upmask) == n_up && __popcll(i & ~upmask) == n_dn);
        if (in_sector) {
            const u64 h1 = mix64(i * 2 + seed * 0x100000001b3ull), h2 = mix64(h1 + 1);
            const double u1 = ((h1 >> 11) + 1.0) * (1.0 / 9007199254740993.0);
            const double u2 = (h2 >> 11) * (1.0 / 9007199254740992.0);
            val = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
        }
        v[i] = make_double2(val, 0.0);
    }
}

// dst[i] = src[i with bit a[k] and bit b[k] exchanged, k < npairs]: the local half of a global<->local qubit
// swap of a sharded state (the other half is an all-to-all over the top local bits).  Writes are consecutive;
// reads stay in >= 2^min(a,b)-amplitude runs.
struct SwapPairs {
    int n;
    unsigned char a[8], b[8];
};

__global__ void __launch_bounds__(256) k_swap_bits(const double2 *__restrict__ src, double2 *__restrict__ dst, u64 dim,
                                                   const SwapPairs sp) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += stride) {
        u64 j = i;
#pragma unroll 1
        for (int k = 0; k < sp.n; ++k) {
            const u64 t = ((j >> sp.a[k]) ^ (j >> sp.b[k])) & 1ull;
            j ^= (t << sp.a[k]) | (t << sp.b[k]);
        }
        dst[i] = src[j];
    }
}

// the same permutation restricted to the destination range [first, first + count): one chunk of a pipelined exchange
__global__ void __launch_bounds__(256) k_swap_bits_range(const double2 *__restrict__ src, double2 *__restrict__ dst, u64 first,
                                                         u64 count, const SwapPairs sp) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride) {
        const u64 i = first + t;
        u64 j = i;
#pragma unroll 1
        for (int k = 0; k < sp.n; ++k) {
            const u64 q = ((j >> sp.a[k]) ^ (j >> sp.b[k])) & 1ull;
            j ^= (q << sp.a[k]) | (q << sp.b[k]);
        }
        dst[i] = src[j];
    }
}

void launch_swap_bits_range(cudaStream_t s, int sm, const double2 *src, double2 *dst, int n, int npairs, const int *a,
                            const int *b, u64 first, u64 count) {
    SwapPairs sp;
    sp.n = npairs;
    for (int k = 0; k < npairs; ++k) {
        sp.a[k] = (unsigned char)a[k];
        sp.b[k] = (unsigned char)b[k];
    }
    (void)n;
    u64 blocks = (count + 255) / 256;
    if (blocks > (u64)sm * 32) blocks = (u64)sm * 32;
    if (blocks < 1) blocks = 1;
    ++g_fh_launch_count;
    k_swap_bits_range<<<(int)blocks, 256, 0, s>>>(src, dst, first, count, sp);
}

void launch_swap_bits(cudaStream_t s, int sm, const double2 *src, double2 *dst, int n, int npairs, const int *a,
                      const int *b) {
    SwapPairs sp;
    sp.n = npairs;
    for (int k = 0; k < npairs; ++k) {
        sp.a[k] = (unsigned char)a[k];
        sp.b[k] = (unsigned char)b[k];
    }
    const u64 dim = 1ull << n;
    u64 blocks = (dim + 255) / 256;
    if (blocks > (u64)sm * 32) blocks = (u64)sm * 32;
    ++g_fh_launch_count;
    k_swap_bits<<<(int)blocks, 256, 0, s>>>(src, dst, dim, sp);
}

// ----------------------------------------------------------------------------------------------
// launch wrappers
// ----------------------------------------------------------------------------------------------
static inline int grid_for(u64 work_items, int threads, int per_thread, int sm, int max_blocks) {
    u64 blocks = (work_items + (u64)threads * per_thread - 1) / ((u64)threads * per_thread);
    if (blocks < 1) blocks = 1;
    if (blocks > (u64)max_blocks) blocks = max_blocks;
    (void)sm;
    return (int)blocks;
}

void launch_pair(cudaStream_t s, int sm, double2 *psi, const PairOp *d_op, int n, int nfix, int dagger) {
    const u64 npairs = 1ull << (n - nfix);
    if (npairs >= (u64)sm * 128 * 16) {
        const int grid = grid_for(npairs, 128, 4, sm, sm * 32);
        ++g_fh_launch_count; k_pair<4><<<grid, 128, 0, s>>>(psi, d_op, npairs, dagger);
    } else {
        const int grid = grid_for(npairs, 128, 1, sm, sm * 32);
        ++g_fh_launch_count; k_pair<1><<<grid, 128, 0, s>>>(psi, d_op, npairs, dagger);
    }
}

void launch_pair_adjoint(cudaStream_t s, int sm, double2 *psi, double2 *lam, const PairOp *d_op, int n, int nfix,
                         double *d_partials, int max_blocks, int *blocks_used) {
    const u64 npairs = 1ull << (n - nfix);
    int grid = grid_for(npairs, 128, npairs >= (u64)sm * 128 * 8 ? 2 : 1, sm, max_blocks);
    ++g_fh_launch_count; k_pair_adjoint<<<grid, 128, 0, s>>>(psi, lam, d_op, npairs, d_partials);
    *blocks_used = grid;
}

// Diagonal op through factor tables (large states): the phase of index i is the product of three table entries
// (index bits 0..11, 12..23, 24..) times the few terms whose z-mask straddles two chunks.  k_diag_build fills the
// tables and compacts the straddling terms; k_diag_tab streams the state once with three L1-resident lookups.
#define DIAG_TAB_A 4096
#define DIAG_TAB_B 4096
#define DIAG_TAB_C 1024
#define DIAG_TAB_ENTRIES (DIAG_TAB_A + DIAG_TAB_B + DIAG_TAB_C)
#define DIAG_MAX_CROSS 512
struct DiagCross {
    u64 z;
    double c, s;
    double pad;
};

__device__ __forceinline__ int diag_chunk_of(u64 z) {     // 0 / 1 / 2: z inside that chunk; -1: straddles
    const u64 ma = 0xfffull, mb = 0xfffull << 12;
    if ((z & ~ma) == 0) return 0;
    if ((z & ~mb) == 0) return 1;
    if ((z & (ma | mb)) == 0) return 2;
    return -1;
}

__global__ void __launch_bounds__(256) k_diag_build(const DiagTerm *__restrict__ terms, int nterms, int dagger,
                                                    double2 *__restrict__ tab, int *__restrict__ ncross,
                                                    DiagCross *__restrict__ cross) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < DIAG_TAB_ENTRIES) {
        const int chunk = v < DIAG_TAB_A ? 0 : (v < DIAG_TAB_A + DIAG_TAB_B ? 1 : 2);
        const u64 idx = (u64)(v - (chunk == 0 ? 0 : (chunk == 1 ? DIAG_TAB_A : DIAG_TAB_A + DIAG_TAB_B)));
        const u64 gl = idx << (12 * chunk);
        double tot = 0.0;
        for (int m = 0; m < nterms; ++m) {
            const u64 z = terms[m].z;
            if (diag_chunk_of(z) == chunk) tot += sign_of(gl & z) * terms[m].angle;
        }
        double sn, cs;
        sincos(tot, &sn, &cs);
        tab[v] = make_double2(cs, dagger ? sn : -sn);
    }
    if (v == 0) {
        int k = 0;
        for (int m = 0; m < nterms; ++m)
            if (diag_chunk_of(terms[m].z) < 0) {
                cross[k].z = terms[m].z;
                cross[k].c = terms[m].c;
                cross[k].s = dagger ? -terms[m].s : terms[m].s;
                ++k;
            }
        *ncross = k;
    }
}

__global__ void __launch_bounds__(256) k_diag_tab(double2 *__restrict__ psi, const double2 *__restrict__ tab,
                                                  const int *__restrict__ ncross, const DiagCross *__restrict__ cross,
                                                  u64 dim) {
    const int nc = *ncross;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i0 = ((u64)blockIdx.x * blockDim.x + threadIdx.x); i0 < dim; i0 += 2 * stride) {
        const u64 i1 = i0 + stride;
        const bool two = i1 < dim;
        const double2 a0 = psi[i0];
        const double2 a1 = two ? psi[i1] : make_double2(0.0, 0.0);
        double2 f0 = cmul(cmul(__ldg(tab + (i0 & 0xfffull)), __ldg(tab + DIAG_TAB_A + ((i0 >> 12) & 0xfffull))),
                          __ldg(tab + DIAG_TAB_A + DIAG_TAB_B + (i0 >> 24)));
        double2 f1 = make_double2(1.0, 0.0);
        if (two)
            f1 = cmul(cmul(__ldg(tab + (i1 & 0xfffull)), __ldg(tab + DIAG_TAB_A + ((i1 >> 12) & 0xfffull))),
                      __ldg(tab + DIAG_TAB_A + DIAG_TAB_B + (i1 >> 24)));
        for (int m = 0; m < nc; ++m) {
            const u64 z = cross[m].z;
            const double c = cross[m].c, sn = cross[m].s;
            f0 = cmul(f0, make_double2(c, -sign_of(i0 & z) * sn));
            if (two) f1 = cmul(f1, make_double2(c, -sign_of(i1 & z) * sn));
        }
        psi[i0] = cmul(f0, a0);
        if (two) psi[i1] = cmul(f1, a1);
    }
}

void launch_diag(cudaStream_t s, int sm, double2 *psi, const DiagTerm *d_terms, int nterms, int n, int dagger,
                 void *scratch) {
    const u64 dim = 1ull << n;
    if (scratch && n >= 20 && n <= 34 && nterms > 4 && nterms <= DIAG_MAX_CROSS) {
        double2 *tab = reinterpret_cast<double2 *>(scratch);
        int *ncross = reinterpret_cast<int *>(tab + DIAG_TAB_ENTRIES);
        DiagCross *cross = reinterpret_cast<DiagCross *>(tab + DIAG_TAB_ENTRIES + 1);
        ++g_fh_launch_count;
        k_diag_build<<<(DIAG_TAB_ENTRIES + 255) / 256, 256, 0, s>>>(d_terms, nterms, dagger, tab, ncross, cross);
        const int grid = grid_for(dim, 256, 4, sm, sm * 16);
        ++g_fh_launch_count;
        k_diag_tab<<<grid, 256, 0, s>>>(psi, tab, ncross, cross, dim);
        return;
    }
    for (int off = 0; off < nterms; off += DIAG_SMEM_TERMS) {
        const int cnt = nterms - off < DIAG_SMEM_TERMS ? nterms - off : DIAG_SMEM_TERMS;
        const int grid = grid_for(dim, 256, dim >= (u64)sm * 256 * 8 ? 2 : 1, sm, sm * 16);
        ++g_fh_launch_count; k_diag<<<grid, 256, 0, s>>>(psi, d_terms + off, cnt, dim, dagger);
    }
}

size_t fh_diag_scratch_bytes() { return sizeof(double2) * (DIAG_TAB_ENTRIES + 1) + sizeof(DiagCross) * DIAG_MAX_CROSS; }

void launch_diag_adjoint(cudaStream_t s, int sm, double2 *psi, double2 *lam, const DiagTerm *d_terms, int nterms, int n,
                         double *d_partials, int max_blocks, int *blocks_used) {
    const u64 dim = 1ull << n;
    int grid = grid_for(dim, 256, 1, sm, max_blocks);
    ++g_fh_launch_count; k_diag_adjoint<<<grid, 256, 0, s>>>(psi, lam, d_terms, nterms, dim, d_partials);
    *blocks_used = grid;
}

// register-staged (LDG/STG) tile kernel: fallback of launch_tile (tile_tma.cu) when the TMA path does not apply
void launch_tile_ldg(cudaStream_t s, double2 *psi, const TileLaunch &tl, const TileRec *d_recs, const TileTerm *d_terms,
                     int n) {
    const int nbits = tl.nbits;
    const size_t smem = ((size_t)1 << nbits) * (sizeof(double2) + sizeof(unsigned int));
    u64 ntiles = 1ull << (n - nbits);
    const int grid = (int)(ntiles > 148ull * 16 ? 148ull * 16 : ntiles);
    // one thread per index pair of the tile (at most 512, at least two warps)
    int threads = nbits >= 1 ? (1 << (nbits - 1)) : 1;
    if (threads > 512) threads = 512;
    if (threads < 64) threads = 64;
    ++g_fh_launch_count;
    // Programmatic dependent launch: the launch may begin while the previous kernel on the stream is finishing;
    // k_tile<true> sets up its records and tables and then waits (griddepcontrol.wait) before it touches the state.
    // Captured into the evaluation graph as a programmatic edge (18-qubit screening step 0.217 -> 0.206 ms).  On by
    // default inside fh_program_evaluate; FHSIM_PDL=1 turns it on for every tile launch, FHSIM_NO_PDL=1 off.
    static const bool pdl_all = getenv("FHSIM_PDL") != nullptr, pdl_off = getenv("FHSIM_NO_PDL") != nullptr;
    const bool pdl = !pdl_off && (pdl_all || g_fh_tile_pdl_scope > 0);
    if (pdl) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3((unsigned)threads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, k_tile<true>, psi, tl, d_recs, d_terms, n);
    } else {
        k_tile<false><<<grid, threads, smem, s>>>(psi, tl, d_recs, d_terms, n);
    }
}

static int g_tile_multi_blocks_per_sm = -1;      // occupancy query result (same for every B200 of the box)

// returns 0 when the cooperative launch cannot be used (caller falls back to one launch per run)
int launch_tile_multi(cudaStream_t s, int sm, double2 *psi, const TileLaunch *d_tls, int nl, int max_bits, int min_bits,
                      const TileRec *d_recs, const TileTerm *d_terms, int n) {
    const size_t smem = ((size_t)1 << max_bits) * (sizeof(double2) + sizeof(unsigned int));
    int threads = max_bits >= 1 ? (1 << (max_bits - 1)) : 1;
    if (threads > 512) threads = 512;
    if (threads < 64) threads = 64;
    if (g_tile_multi_blocks_per_sm < 0) {
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_tile_multi, 512, ((size_t)1 << 12) * 20) != cudaSuccess) nb = 0;
        g_tile_multi_blocks_per_sm = nb;
    }
    int per_sm = g_tile_multi_blocks_per_sm;
    if (max_bits > 12) per_sm = per_sm > 1 ? 1 : per_sm;
    if (per_sm <= 0) return 0;
    const u64 ntiles = 1ull << (n - min_bits);
    u64 grid = (u64)sm * per_sm;
    if (grid > ntiles) grid = ntiles;
    void *args[] = {(void *)&psi, (void *)&d_tls, (void *)&nl, (void *)&d_recs, (void *)&d_terms, (void *)&n};
    ++g_fh_launch_count;
    if (cudaLaunchCooperativeKernel((const void *)k_tile_multi, dim3((unsigned)grid), dim3(threads), args, smem, s) !=
        cudaSuccess) {
        cudaGetLastError();
        --g_fh_launch_count;
        return 0;
    }
    return 1;
}

void launch_tile_adjoint(cudaStream_t s, double2 *psi, double2 *lam, const TileLaunch &tl, const TileRec *d_recs,
                         const TileTerm *d_terms, int n, double *d_gpart, int seg_base) {
    const int nbits = tl.nbits;
    const size_t smem = ((size_t)1 << nbits) * (2 * sizeof(double2) + sizeof(unsigned int));
    const u64 ntiles = 1ull << (n - nbits);
    const int grid = (int)(ntiles > (u64)FH_GRAD_BLOCKS ? (u64)FH_GRAD_BLOCKS : ntiles);
    int threads = nbits >= 1 ? (1 << (nbits - 1)) : 1;
    if (threads > 512) threads = 512;
    if (threads < 64) threads = 64;
    ++g_fh_launch_count;
    k_tile_adjoint<<<grid, threads, smem, s>>>(psi, lam, tl, d_recs, d_terms, n, d_gpart, seg_base);
}

void launch_apply_table(cudaStream_t s, int sm, const fh_table *tab, const double2 *in, double2 *out, int mode,
                        double *d_partials, double *d_result) {
    const u64 dim = 1ull << tab->n;
    const int ngroups = (int)tab->groups.size(), nclasses = (int)tab->classes.size(), nvals = (int)tab->vals.size();
    const size_t need = (size_t)ngroups * sizeof(TabGroup) + (size_t)nclasses * sizeof(TabClass) +
                        (size_t)nvals * sizeof(double2);
    const int use_smem = need <= 40 * 1024;
    const size_t smem = use_smem ? need : 0;
    if (!out) mode = 0;
    // n >= 22 and a table whose x-mask groups are covered by <= 3 sets of 12 index bits: shared-memory tile passes
    if (tab->tiles && launch_apply_table_tiles(s, sm, tab, in, out, mode, d_partials, d_result)) return;
    int grid;
    if (tab->n >= 11 && use_smem) {
        const int rl = tab->n >= 22 ? 3 : 2;          // 8 outputs per thread once there is parallelism to spare
        const unsigned nblk = (unsigned)(dim >> (8 + rl));
        grid = nblk < (unsigned)(sm * 8) ? (int)nblk : sm * 8;
        // Window traversal (see the kernel): measured on the 3x4 lattice it moves K2 by -4 % (tools/run_k2.py) to +17 %
        // (bench.py hbm_regime) -- K2 is bound by its instruction stream, not by the DRAM gathers -- so the linear order
        // stays the default and the windows are opt-in: FHSIM_K2_SLOW_BITS=n-21 gives 32-MiB windows
        int slow_bits = 0;
        if (const char *env = getenv("FHSIM_K2_SLOW_BITS")) slow_bits = atoi(env);      // tuning / tests (read per call)
        if (slow_bits < 0 || (1u << slow_bits) >= nblk) slow_bits = 0;
#define LAUNCH_TAB4(I, R, M, RLV)                                                                                   \
    do {                                                                                                            \
        ++g_fh_launch_count;                                                                                        \
        k_apply_table4<I, R, M, RLV><<<grid, 256, smem, s>>>(tab->d_groups, ngroups, tab->d_classes, nclasses,      \
                                                             tab->d_vals, nvals, in, out, nblk, d_partials,         \
                                                             tab->d_diag, tab->ctx->d_counter, d_result, slow_bits); \
    } while (0)
#define LAUNCH_TAB4_M(I, R, RLV)                                                                                    \
    do {                                                                                                            \
        if (mode == 0) LAUNCH_TAB4(I, R, 0, RLV); else if (mode == 1) LAUNCH_TAB4(I, R, 1, RLV); else LAUNCH_TAB4(I, R, 2, RLV); \
    } while (0)
#define LAUNCH_TAB4_I(I)                                                                                            \
    do {                                                                                                            \
        if (tab->all_real) { if (rl == 3) LAUNCH_TAB4_M(I, true, 3); else LAUNCH_TAB4_M(I, true, 2); }              \
        else { if (rl == 3) LAUNCH_TAB4_M(I, false, 3); else LAUNCH_TAB4_M(I, false, 2); }                          \
    } while (0)
        if (tab->n <= 31) LAUNCH_TAB4_I(unsigned); else LAUNCH_TAB4_I(u64);
#undef LAUNCH_TAB4_I
#undef LAUNCH_TAB4_M
#undef LAUNCH_TAB4
        return;
    }
    grid = grid_for(dim, 256, 1, sm, FH_MAX_PARTIALS);
    if (grid > sm * 8 && dim > (u64)sm * 8 * 256) grid = sm * 8;
#define LAUNCH_TAB(R, M)                                                                                           \
    do {                                                                                                           \
        ++g_fh_launch_count;                                                                                       \
        k_apply_table<R, M><<<grid, 256, smem, s>>>(tab->d_groups, ngroups, tab->d_classes, nclasses, tab->d_vals, \
                                                    nvals, use_smem, in, out, dim, d_partials, tab->d_diag);       \
    } while (0)
    if (tab->all_real) {
        if (mode == 0) LAUNCH_TAB(true, 0); else if (mode == 1) LAUNCH_TAB(true, 1); else LAUNCH_TAB(true, 2);
    } else {
        if (mode == 0) LAUNCH_TAB(false, 0); else if (mode == 1) LAUNCH_TAB(false, 1); else LAUNCH_TAB(false, 2);
    }
#undef LAUNCH_TAB
    ++g_fh_launch_count; k_finalize_pairs<<<1, 256, 0, s>>>(d_partials, grid, d_result);
}

void launch_pool(cudaStream_t s, const PoolEntry *entries, int first_entry, int n_entries, int chunks, int n,
                 const double2 *psi, const double2 *lam, double *d_partials, const int *entry_ids, int e0, int e1,
                 int row, int narrow) {
    if (n_entries <= 0) return;
    if (row <= 0) row = chunks;
    // gridDim.y is limited to 65535
    for (int off = 0; off < n_entries; off += 32768) {
        const int cnt = n_entries - off < 32768 ? n_entries - off : 32768;
        dim3 grid(chunks, cnt);
        ++g_fh_launch_count;
        if (narrow && n <= 31)
            k_pool32<<<grid, 128, 0, s>>>(entries, first_entry + off, n, psi, lam, d_partials, entry_ids, e0, e1, row);
        else
            k_pool<<<grid, 128, 0, s>>>(entries, first_entry + off, n, psi, lam, d_partials, entry_ids, e0, e1, row);
    }
}

void launch_pool_tiles(cudaStream_t s, const PoolPass *d_passes, int npasses, const PoolTileRec *d_recs, int tile_bits,
                       int grid_x, int chunks, int n, const double2 *psi, const double2 *lam, double *d_partials, int e0,
                       int e1) {
    if (npasses <= 0) return;
    const size_t smem = ((size_t)1 << tile_bits) * 2 * sizeof(double2);
    int threads = 1 << (tile_bits > 3 ? tile_bits - 3 : 0);
    if (threads > 512) threads = 512;
    if (threads < 64) threads = 64;
    for (int off = 0; off < npasses; off += 32768) {
        const int cnt = npasses - off < 32768 ? npasses - off : 32768;
        dim3 grid(grid_x, cnt);
        ++g_fh_launch_count;
        k_pool_tile<<<grid, threads, smem, s>>>(d_passes + off, d_recs, n, chunks, psi, lam, d_partials, e0, e1);
    }
}

void launch_pool_finalize(cudaStream_t s, const double *d_partials, const int *d_out_first, int chunks, int first_out,
                          int count, double *d_out) {
    if (count <= 0) return;
    ++g_fh_launch_count; k_pool_finalize<<<count, 32, 0, s>>>(d_partials, d_out_first, chunks, first_out, d_out);
}

void launch_inner(cudaStream_t s, int sm, const double2 *a, const double2 *b, u64 dim, double *d_partials,
                  double *d_result) {
    int grid = grid_for(dim, 256, 4, sm, sm * 8);
    ++g_fh_launch_count; k_inner<<<grid, 256, 0, s>>>(a, b, dim, d_partials);
    ++g_fh_launch_count; k_finalize_pairs<<<1, 256, 0, s>>>(d_partials, grid, d_result);
}

void launch_sum_segments(cudaStream_t s, const double *d_partials, const int *d_first, int nseg, double *d_out) {
    if (nseg <= 0) return;
    ++g_fh_launch_count; k_sum_segments<<<nseg, 32, 0, s>>>(d_partials, d_first, d_out);
}

// out[s] = sum_{t < stride} partials[s*stride + t]   (fixed order; unused slots are pre-zeroed)
__global__ void __launch_bounds__(32) k_sum_strided(const double *__restrict__ partials, int stride,
                                                    double *__restrict__ out) {
    const double *seg = partials + (size_t)blockIdx.x * stride;
    double acc = 0.0;
    for (int t = threadIdx.x; t < stride; t += 32) acc += seg[t];
    acc = warp_sum(acc);
    if (threadIdx.x == 0) out[blockIdx.x] = acc;
}

void launch_sum_strided(cudaStream_t s, const double *d_partials, int stride, int nseg, double *d_out) {
    if (nseg <= 0) return;
    ++g_fh_launch_count; k_sum_strided<<<nseg, 32, 0, s>>>(d_partials, stride, d_out);
}

void launch_set_basis(cudaStream_t s, double2 *psi, u64 dim, u64 index) {
    int grid = (int)((dim + 255) / 256 > 148 * 16 ? 148 * 16 : (dim + 255) / 256);
    ++g_fh_launch_count; k_set_basis<<<grid, 256, 0, s>>>(psi, dim, index);
}

void launch_flush(cudaStream_t s, void *buf, size_t bytes) {
    ++g_fh_launch_count; k_flush<<<148 * 8, 256, 0, s>>>(reinterpret_cast<double4 *>(buf), bytes / sizeof(double4));
}

static inline int vec_grid(u64 dim, int sm) { return grid_for(dim, 256, 2, sm, sm * 16); }

void launch_axpby(cudaStream_t s, int sm, double2 *y, double a, const double2 *x, double b, u64 dim) {
    ++g_fh_launch_count; k_axpby<<<vec_grid(dim, sm), 256, 0, s>>>(y, a, x, b, dim);
}
void launch_caxpy(cudaStream_t s, int sm, double2 *y, double ar, double ai, const double2 *x, u64 dim) {
    ++g_fh_launch_count; k_caxpy<<<vec_grid(dim, sm), 256, 0, s>>>(y, ar, ai, x, dim);
}
void launch_lanczos_update(cudaStream_t s, int sm, double2 *w, const double2 *v, const double2 *vprev, double alpha,
                           double beta, u64 dim) {
    ++g_fh_launch_count; k_lanczos_update<<<vec_grid(dim, sm), 256, 0, s>>>(w, v, vprev, alpha, beta, dim);
}
void launch_scale(cudaStream_t s, int sm, double2 *y, double a, u64 dim) {
    ++g_fh_launch_count; k_scale<<<vec_grid(dim, sm), 256, 0, s>>>(y, a, dim);
}
void launch_sector_random(cudaStream_t s, int sm, double2 *v, int n, int n_up, int n_dn, u64 seed) {
    ++g_fh_launch_count; k_sector_random<<<vec_grid(1ull << n, sm), 256, 0, s>>>(v, n, n_up, n_dn, seed);
}


// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: one process may drive several GPUs
// (fhsim.backend.default_context(device)), so every context sets it for its own device.
int fh_tile_tma_init_device();

int fh_kernels_init_device() {
    FH_TRY(fh_tile_tma_init_device());
    FH_CUDA(cudaFuncSetAttribute(k_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 180 * 1024));
    FH_CUDA(cudaFuncSetAttribute(k_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 180 * 1024));
    FH_CUDA(cudaFuncSetAttribute(k_tile_multi, cudaFuncAttributeMaxDynamicSharedMemorySize, 180 * 1024));
    FH_CUDA(cudaFuncSetAttribute(k_tile_adjoint, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    FH_CUDA(cudaFuncSetAttribute(k_pool_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024));
    return FH_OK;
}
