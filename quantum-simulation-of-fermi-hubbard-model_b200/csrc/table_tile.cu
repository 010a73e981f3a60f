// K2 in shared-memory tile passes (n >= 22): out = H in / out += H in fused with <in|H|in>.
//
// The gather kernel (k_apply_table4) reads one partner amplitude per (amplitude, live x-mask group) through L1/L2: on the
// 3x4 lattice that is ~24 x 16 B per amplitude, 7 GB per launch at 24 qubits, of which 2 GB miss to DRAM (profiles/r01_*).
// Here the x-mask groups are covered by a few PASSES; a pass fixes 12 index bits that contain the x-masks of its groups
// (for the Hubbard Hamiltonian: all up-orbital bits, then all down-orbital bits -- every hopping term stays inside one
// species), stages the 2^12 amplitudes that differ only in those bits in shared memory once, and serves every partner of
// every group of the pass from there.  Traffic per pass: 16 B read of `in` + 16 B write (first pass) or 32 B read-modify-
// write of `out` per amplitude, i.e. 80 B per amplitude for two passes against ~416 B through L2 before.
// A thread owns the 2^(12 - TT_LOW) amplitudes that differ in the top tile bits: group decode (partner slot, x-bit pattern,
// sign parity) is done once for all of them; per amplitude only the pattern / sign deltas of those top bits are applied
// (4-bit and 1-bit table lookups in registers).  Tiles t and t+1 differ in the lowest non-tile bit, so when that is index
// bit 0 (pass over the odd bits) the two share every 32-byte sector and run on neighbouring CTAs at the same time.
// Measured on B200 (profiles/r02_k2_tile_*): 3x4 / 24 qubits 873 -> 775 us, 2x6 810 -> 664 us; DRAM traffic 1.27 GB per apply
// (0.27 + 0.54 GB read, 0.46 GB written) against 2.0 GB read by the gather kernel, but the pass is bound by the shared-memory
// pipe and the dependent LDS -> DFMA chain (ncu: L1 / shared 88 % busy, short-scoreboard stalls 50-60 %), not by HBM.
// replaces qml.expval(qml.Hamiltonian) on the state, models/adapt_vqe.py:357,361 (reference: PennyLane default.qubit).
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "devmath.cuh"

extern long long g_fh_launch_count;

#define TT_BITS 12
#define TT_LOW 9                // tile bits covered by the thread index
#define TT_THREADS (1 << TT_LOW)
#define TT_R (1 << (TT_BITS - TT_LOW))      // amplitudes per thread: the settings of the top tile bits
#define TT_LOWMASK ((1u << TT_LOW) - 1u)
#if TT_BITS - TT_LOW <= 3
typedef unsigned short tt_code_t;      // live bits [0, TT_R), sign bits [TT_R, 2 TT_R)
#else
typedef unsigned tt_code_t;
#endif
#define TT_MAX_PASSES 3
#define TT_MAX_GROUPS 96        // per pass
#define TT_MAX_CLASSES 192      // per pass
#define TT_MAX_VALS 1024        // whole table

struct __align__(16) TTGroup {      // 48 bytes
    unsigned xl;                    // x-mask in tile-local coordinates
    unsigned live;                  // live x-bit patterns
    unsigned char pos[4];           // tile-local positions of the x bits (unused: 31)
    int first_class;                // classes of this group in the pass's class array
    int n_class;
    unsigned rpat_lo, rpat_hi;      // pattern(j ^ (r << TT_LOW)) = pattern(j) ^ nibble r of (rpat_hi:rpat_lo)
    unsigned vneg;                  // uniform groups: bit p set when the table entry of pattern p is negative
    double v;                       // uniform groups: common magnitude of the live table entries
    double pad;
};
static_assert(sizeof(TTGroup) == 48, "TTGroup layout");
struct __align__(16) TTClass {      // 16 bytes
    unsigned zl;                    // in-tile zeta bits, tile-local coordinates
    unsigned zo;                    // out-of-tile zeta bits, global coordinates
    int vofs;                       // first entry of the class table in the value array
    unsigned rsgn;                  // bit r: parity((r << TT_LOW) & zl)
};
struct TTPass {
    int ngroups, nclasses, first_group, first_class;
    int ndiag;                      // x = 0 classes (terms straddling the 12-bit chunks of the factor tables): first pass only
    int first_diag;
    int uniform;                    // 1: every group is real, single-class, with one magnitude over its live patterns
    int pad;
    unsigned char bits[16];         // ascending global positions of the 12 tile bits
    unsigned char rest[32];         // ascending global positions of the n - 12 other bits
};

struct fh_table_tiles {
    int npasses = 0;
    TTPass pass[TT_MAX_PASSES];
    TTGroup *d_groups = nullptr;
    TTClass *d_classes = nullptr;
};

void fh_table_tiles_free(fh_table_tiles *t) {
    if (!t) return;
    cudaFree(t->d_groups);
    cudaFree(t->d_classes);
    delete t;
}

__device__ __forceinline__ double flip_sign64(double v, unsigned sbit) {     // sbit = 0 or 0x80000000
    return __hiloint2double(__double2hiint(v) ^ (int)sbit, __double2loint(v));
}

// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 tt_ld_shared_if(unsigned addr, unsigned pred) {
    double2 v;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\tmov.f64 %1, 0d0000000000000000;\n\t"
                 "@p ld.shared.v2.f64 {%0, %1}, [%2];\n\t}"
                 : "=d"(v.x), "=d"(v.y)
                 : "r"(addr), "r"(pred));
    return v;
}

// UNIFORM (implies REAL): every group of the pass has ONE class and ONE magnitude over its live patterns (every hopping
// group of a Hubbard model): which of a thread's amplitudes are live for a group, and with which sign, does not depend
// on the tile -- a 32-bit code per (group, thread) computed once per launch; the op loop is then predicated partner load +
// sign flip + two FMAs per (amplitude, group), with no table lookup.
template <bool REAL, int MODE, bool UNIFORM>      // MODE 0: expectation only; 1: out = acc (first pass) / out += acc; 2: out += acc
__global__ void __launch_bounds__(TT_THREADS, 2)
    k_table_pass(const TTPass P, const TTGroup *__restrict__ groups, const TTClass *__restrict__ classes,
                 const double2 *__restrict__ vals, int nvals, const double2 *__restrict__ dtab, const double2 *__restrict__ in,
                 double2 *__restrict__ out, int n, int first_pass, int last_pass, int pass_index, double *__restrict__ partials,
                 unsigned *__restrict__ counter, double *__restrict__ result, int total_partials) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[TT_THREADS / 32];
    __shared__ unsigned s_hi[TT_R];          // global offset of the four top tile bits
    __shared__ unsigned s_base;
    double2 *tile = reinterpret_cast<double2 *>(smem_raw);                         // 2^12 amplitudes
    TTGroup *sg = reinterpret_cast<TTGroup *>(tile + (1 << TT_BITS));
    TTClass *sc = reinterpret_cast<TTClass *>(sg + P.ngroups);
    double2 *sv = reinterpret_cast<double2 *>(sc + P.nclasses + P.ndiag);
    unsigned char *bsgn = reinterpret_cast<unsigned char *>(sv + nvals);           // per class: parity(base & zo)
    tt_code_t *codes = reinterpret_cast<tt_code_t *>(bsgn + ((P.nclasses + P.ndiag + 15) & ~15));     // UNIFORM: [group][thread]
    const unsigned tid = threadIdx.x;

    for (unsigned t = tid; t < (unsigned)P.ngroups * 3u; t += TT_THREADS)
        reinterpret_cast<uint4 *>(sg)[t] = __ldg(reinterpret_cast<const uint4 *>(groups + P.first_group) + t);
    for (unsigned t = tid; t < (unsigned)P.nclasses; t += TT_THREADS)
        reinterpret_cast<uint4 *>(sc)[t] = __ldg(reinterpret_cast<const uint4 *>(classes + P.first_class) + t);
    for (unsigned t = tid; t < (unsigned)P.ndiag; t += TT_THREADS)
        reinterpret_cast<uint4 *>(sc)[P.nclasses + t] = __ldg(reinterpret_cast<const uint4 *>(classes + P.first_diag) + t);
    for (unsigned t = tid; t < (unsigned)nvals; t += TT_THREADS) sv[t] = vals[t];
    // global offset of this thread's low tile bits and of the 16 settings of the top four
    unsigned lo_off = 0;
#pragma unroll
    for (int b = 0; b < TT_LOW; ++b) lo_off |= ((tid >> b) & 1u) << P.bits[b];
    if (tid < TT_R) {
        unsigned o = 0;
#pragma unroll
        for (int b = 0; b < TT_BITS - TT_LOW; ++b) o |= ((tid >> b) & 1u) << P.bits[TT_LOW + b];
        s_hi[tid] = o;
    }
    if (UNIFORM) {
        __syncthreads();
        for (int g = 0; g < P.ngroups; ++g) {
            const TTGroup G = sg[g];
            const TTClass Cl = sc[G.first_class];
            const unsigned j0 = tid ^ G.xl;
            const unsigned pat0 = ((j0 >> G.pos[0]) & 1u) | (((j0 >> G.pos[1]) & 1u) << 1) | (((j0 >> G.pos[2]) & 1u) << 2) |
                                  (((j0 >> G.pos[3]) & 1u) << 3);
            const unsigned s0 = (unsigned)(__popc(j0 & Cl.zl) & 1);
            unsigned code = 0;
#pragma unroll
            for (int r = 0; r < TT_R; ++r) {
                const unsigned pat = pat0 ^ (((r < 8 ? G.rpat_lo : G.rpat_hi) >> (4 * (r & 7))) & 15u);
                code |= ((G.live >> pat) & 1u) << r;
                code |= ((s0 ^ (Cl.rsgn >> r) ^ (G.vneg >> pat)) & 1u) << (TT_R + r);
            }
            codes[g * TT_THREADS + tid] = (tt_code_t)code;
        }
    }
    const unsigned tile_u32 = (unsigned)__cvta_generic_to_shared(tile);
    const unsigned ntiles = 1u << (n - TT_BITS);
    double er = 0.0, ei = 0.0;
    for (unsigned t = blockIdx.x; t < ntiles; t += gridDim.x) {
        __syncthreads();                 // previous tile fully consumed (and the descriptors are in place)
        if (tid == 0) {
            unsigned base = 0;
#pragma unroll
            for (int b = 0; b < 20; ++b)          // compile-time subscripts: the by-value pass stays in the constant bank
                if (b < n - TT_BITS) base |= ((t >> b) & 1u) << P.rest[b];
            s_base = base;
        }
        __syncthreads();
        const unsigned base = s_base;
        for (unsigned c = tid; c < (unsigned)(P.nclasses + P.ndiag); c += TT_THREADS) bsgn[c] = (unsigned char)(__popc(base & sc[c].zo) & 1);
        const unsigned g0 = base | lo_off;
#pragma unroll
        for (int r = 0; r < TT_R; ++r) tile[tid | ((unsigned)r << TT_LOW)] = in[g0 | s_hi[r]];
        __syncthreads();

        double ar[TT_R], ai[TT_R];
#pragma unroll
        for (int r = 0; r < TT_R; ++r) ar[r] = ai[r] = 0.0;
        for (int g = 0; g < P.ngroups; ++g) {
            const TTGroup G = sg[g];
            const unsigned j0 = tid ^ G.xl;                  // partner of r = 0 (12-bit local index)
            const unsigned jlow = j0 & TT_LOWMASK, xh = j0 >> TT_LOW;
            if (UNIFORM) {
                const unsigned code = codes[g * TT_THREADS + tid];
                const unsigned neg = (code >> TT_R) ^ (bsgn[G.first_class] ? 0xffffu : 0u);
                const unsigned plow = tile_u32 + (jlow << 4), xh12 = xh << (TT_LOW + 4);      // r flips the top bits of the tile OFFSET
                const int vhi = __double2hiint(G.v), vlo = __double2loint(G.v);
#pragma unroll
                for (int r = 0; r < TT_R; ++r) {
                    // the partner load is predicated (a dead amplitude costs no shared-memory wavefront and adds 0)
                    const double wr = __hiloint2double(vhi ^ (int)((neg << (31 - r)) & 0x80000000u), vlo);
                    asm volatile(
                        "{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t.reg .f64 x, y;\n\t"
                        "and.b32 t, %3, %4;\n\tsetp.ne.u32 p, t, 0;\n\t"
                        "mov.f64 x, 0d0000000000000000;\n\tmov.f64 y, 0d0000000000000000;\n\t"
                        "@p ld.shared.v2.f64 {x, y}, [%2];\n\t"
                        "fma.rn.f64 %0, %5, x, %0;\n\tfma.rn.f64 %1, %5, y, %1;\n\t}"
                        : "+d"(ar[r]), "+d"(ai[r])
                        : "r"(plow + (xh12 ^ ((unsigned)r << (TT_LOW + 4)))), "r"(code), "r"(1u << r), "d"(wr));
                }
                continue;
            }
            const unsigned pat0 = ((j0 >> G.pos[0]) & 1u) | (((j0 >> G.pos[1]) & 1u) << 1) | (((j0 >> G.pos[2]) & 1u) << 2) |
                                  (((j0 >> G.pos[3]) & 1u) << 3);
            if (G.n_class == 1) {
                const TTClass Cl = sc[G.first_class];
                const unsigned s0 = (unsigned)(__popc(j0 & Cl.zl) & 1) ^ (unsigned)bsgn[G.first_class];
                const double2 *V = sv + Cl.vofs;
#pragma unroll
                for (int r = 0; r < TT_R; ++r) {
                    // no liveness test: a dead pattern has a zero table entry, and a branch here diverges in every warp whose
                    // lanes differ in an x bit of the group
                    const unsigned pat = pat0 ^ (((r < 8 ? G.rpat_lo : G.rpat_hi) >> (4 * (r & 7))) & 15u);
                    const double2 w = V[pat];
                    const double2 pv = tile[jlow | (((unsigned)r ^ xh) << TT_LOW)];
                    const unsigned sbit = ((s0 ^ (Cl.rsgn >> r)) & 1u) << 31;
                    const double wr = flip_sign64(w.x, sbit);
                    if (REAL) {
                        ar[r] += wr * pv.x;
                        ai[r] += wr * pv.y;
                    } else {
                        const double wi = flip_sign64(w.y, sbit);
                        ar[r] += wr * pv.x - wi * pv.y;
                        ai[r] += wr * pv.y + wi * pv.x;
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < TT_R; ++r) {
                    const unsigned pat = pat0 ^ (((r < 8 ? G.rpat_lo : G.rpat_hi) >> (4 * (r & 7))) & 15u);
                    double wr = 0.0, wi = 0.0;
                    for (int c = G.first_class; c < G.first_class + G.n_class; ++c) {
                        const TTClass Cl = sc[c];
                        const unsigned sbit = (((unsigned)(__popc(j0 & Cl.zl) & 1) ^ (unsigned)bsgn[c] ^ (Cl.rsgn >> r)) & 1u) << 31;
                        const double2 w = sv[Cl.vofs + pat];
                        wr += flip_sign64(w.x, sbit);
                        if (!REAL) wi += flip_sign64(w.y, sbit);
                    }
                    const double2 pv = tile[jlow | (((unsigned)r ^ xh) << TT_LOW)];
                    if (REAL) {
                        ar[r] += wr * pv.x;
                        ai[r] += wr * pv.y;
                    } else {
                        ar[r] += wr * pv.x - wi * pv.y;
                        ai[r] += wr * pv.y + wi * pv.x;
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < TT_R; ++r) {
            const unsigned i = g0 | s_hi[r];
            const double2 self = tile[tid | ((unsigned)r << TT_LOW)];
            if (first_pass) {
                // diagonal: three additive factor tables over 12-bit chunks of the index + the chunk-straddling terms
                double dr = 0.0, di = 0.0;
                if (dtab) {
                    const double2 d = cadd(cadd(__ldg(dtab + (i & 0xfffu)), __ldg(dtab + 4096 + ((i >> 12) & 0xfffu))),
                                           __ldg(dtab + 8192 + (i >> 24)));
                    dr = d.x;
                    di = d.y;
                }
                for (int c = 0; c < P.ndiag; ++c) {
                    const TTClass Cl = sc[P.nclasses + c];
                    const double2 w = sv[Cl.vofs];
                    const bool neg = (__popc(i & Cl.zo) & 1) != 0;
                    dr += neg ? -w.x : w.x;
                    di += neg ? -w.y : w.y;
                }
                ar[r] += dr * self.x - di * self.y;
                ai[r] += dr * self.y + di * self.x;
            }
            er += self.x * ar[r] + self.y * ai[r];
            ei += self.x * ai[r] - self.y * ar[r];
            if (MODE == 1 && first_pass) {
                out[i] = make_double2(ar[r], ai[r]);
            } else if (MODE != 0) {
                const double2 o = out[i];
                out[i] = make_double2(o.x + ar[r], o.y + ai[r]);
            }
        }
    }
    const double sr = block_sum<TT_THREADS>(er, red);
    const double si = block_sum<TT_THREADS>(ei, red);
    // per-CTA partials of every pass; the last CTA of the LAST pass folds them all in slot order (deterministic)
    __shared__ unsigned is_last;
    if (tid == 0) {
        partials[2 * (pass_index * (int)gridDim.x + (int)blockIdx.x)] = sr;
        partials[2 * (pass_index * (int)gridDim.x + (int)blockIdx.x) + 1] = si;
        __threadfence();
        is_last = (last_pass && atomicAdd(counter, 1u) == gridDim.x - 1u) ? 1u : 0u;
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double r = 0.0, im = 0.0;
        for (int t = (int)tid; t < total_partials; t += TT_THREADS) {
            r += __ldcg(partials + 2 * t);
            im += __ldcg(partials + 2 * t + 1);
        }
        const double fr = block_sum<TT_THREADS>(r, red);
        const double fi = block_sum<TT_THREADS>(im, red);
        if (tid == 0) {
            result[0] = fr;
            result[1] = fi;
            *counter = 0u;
        }
    }
}

// ----------------------------------------------------------------------------------------------
// host: cover the x-mask groups by passes of 12 index bits
// ----------------------------------------------------------------------------------------------
int fh_table_plan_tiles(fh_table *tab) {
    tab->tiles = nullptr;
    const int n = tab->n;
    if (n < 22 || n > 31) return FH_OK;
    // FHSIM_K2_TILE: 0 never; 1 (default) only when every pass takes the uniform fast path (measured 11-18 % faster than the
    // gather kernel at 24 qubits; the general tile path is slower than the gather kernel); 2 every coverable table
    int policy = 1;
    if (const char *env = getenv("FHSIM_K2_TILE")) policy = atoi(env);
    if (policy == 0) return FH_OK;
    if (tab->vals.size() > TT_MAX_VALS) return FH_OK;
    std::vector<int> remaining;
    int diag_group = -1;
    for (int g = 0; g < (int)tab->groups.size(); ++g) {
        const TabGroup &G = tab->groups[g];
        if (G.x == 0) {
            diag_group = g;
            continue;
        }
        if (G.kbits == 0) return FH_OK;           // more than 4 x bits: only the gather kernel takes those
        remaining.push_back(g);
    }
    if (remaining.empty()) return FH_OK;
    auto popc = [](u64 v) { return __builtin_popcountll(v); };
    std::vector<TTGroup> groups;
    std::vector<TTClass> classes;
    fh_table_tiles *T = new fh_table_tiles();
    while (!remaining.empty()) {
        if (T->npasses == TT_MAX_PASSES) {
            delete T;
            return FH_OK;
        }
        u64 S = tab->groups[remaining[0]].x;
        for (;;) {
            // grow along connected groups first (a group sharing a bit with S), among those the one that needs the fewest
            // new bits; a disconnected group only when no connected one fits
            int best = -1, best_new = 99, best_conn = 0;
            for (int g : remaining) {
                const u64 x = tab->groups[g].x;
                const int add = popc(x & ~S), conn = (x & S) ? 1 : 0;
                if (add == 0 || popc(S | x) > TT_BITS) continue;
                if (conn > best_conn || (conn == best_conn && add < best_new)) {
                    best = g;
                    best_new = add;
                    best_conn = conn;
                }
            }
            if (best < 0) break;
            S |= tab->groups[best].x;
        }
        for (int b = 0; b < n && popc(S) < TT_BITS; ++b) S |= 1ull << b;       // pad with the lowest free bits
        TTPass &P = T->pass[T->npasses];
        memset(&P, 0, sizeof(P));
        int nb = 0, nr = 0;
        for (int b = 0; b < n; ++b) {
            if (S >> b & 1ull) P.bits[nb++] = (unsigned char)b;
            else P.rest[nr++] = (unsigned char)b;
        }
        auto local = [&](u64 m) {
            unsigned o = 0;
            for (int k = 0; k < TT_BITS; ++k)
                if (m >> P.bits[k] & 1ull) o |= 1u << k;
            return o;
        };
        P.first_group = (int)groups.size();
        P.first_class = (int)classes.size();
        bool pass_uniform = true;
        std::vector<int> left;
        for (int g : remaining) {
            const TabGroup &G = tab->groups[g];
            if (G.x & ~S) {
                left.push_back(g);
                continue;
            }
            TTGroup tg;
            memset(&tg, 0, sizeof(tg));
            tg.xl = local(G.x);
            tg.live = G.live;
            tg.first_class = (int)classes.size() - P.first_class;
            tg.n_class = G.n_class;
            u64 rp = 0;
            for (int k = 0; k < 4; ++k) {
                if (k >= G.kbits) {
                    tg.pos[k] = 31;
                    continue;
                }
                int lp = 0;
                while (P.bits[lp] != G.pos[k]) ++lp;
                tg.pos[k] = (unsigned char)lp;
                if (lp >= TT_LOW)
                    for (int r = 0; r < TT_R; ++r)
                        if (r >> (lp - TT_LOW) & 1) rp |= (u64)(1u << k) << (4 * r);
            }
            tg.rpat_lo = (unsigned)(rp & 0xffffffffull);
            tg.rpat_hi = (unsigned)(rp >> 32);
            // one class, real, one magnitude over the live patterns?
            bool uni = tab->all_real && G.n_class == 1;
            if (uni) {
                const int vofs = tab->classes[G.first_class].vofs;
                double mag = -1.0;
                for (int pat = 0; pat < (1 << G.kbits) && uni; ++pat) {
                    const double2 w = tab->vals[vofs + pat];
                    if (w.x == 0.0 && w.y == 0.0) continue;
                    if (mag < 0.0) mag = fabs(w.x);
                    if (w.y != 0.0 || fabs(w.x) != mag) uni = false;
                    if (w.x < 0.0) tg.vneg |= 1u << pat;
                }
                tg.v = mag > 0.0 ? mag : 0.0;
            }
            if (!uni) pass_uniform = false;
            for (int c = G.first_class; c < G.first_class + G.n_class; ++c) {
                TTClass tc;
                tc.zl = local(tab->classes[c].zeta & S);
                tc.zo = (unsigned)(tab->classes[c].zeta & ~S);
                tc.vofs = tab->classes[c].vofs;
                tc.rsgn = 0;
                for (int r = 0; r < TT_R; ++r)
                    if (__builtin_popcount(((unsigned)r << TT_LOW) & tc.zl) & 1) tc.rsgn |= 1u << r;
                classes.push_back(tc);
            }
            groups.push_back(tg);
        }
        P.ngroups = (int)groups.size() - P.first_group;
        P.nclasses = (int)classes.size() - P.first_class;
        P.uniform = (pass_uniform && P.ngroups <= 32) ? 1 : 0;
        if (P.ngroups > TT_MAX_GROUPS || P.nclasses > TT_MAX_CLASSES) {
            delete T;
            return FH_OK;
        }
        remaining.swap(left);
        T->npasses++;
    }
    if (policy < 2)
        for (int p = 0; p < T->npasses; ++p)
            if (!T->pass[p].uniform) {
                delete T;
                return FH_OK;
            }
    // x = 0 classes (diagonal terms that straddle the factor-table chunks): evaluated per amplitude in the first pass
    T->pass[0].first_diag = (int)classes.size();
    if (diag_group >= 0) {
        const TabGroup &G = tab->groups[diag_group];
        for (int c = G.first_class; c < G.first_class + G.n_class; ++c) {
            TTClass tc;
            tc.zl = 0;
            tc.zo = (unsigned)tab->classes[c].zeta;
            tc.vofs = tab->classes[c].vofs;
            tc.rsgn = 0;
            classes.push_back(tc);
        }
        T->pass[0].ndiag = G.n_class;
        if (G.n_class > 64) {
            delete T;
            return FH_OK;
        }
    }
    if (cudaMalloc(&T->d_groups, sizeof(TTGroup) * groups.size()) != cudaSuccess ||
        cudaMalloc(&T->d_classes, sizeof(TTClass) * std::max<size_t>(1, classes.size())) != cudaSuccess) {
        cudaGetLastError();
        fh_table_tiles_free(T);
        return FH_OK;
    }
    cudaMemcpy(T->d_groups, groups.data(), sizeof(TTGroup) * groups.size(), cudaMemcpyHostToDevice);
    if (!classes.empty()) cudaMemcpy(T->d_classes, classes.data(), sizeof(TTClass) * classes.size(), cudaMemcpyHostToDevice);
    tab->tiles = T;
    return FH_OK;
}

extern "C" int fh_table_tile_passes(const fh_table *tab, int *n_passes) {
    FH_REQUIRE(tab && n_passes, "fh_table_tile_passes: NULL argument");
    *n_passes = tab->tiles ? tab->tiles->npasses : 0;
    return FH_OK;
}

static bool g_tt_attr[64];
static void tt_init_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (g_tt_attr[dev]) return;
    const int bytes = 110 * 1024;
#define TT_SET(R, M)                                                                                            \
    cudaFuncSetAttribute(k_table_pass<R, M, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);        \
    if (R) cudaFuncSetAttribute(k_table_pass<true, M, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)
    TT_SET(true, 0); TT_SET(true, 1); TT_SET(true, 2); TT_SET(false, 0); TT_SET(false, 1); TT_SET(false, 2);
#undef TT_SET
    g_tt_attr[dev] = true;
}

// true: the tile passes were enqueued (result -> d_result[0..1]); false: not applicable, the caller uses the gather kernel
bool launch_apply_table_tiles(cudaStream_t s, int sm, const fh_table *tab, const double2 *in, double2 *out, int mode,
                              double *d_partials, double *d_result) {
    const fh_table_tiles *T = tab->tiles;
    if (!T || T->npasses == 0) return false;
    if (getenv("FHSIM_K2_GATHER")) return false;             // read per call: tests / A-B runs flip it
    tt_init_device();
    const int nvals = (int)tab->vals.size();
    const unsigned ntiles = 1u << (tab->n - TT_BITS);
    int grid = sm * 2;
    if ((unsigned)grid > ntiles) grid = (int)ntiles;
    if (T->npasses * grid > FH_MAX_PARTIALS) return false;
    auto pass_smem = [&](const TTPass &P) {
        return ((size_t)16 << TT_BITS) + (size_t)P.ngroups * sizeof(TTGroup) + (size_t)(P.nclasses + P.ndiag) * sizeof(TTClass) +
               (size_t)nvals * 16 + (size_t)((P.nclasses + P.ndiag + 15) & ~15) + (P.uniform ? (size_t)P.ngroups * TT_THREADS * sizeof(tt_code_t) : 0) + 16;
    };
    for (int p = 0; p < T->npasses; ++p)
        if (pass_smem(T->pass[p]) > 110 * 1024) return false;
    for (int p = 0; p < T->npasses; ++p) {
        const TTPass &P = T->pass[p];
        const size_t smem = pass_smem(P);
        const int first = p == 0, last = p == T->npasses - 1;
        ++g_fh_launch_count;
#define TT_LAUNCH(R, M, U)                                                                                                   \
    k_table_pass<R, M, U><<<grid, TT_THREADS, smem, s>>>(P, T->d_groups, T->d_classes, tab->d_vals, nvals, tab->d_diag, in, out, \
                                                         tab->n, first, last, p, d_partials, tab->ctx->d_counter, d_result,    \
                                                         T->npasses * grid)
        if (tab->all_real && P.uniform) {
            if (mode == 0) TT_LAUNCH(true, 0, true); else if (mode == 1) TT_LAUNCH(true, 1, true); else TT_LAUNCH(true, 2, true);
        } else if (tab->all_real) {
            if (mode == 0) TT_LAUNCH(true, 0, false); else if (mode == 1) TT_LAUNCH(true, 1, false); else TT_LAUNCH(true, 2, false);
        } else {
            if (mode == 0) TT_LAUNCH(false, 0, false); else if (mode == 1) TT_LAUNCH(false, 1, false); else TT_LAUNCH(false, 2, false);
        }
#undef TT_LAUNCH
    }
    return true;
}
