// K1, fused tile runs with TMA tile movement (sm_100a).
//
// A tile = all 2^T amplitudes that differ only in T chosen index bits.  The gather "every amplitude whose other bits
// equal this tile's" is done by the Tensor Memory Accelerator instead of by per-thread address arithmetic: the state
// is described to cuTensorMapEncodeTiled as a tensor of doubles whose dim 0 is the WHOLE flat state (stride 8 B) and
// whose higher dims are the runs of consecutive tile bits, each with stride 16 B << (first bit of the run) -- the
// strides deliberately overlap dim 0.  One cp.async.bulk.tensor.5d with box = (low run, run, run, run, run) and dim-0
// coordinate = 2 * (tile base index) then lands the tile in shared memory in tile-local index order; tile bits that do
// not fit the 5 dims are iterated (2^extra boxes per tile, issued by the lanes of warp 0).  The same map stores the
// tile back (cp.async.bulk.tensor ... bulk_group).  Verified on B200 by tools/probes/probe_tma.cu (profiles/r02_*).
//
// Shared-memory layout: with dim 0 = exactly 8 amplitudes (128 B) the map uses CU_TENSOR_MAP_SWIZZLE_128B, i.e. local
// index l lives in 16-byte slot l ^ ((l >> 3) & 7): the strided pair accesses of the op loop spread over the eight
// 16-byte bank groups.  Tiles whose lowest run is shorter than 3 bits use the linear layout.
//
// Persistent CTAs: when there are more tiles than resident CTAs every CTA owns two tile buffers and the load of tile
// i+1 (and the store of tile i-1) overlap the op loop of tile i (mbarrier complete_tx for loads, bulk groups for
// stores).  Op loop: the index/sign decode of op k+1 is computed before the barrier that ends op k, so the dependent
// chain per op is LDS -> FP64 -> STS -> barrier.
#include <cuda.h>
#include <string.h>

#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "devmath.cuh"

extern long long g_fh_launch_count;
extern thread_local int g_fh_tile_pdl_scope;

void launch_tile_ldg(cudaStream_t s, double2 *psi, const TileLaunch &tl, const TileRec *d_recs, const TileTerm *d_terms,
                     int n);

// ----------------------------------------------------------------------------------------------
// host: tensor maps
// ----------------------------------------------------------------------------------------------
struct TmaPlan {
    int map_bits;              // tile-local bits [0, map_bits) are covered by one box
    int n_extra;               // remaining local bits: 2^n_extra boxes per tile
    int swizzle;               // 1: CU_TENSOR_MAP_SWIZZLE_128B layout
    int pad;
    unsigned char extra[8];    // global positions of the extra bits (ascending)
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        return q == cudaDriverEntryPointSuccess ? reinterpret_cast<EncodeTiledFn>(p) : nullptr;
    }();
    return fn;
}

struct TmaEntry {
    CUtensorMap map;
    TmaPlan plan;
    bool ok;
};

struct TmaKey {
    const void *ptr;
    int n, T;
    unsigned char bits[16];
    bool operator<(const TmaKey &o) const { return memcmp(this, &o, sizeof(TmaKey)) < 0; }
};

// dims of the box for a tile bit set: (first global bit, number of bits) per dim, dim 0 first
static void plan_dims(const unsigned char *bits, int T, bool want_swizzle, std::vector<std::pair<int, int>> &dims,
                      TmaPlan &plan) {
    std::vector<std::pair<int, int>> runs;
    for (int b = 0; b < T; ++b) {
        if (!runs.empty() && runs.back().first + runs.back().second == bits[b]) runs.back().second++;
        else runs.push_back({bits[b], 1});
    }
    dims.clear();
    const int r0 = runs[0].second;
    const bool swz = want_swizzle && r0 >= 3;
    int first = swz ? 3 : (r0 < 7 ? r0 : 7);          // dim 0 box <= 256 doubles
    dims.push_back({0, first});
    for (int done = first; done < r0;) {
        const int take = r0 - done > 8 ? 8 : r0 - done;
        dims.push_back({done, take});
        done += take;
    }
    for (size_t k = 1; k < runs.size(); ++k)
        for (int done = 0; done < runs[k].second;) {
            const int take = runs[k].second - done > 8 ? 8 : runs[k].second - done;
            dims.push_back({runs[k].first + done, take});
            done += take;
        }
    if (dims.size() > 5) dims.resize(5);
    int covered = 0;
    for (auto &d : dims) covered += d.second;
    memset(&plan, 0, sizeof(plan));
    plan.map_bits = covered;
    plan.n_extra = T - covered;
    plan.swizzle = swz ? 1 : 0;
    for (int b = covered; b < T && b - covered < 8; ++b) plan.extra[b - covered] = bits[b];
}

static std::mutex g_tma_mutex;
static std::map<TmaKey, TmaEntry> g_tma_cache;

// tensor map + box plan for (state buffer, tile bit set); nullptr when the TMA path does not apply
static const TmaEntry *tma_entry(const double2 *psi, int n, const TileLaunch &tl) {
    const int T = tl.nbits;
    if (n > 30 || T < 3 || T > FH_MAX_TILE_BITS || tl.bits[0] != 0) return nullptr;       // dim-0 coordinate is an int32 count of doubles
    TmaKey key;
    memset(&key, 0, sizeof(key));
    key.ptr = psi;
    key.n = n;
    key.T = T;
    memcpy(key.bits, tl.bits, 16);
    std::lock_guard<std::mutex> lock(g_tma_mutex);
    auto it = g_tma_cache.find(key);
    if (it != g_tma_cache.end()) return it->second.ok ? &it->second : nullptr;
    TmaEntry e;
    memset(&e, 0, sizeof(e));
    e.ok = false;
    EncodeTiledFn enc = encode_fn();
    if (enc) {
        std::vector<std::pair<int, int>> dims;
        plan_dims(tl.bits, T, true, dims, e.plan);
        // every box must start on a 1024-byte boundary of the tile buffer for the 128-byte swizzle, and the bulk copy
        // of 2^n_extra boxes should stay a handful of instructions
        if (e.plan.swizzle && e.plan.map_bits < 6) plan_dims(tl.bits, T, false, dims, e.plan);
        if (e.plan.n_extra <= 6) {
            cuuint64_t gdim[5], gstride[4];
            cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
            gdim[0] = 2ull << n;
            box[0] = 2u << dims[0].second;
            for (int k = 1; k < 5; ++k) {
                if (k < (int)dims.size()) {
                    gdim[k] = 1ull << dims[k].second;
                    box[k] = 1u << dims[k].second;
                    gstride[k - 1] = 16ull << dims[k].first;
                } else {
                    gdim[k] = 1;
                    box[k] = 1;
                    gstride[k - 1] = 16ull << n;
                }
            }
            const CUresult r = enc(&e.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, const_cast<double2 *>(psi), gdim, gstride, box,
                                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   e.plan.swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            e.ok = (r == CUDA_SUCCESS);
        }
    }
    if (g_tma_cache.size() > 4096) g_tma_cache.clear();
    auto ins = g_tma_cache.emplace(key, e);
    return ins.first->second.ok ? &ins.first->second : nullptr;
}

// ----------------------------------------------------------------------------------------------
// device: TMA / mbarrier primitives
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra WAIT_%=;\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(unsigned dst, const CUtensorMap *map, unsigned bar, int c0) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %4, %4, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(0)
        : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap *map, unsigned src, int c0) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %3, %3, %3}], [%1];" ::"l"(map), "r"(src),
                 "r"(c0), "r"(0)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <bool SWZ>
__device__ __forceinline__ unsigned tslot(unsigned l) {
    return SWZ ? (l ^ ((l >> 3) & 7u)) : l;
}

// dim-0 coordinate (in doubles) of box q of the tile with base index `base`
__device__ __forceinline__ int box_coord(const TmaPlan &plan, unsigned base, unsigned q) {
    unsigned g = base;
#pragma unroll
    for (int e = 0; e < 8; ++e)
        if (e < plan.n_extra) g |= ((q >> e) & 1u) << plan.extra[e];
    return (int)(g << 1);
}

// all lanes of warp 0 call these; lane q moves box q, q + 32, ...
__device__ __forceinline__ void warp_load_tile(const CUtensorMap *map, const TmaPlan &plan, unsigned base, unsigned dst,
                                               unsigned bar, unsigned tile_bytes, int lane) {
    if (lane == 0) mbar_expect_tx(bar, tile_bytes);
    __syncwarp();
    const unsigned nbox = 1u << plan.n_extra, box_bytes = 16u << plan.map_bits;
    for (unsigned q = lane; q < nbox; q += 32) tma_load_5d(dst + q * box_bytes, map, bar, box_coord(plan, base, q));
}
__device__ __forceinline__ void warp_store_tile(const CUtensorMap *map, const TmaPlan &plan, unsigned base, unsigned src,
                                                int lane) {
    const unsigned nbox = 1u << plan.n_extra, box_bytes = 16u << plan.map_bits;
    for (unsigned q = lane; q < nbox; q += 32) tma_store_5d(map, src + q * box_bytes, box_coord(plan, base, q));
    bulk_commit();
}

__device__ __forceinline__ double flip_sign(double v, unsigned sbit) {
    return __hiloint2double(__double2hiint(v) ^ (int)sbit, __double2loint(v));
}

// Pair ops are driven by a per-thread table built on the host once per program (program.cu: build_tile_records): 16-bit
// entry [rep * threads + tid] of an op = slot of the pattern side | in-tile sign parity << 13 | valid << 14; the partner
// slot is entry ^ xs with xs = slot image of the op's x-mask (the slot map is XOR-linear).  The table of a run arrives in
// shared memory by bulk copy next to the op records, so the op loop does no index arithmetic at all:
// LDS entry -> LDS amplitudes -> FP64 -> STS.
struct OpHdr {
    unsigned fixmask_out, fixval_out, zeta;   // pattern / sign bits outside the tile (uniform per tile)
    unsigned xs;                              // slot image of the x-mask
    int type, reps, off;                      // record type, table entries per thread, first table entry (or first term)
    int run_len;                              // >= 2: first record of a register-fused run
    unsigned run_bits;
    bool full4;                               // no invalid table entry, entries per thread a multiple of 4, type 3 or 6
    double2 m0, m1, m2, m3;                   // 2x2 block
};

__device__ __forceinline__ void load_hdr(const TileRec *rec, OpHdr &h) {
    const uint4 *rp = reinterpret_cast<const uint4 *>(rec);
    const uint4 q0 = rp[0], q1 = rp[1], q3 = rp[3];
    h.fixmask_out = q0.x;
    h.fixval_out = q0.y;
    h.zeta = q0.z;
    h.type = (int)q1.y;
    h.off = (int)q1.w;
    h.reps = (int)q3.w;
    h.xs = q3.x;
    const uint2 q9 = *reinterpret_cast<const uint2 *>(&rec->run_len);
    h.run_len = (int)q9.x;
    h.run_bits = q9.y;
    // pairs of the op in one tile = 2^T >> nlfix = reps * threads exactly  <=>  no invalid entry
    h.full4 = (h.reps & 3) == 0 && h.reps > 0 && (h.type == 3 || h.type == 6) && rec->run_pad[0] != 0;
    const double2 *mp = reinterpret_cast<const double2 *>(rec->m);
    h.m0 = mp[0];
    h.m1 = mp[1];
    h.m2 = mp[2];
    h.m3 = mp[3];
}

// 2x2 block on one pair held in registers.  TYPE 3: real matrix (Givens); 6: real diagonal, complex off-diagonal (every
// rotation exp(-i a G)); anything else: general complex.
template <int TYPE>
__device__ __forceinline__ void pair_math(const OpHdr &h, unsigned sbit, double2 &a, double2 &b) {
    if (TYPE == 3) {
        const double m01 = flip_sign(h.m1.x, sbit), m10 = flip_sign(h.m2.x, sbit);
        const double2 ra = make_double2(h.m0.x * a.x + m01 * b.x, h.m0.x * a.y + m01 * b.y);
        const double2 rb = make_double2(m10 * a.x + h.m3.x * b.x, m10 * a.y + h.m3.x * b.y);
        a = ra;
        b = rb;
    } else if (TYPE == 6) {
        const double2 mb = make_double2(flip_sign(h.m1.x, sbit), flip_sign(h.m1.y, sbit));
        const double2 mc = make_double2(flip_sign(h.m2.x, sbit), flip_sign(h.m2.y, sbit));
        const double2 ra = make_double2(h.m0.x * a.x + (mb.x * b.x - mb.y * b.y), h.m0.x * a.y + (mb.x * b.y + mb.y * b.x));
        const double2 rb = make_double2(h.m3.x * b.x + (mc.x * a.x - mc.y * a.y), h.m3.x * b.y + (mc.x * a.y + mc.y * a.x));
        a = ra;
        b = rb;
    } else {
        const double2 mb = make_double2(flip_sign(h.m1.x, sbit), flip_sign(h.m1.y, sbit));
        const double2 mc = make_double2(flip_sign(h.m2.x, sbit), flip_sign(h.m2.y, sbit));
        const double2 ra = cadd(cmul(h.m0, a), cmul(mb, b));
        const double2 rb = cadd(cmul(mc, a), cmul(h.m3, b));
        a = ra;
        b = rb;
    }
}

// All pairs of one op that belong to this thread, four at a time: four table entries, then eight independent
// shared-memory loads in flight, then the arithmetic, then the stores (a CTA is only four warps, so the per-op overhead
// -- header, barrier -- is paid by four warps and the latency is hidden by the four independent pairs of each thread).
#ifndef PAIR_UNROLL
#define PAIR_UNROLL 1
#endif
// FH_OPLOOP_VARIANT (tools/probe_timeline.py only): 0 full op loop; 1 no arithmetic; 2 no amplitude loads / stores either;
// 3 no table loads either (header + barrier); 4 barrier only.  Strips the op loop piece by piece to see where the cycles go.
#ifndef FH_OPLOOP_VARIANT
#define FH_OPLOOP_VARIANT 0
#endif
template <int TYPE>
__device__ __forceinline__ void pair_block(double2 *buf, const OpHdr &h, const unsigned short *tp, int nthr, unsigned sb,
                                           unsigned efirst) {
    for (int rep = 0; rep < h.reps; rep += PAIR_UNROLL) {
        unsigned e[PAIR_UNROLL];
#pragma unroll
        for (int u = 0; u < PAIR_UNROLL; ++u)
            e[u] = (rep + u == 0) ? efirst : ((rep + u < h.reps) ? (unsigned)tp[(rep + u) * nthr] : 0u);
        double2 a[PAIR_UNROLL], b[PAIR_UNROLL];
#if FH_OPLOOP_VARIANT >= 2
        if (e[0] == 0xffffffffu) buf[0] = make_double2(0.0, 0.0);          // keep the table loads alive
        (void)a; (void)b; (void)sb;
#else
#pragma unroll
        for (int u = 0; u < PAIR_UNROLL; ++u)
            if (e[u] >> 14) {
                const unsigned si = e[u] & 0x1fffu;
                a[u] = buf[si];
                b[u] = buf[si ^ h.xs];
            }
#if FH_OPLOOP_VARIANT == 0
#pragma unroll
        for (int u = 0; u < PAIR_UNROLL; ++u)
            if (e[u] >> 14) pair_math<TYPE>(h, (((e[u] >> 13) ^ sb) & 1u) << 31, a[u], b[u]);
#endif
#pragma unroll
        for (int u = 0; u < PAIR_UNROLL; ++u)
            if (e[u] >> 14) {
                const unsigned si = e[u] & 0x1fffu;
                buf[si] = a[u];
                buf[si ^ h.xs] = b[u];
            }
#endif
    }
}

// The same for ops whose table has no invalid entry and a multiple of four entries per thread (every Givens of a 2^11
// tile on a 128-thread CTA): no predicates, so the four pairs of a thread form one straight-line block -- four table
// entries, eight shared-memory loads in flight, then the arithmetic of four independent pairs, then eight stores.
template <int TYPE>
__device__ __forceinline__ void pair_block_full4(double2 *buf, const OpHdr &h, const unsigned short *tp, int nthr, unsigned sb,
                                                 unsigned efirst) {
    for (int rep = 0; rep < h.reps; rep += 4) {
        const unsigned e0 = rep == 0 ? efirst : (unsigned)tp[rep * nthr], e1 = tp[(rep + 1) * nthr], e2 = tp[(rep + 2) * nthr], e3 = tp[(rep + 3) * nthr];
        const unsigned s0 = e0 & 0x1fffu, s1 = e1 & 0x1fffu, s2 = e2 & 0x1fffu, s3 = e3 & 0x1fffu;
        double2 a0 = buf[s0], b0 = buf[s0 ^ h.xs], a1 = buf[s1], b1 = buf[s1 ^ h.xs];
        double2 a2 = buf[s2], b2 = buf[s2 ^ h.xs], a3 = buf[s3], b3 = buf[s3 ^ h.xs];
        pair_math<TYPE>(h, (((e0 >> 13) ^ sb) & 1u) << 31, a0, b0);
        pair_math<TYPE>(h, (((e1 >> 13) ^ sb) & 1u) << 31, a1, b1);
        pair_math<TYPE>(h, (((e2 >> 13) ^ sb) & 1u) << 31, a2, b2);
        pair_math<TYPE>(h, (((e3 >> 13) ^ sb) & 1u) << 31, a3, b3);
        buf[s0] = a0;
        buf[s0 ^ h.xs] = b0;
        buf[s1] = a1;
        buf[s1 ^ h.xs] = b1;
        buf[s2] = a2;
        buf[s2 ^ h.xs] = b2;
        buf[s3] = a3;
        buf[s3 ^ h.xs] = b3;
    }
}

// diagonal op on one tile: exp(-i sum_m angle_m sgn_m(index)).  Terms whose in-tile z bits sit entirely in local bits
// 0..5 (or entirely in 6..) are folded into two phase tables built once per (op, tile): 2^min(T,6) + 2^(T-6) entries, each
// the sincos of a signed angle sum; the terms straddling both halves come first in the op's term list (rec->reps of them,
// sorted by the host) and are evaluated per amplitude from the tile-local z-mask.
template <bool SWZ>
__device__ __forceinline__ void apply_diag(double2 *buf, double2 *ph, const TileRec *rec, const TileTerm *tterm,
                                           const TileLaunch &tl, int T, unsigned base, unsigned lomask_g, unsigned himask_g) {
    const TileTerm *dt = tterm + rec->term_off;
    const int cnt = rec->nterms, nstr = rec->reps;
    const unsigned L = 1u << T;
    const unsigned nlo = T < 6 ? L : 64u, nhi = T > 6 ? (L >> 6) : 1u;      // entries of the two phase tables
    for (unsigned v = threadIdx.x; v < nlo + nhi; v += blockDim.x) {
        const bool lo = v < nlo;
        const unsigned gl = base | (lo ? tile_scatter(tl, T, v, 0, 6) : tile_scatter(tl, T, v - nlo, 6, TILE_BITS_CAP));
        // two partial sums of the signed angles (independent add chains), one sincos per entry
        double t0 = 0.0, t1 = 0.0;
        int m = nstr;
        for (; m + 1 < cnt; m += 2) {
            const unsigned z0 = (unsigned)dt[m].z, z1 = (unsigned)dt[m + 1].z;
            const double a0 = (__popc(gl & z0) & 1) ? -dt[m].angle : dt[m].angle;
            const double a1 = (__popc(gl & z1) & 1) ? -dt[m + 1].angle : dt[m + 1].angle;
            if (lo == ((z0 & himask_g) == 0u)) t0 += a0;
            if (lo == ((z1 & himask_g) == 0u)) t1 += a1;
        }
        if (m < cnt) {
            const unsigned z0 = (unsigned)dt[m].z;
            if (lo == ((z0 & himask_g) == 0u)) t0 += (__popc(gl & z0) & 1) ? -dt[m].angle : dt[m].angle;
        }
        double sn, cs;
        sincos(t0 + t1, &sn, &cs);
        ph[lo ? v : 64u + (v - nlo)] = make_double2(cs, -sn);
    }
    __syncthreads();
    for (unsigned l = threadIdx.x; l < L; l += blockDim.x) {
        double2 f = cmul(ph[l & 63u], ph[64u + (l >> 6)]);
        for (int m = 0; m < nstr; ++m) {
            const unsigned par = (unsigned)(__popc(base & (unsigned)dt[m].z) + __popc(l & dt[m].zlocal)) & 1u;
            f = cmul(f, make_double2(dt[m].c, par ? dt[m].s : -dt[m].s));
        }
        const unsigned sl = tslot<SWZ>(l);
        buf[sl] = cmul(f, buf[sl]);
    }
}

// Optional in-kernel timeline (make TIMELINE=1 / tools/probe_timeline.py): thread 0 of CTA 0 stamps clock64() at the phase
// boundaries.  Slots: 0 entry, 1 prologue done, 2 first tile landed, 4+k start of op k (k < 40), 3 ops done, 62 store
// issued, 63 stores complete.  Compiled out by default.
#ifdef FH_TILE_TIMELINE
__device__ long long g_fh_tma_timeline[64];
#define TMA_TLMARK(k)                                                              \
    do {                                                                           \
        if (blockIdx.x == 0 && threadIdx.x == 0) g_fh_tma_timeline[k] = clock64(); \
    } while (0)
extern "C" int fh_debug_tile_tma_timeline(long long *out64) {
    return (int)cudaMemcpyFromSymbol(out64, g_fh_tma_timeline, sizeof(long long) * 64);
}
#else
#define TMA_TLMARK(k) do { } while (0)
#endif

// One Givens-like op of a fused run on the 8 amplitudes v[] of a group (sub-index = the group's three run bits): the op
// rotates the pairs (sub_i, sub_i ^ x3) for both values of the run bit that is not in x3.  The six (x3, sub_i) shapes are
// spelled out so that v[] is only indexed with compile-time constants (it must stay in registers).
template <int TYPE>
__device__ __forceinline__ void run_op(double2 (&v)[8], const OpHdr &h, unsigned shape, unsigned par0, unsigned par1) {
    const unsigned s0 = (par0 & 1u) << 31, s1 = (par1 & 1u) << 31;
    switch (shape) {
        case 0: pair_math<TYPE>(h, s0, v[1], v[2]); pair_math<TYPE>(h, s1, v[5], v[6]); break;   // x3 = 011, pattern 001
        case 1: pair_math<TYPE>(h, s0, v[2], v[1]); pair_math<TYPE>(h, s1, v[6], v[5]); break;   // x3 = 011, pattern 010
        case 2: pair_math<TYPE>(h, s0, v[1], v[4]); pair_math<TYPE>(h, s1, v[3], v[6]); break;   // x3 = 101, pattern 001
        case 3: pair_math<TYPE>(h, s0, v[4], v[1]); pair_math<TYPE>(h, s1, v[6], v[3]); break;   // x3 = 101, pattern 100
        case 4: pair_math<TYPE>(h, s0, v[2], v[4]); pair_math<TYPE>(h, s1, v[3], v[5]); break;   // x3 = 110, pattern 010
        default: pair_math<TYPE>(h, s0, v[4], v[2]); pair_math<TYPE>(h, s1, v[5], v[3]); break;  // x3 = 110, pattern 100
    }
}

// A run of `m` Givens-like ops inside three tile-local bits, applied to 8-amplitude groups held in registers: one load and
// one store of the tile and ONE barrier for the whole run (a 3-site Fourier transform of the separable W is one run).
template <bool SWZ>
__device__ __forceinline__ void fused_run(double2 *buf, const TileRec *rec, int m, unsigned run_bits, int T, unsigned base) {
    const unsigned b0 = run_bits & 255u, b1 = (run_bits >> 8) & 255u, b2 = (run_bits >> 16) & 255u;
    const unsigned m0 = (1u << b0) - 1u, m1 = (1u << b1) - 1u, m2 = (1u << b2) - 1u;
    const unsigned t0 = tslot<SWZ>(1u << b0), t1 = tslot<SWZ>(1u << b1), t2 = tslot<SWZ>(1u << b2);   // slot map is XOR-linear
    const unsigned ngroups = (1u << T) >> 3;
    for (unsigned g = threadIdx.x; g < ngroups; g += blockDim.x) {
        unsigned gl = g;                                        // group index -> tile-local index with the run bits clear
        gl = ((gl & ~m0) << 1) | (gl & m0);
        gl = ((gl & ~m1) << 1) | (gl & m1);
        gl = ((gl & ~m2) << 1) | (gl & m2);
        const unsigned s = tslot<SWZ>(gl);
        double2 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = buf[s ^ ((k & 1) ? t0 : 0u) ^ ((k & 2) ? t1 : 0u) ^ ((k & 4) ? t2 : 0u)];
        for (int o = 0; o < m; ++o) {
            OpHdr h;
            load_hdr(&rec[o], h);
            const unsigned si = rec[o].sub_i, x3 = rec[o].x3;
            const unsigned third = 7u & ~x3;                    // the run bit this op does not flip
            // tile-local index of the two pattern-side amplitudes (third bit 0 / 1) for the sign parity
            const unsigned l0 = gl | ((si & 1u) << b0) | (((si >> 1) & 1u) << b1) | (((si >> 2) & 1u) << b2);
            const unsigned l1 = l0 | ((third & 1u) << b0) | (((third >> 1) & 1u) << b1) | (((third >> 2) & 1u) << b2);
            const unsigned zl = rec[o].zeta_local, pb = (unsigned)__popc(base & h.zeta);
            const unsigned par0 = pb + (unsigned)__popc(l0 & zl), par1 = pb + (unsigned)__popc(l1 & zl);
            const unsigned lowbit = x3 & (0u - x3);             // lower x bit of the op
            const unsigned shape = (x3 == 3u ? 0u : (x3 == 5u ? 2u : 4u)) + ((si & lowbit) ? 0u : 1u);
            if (h.type == 3) run_op<3>(v, h, shape, par0, par1);
            else if (h.type == 6) run_op<6>(v, h, shape, par0, par1);
            else run_op<1>(v, h, shape, par0, par1);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) buf[s ^ ((k & 1) ? t0 : 0u) ^ ((k & 2) ? t1 : 0u) ^ ((k & 4) ? t2 : 0u)] = v[k];
    }
}

// The op loop of one tile: every op of the run in order, one CTA barrier between ops.  The header (pattern test, sign
// mask, 2x2 block) of op k+1 is fetched before the barrier that ends op k.
template <bool SWZ>
__device__ __forceinline__ void tile_ops(double2 *buf, double2 *ph, const TileRec *rec, const TileTerm *tterm,
                                         const unsigned short *ptab, const TileLaunch &tl, int T, int nsub, unsigned base,
                                         unsigned lomask_g, unsigned himask_g) {
    OpHdr h;
    load_hdr(&rec[0], h);
    const int nthr = (int)blockDim.x;
    // first table entry of this thread for the current op: fetched together with the header, i.e. before the barrier
    unsigned efirst = (h.type != 2 && h.reps > 0) ? (unsigned)ptab[h.off + threadIdx.x] : 0u;
    int sidx = 0, mark = 0;
    while (sidx < nsub) {
        if (mark < 40) TMA_TLMARK(4 + mark);
        ++mark;
        int step = 1;
        if (h.type != 2) {
#if FH_OPLOOP_VARIANT < 3
#ifdef FH_TILE_RUNS
            if (h.run_len >= 2) {
                fused_run<SWZ>(buf, &rec[sidx], h.run_len, h.run_bits, T, base);
                step = h.run_len;
            } else
#endif
            if ((base & h.fixmask_out) == h.fixval_out) {
                const unsigned sb = (unsigned)__popc(base & h.zeta);
                const unsigned short *tp = ptab + h.off + threadIdx.x;
                if (h.full4) {
                    if (h.type == 3) pair_block_full4<3>(buf, h, tp, nthr, sb, efirst);
                    else pair_block_full4<6>(buf, h, tp, nthr, sb, efirst);
                } else if (h.type == 3) pair_block<3>(buf, h, tp, nthr, sb, efirst);
                else if (h.type == 6) pair_block<6>(buf, h, tp, nthr, sb, efirst);
                else pair_block<1>(buf, h, tp, nthr, sb, efirst);
            }
#endif
        } else {
            apply_diag<SWZ>(buf, ph, &rec[sidx], tterm, tl, T, base, lomask_g, himask_g);
        }
        sidx += step;
        if (sidx < nsub) {
#if FH_OPLOOP_VARIANT < 4
            load_hdr(&rec[sidx], h);
            efirst = (h.type != 2 && h.reps > 0) ? (unsigned)ptab[h.off + threadIdx.x] : 0u;
#endif
            __syncthreads();
        }
    }
}

// ----------------------------------------------------------------------------------------------
// forward / dagger run on one state
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// op records, diagonal terms and pair table of one run -> shared memory by three bulk copies signalling `bar`
__device__ __forceinline__ void fetch_run_descriptors(const TileLaunch &tl, const TileRec *recs, const TileTerm *terms,
                                                      const unsigned short *ptab, TileRec *s_rec, TileTerm *s_term, unsigned short *s_tab,
                                                      unsigned bar) {
    const unsigned rb = (unsigned)tl.nsub * (unsigned)sizeof(TileRec), tb = (unsigned)tl.nterms * (unsigned)sizeof(TileTerm),
                   pb = (unsigned)tl.ptab_words * 2u;
    mbar_expect_tx(bar, rb + tb + pb);
    bulk_g2s(smem_u32(s_rec), recs + tl.first_rec, rb, bar);
    if (tb) bulk_g2s(smem_u32(s_term), terms + tl.first_term, tb, bar);
    if (pb) bulk_g2s(smem_u32(s_tab), ptab + tl.ptab_first, pb, bar);
}

template <bool PDL, bool SWZ>
__global__ void __launch_bounds__(128, 2)
    k_tile_tma(const __grid_constant__ CUtensorMap map, const TileLaunch tl, const TmaPlan plan,
               const TileRec *__restrict__ recs, const TileTerm *__restrict__ terms, const unsigned short *__restrict__ ptab, int n,
               int stages) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long full_bar[2], desc_bar;
    const int T = tl.nbits, nsub = tl.nsub;
    const unsigned L = 1u << T, tile_bytes = L * 16u;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // carve: [tile buffers, 1024-byte aligned][records][terms][pair table][phase tables]
    const unsigned raw = smem_u32(smem_raw);
    const unsigned pad = ((raw + 1023u) & ~1023u) - raw;
    unsigned char *tiles = smem_raw + pad;
    TileRec *rec = reinterpret_cast<TileRec *>(tiles + (size_t)stages * tile_bytes);
    TileTerm *tterm = reinterpret_cast<TileTerm *>(rec + nsub);
    unsigned short *stab = reinterpret_cast<unsigned short *>(tterm + tl.nterms);
    double2 *ph = reinterpret_cast<double2 *>(stab + tl.ptab_words);
    const unsigned lomask_g = tile_mask(tl, T, 0, 6), himask_g = tile_mask(tl, T, 6, TILE_BITS_CAP);
    TMA_TLMARK(0);

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&full_bar[0]), 1);
        mbar_init(smem_u32(&full_bar[1]), 1);
        mbar_init(smem_u32(&desc_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        prefetch_tensormap(&map);
        // records / terms / table were written before the previous kernel started (payload copy at the head of the
        // graph, static table): they may be fetched while that kernel is still running
        fetch_run_descriptors(tl, recs, terms, ptab, rec, tterm, stab, smem_u32(&desc_bar));
    }
    if (PDL) asm volatile("griddepcontrol.launch_dependents;");
    __syncthreads();                                                // barrier objects are initialised
    if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");     // the previous kernel's amplitudes are complete
    TMA_TLMARK(1);

    const unsigned ntiles = 1u << (n - T);
    const unsigned tiles_u32 = smem_u32(tiles);
    if (warp == 0 && blockIdx.x < ntiles)
        warp_load_tile(&map, plan, (unsigned)tile_base(tl, T, blockIdx.x), tiles_u32, smem_u32(&full_bar[0]), tile_bytes, lane);
    mbar_wait(smem_u32(&desc_bar), 0);

    unsigned it = 0;
    for (unsigned t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const unsigned st = stages == 2 ? (it & 1u) : 0u;
        const unsigned parity = stages == 2 ? ((it >> 1) & 1u) : (it & 1u);
        if (stages == 2 && warp == 0 && t + gridDim.x < ntiles) {
            bulk_wait_read0();          // the store of tile it-1 has finished reading buffer st^1
            warp_load_tile(&map, plan, (unsigned)tile_base(tl, T, t + gridDim.x), tiles_u32 + (st ^ 1u) * tile_bytes,
                           smem_u32(&full_bar[st ^ 1u]), tile_bytes, lane);
        }
        const unsigned base = (unsigned)tile_base(tl, T, t);
        double2 *buf = reinterpret_cast<double2 *>(tiles + (size_t)st * tile_bytes);
        mbar_wait(smem_u32(&full_bar[st]), parity);
        TMA_TLMARK(2);
        tile_ops<SWZ>(buf, ph, rec, tterm, stab, tl, T, nsub, base, lomask_g, himask_g);
        TMA_TLMARK(3);
        fence_async_smem();             // generic-proxy writes of the op loop -> visible to the bulk store
        __syncthreads();
        if (warp == 0) {
            warp_store_tile(&map, plan, base, tiles_u32 + st * tile_bytes, lane);
            TMA_TLMARK(62);
            if (stages == 1 && t + gridDim.x < ntiles) {
                bulk_wait_read0();
                warp_load_tile(&map, plan, (unsigned)tile_base(tl, T, t + gridDim.x), tiles_u32, smem_u32(&full_bar[0]),
                               tile_bytes, lane);
            }
        }
    }
    // shared memory must outlive the bulk stores' reads; their global writes are complete at grid end like any store
    if (warp == 0) bulk_wait_read0();
    TMA_TLMARK(63);
}

// ----------------------------------------------------------------------------------------------
// chain: several consecutive tile runs of one state in ONE cooperative launch
// ----------------------------------------------------------------------------------------------
// At 18-20 qubits the whole state is a few MiB and a tile run is a few microseconds, so a kernel boundary per run
// (drain, launch, prologue, descriptor fetch, cold op records) costs more than the run itself.  k_tile_chain keeps one
// CTA per tile resident for the whole sequence of runs: per run it waits until every tile of the previous run has been
// stored (one counter per run boundary in global memory, release/acquire at gpu scope; all CTAs are co-resident because
// the launch is cooperative), pulls its tile with TMA, applies the ops, stores it with TMA.  The descriptors of run r+1
// (geometry, tensor maps, op records) are prefetched into the alternate shared-memory buffers while run r computes.
// Extras: the first run of a forward pass can synthesise the basis state instead of loading it (no set-basis kernel),
// and any run can store its tile to a second buffer as well (the psi checkpoint of the adjoint sweep).
struct __align__(16) ChainRun {
    TileLaunch tl;
    TmaPlan plan;
    int map_index;       // tensor map of the state for this run's tile bits
    int map2_index;      // second store target, or -1
    int init_basis;      // 1: the tile is |basis> restricted to it (nothing is loaded)
    int ntiles;
    int pad[3];
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned *p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

template <typename T>
__device__ __forceinline__ void smem_copy16(T *dst, const T *src, int count) {     // count objects, sizeof(T) % 16 == 0
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
    uint4 *d4 = reinterpret_cast<uint4 *>(dst);
    const int chunks = count * (int)(sizeof(T) / 16);
    for (int c = threadIdx.x; c < chunks; c += blockDim.x) d4[c] = __ldg(s4 + c);
}

#define FH_CHAIN_MAX_RUNS 48

__global__ void __launch_bounds__(128, 2)
    k_tile_chain(const ChainRun *__restrict__ runs, int nruns, const CUtensorMap *__restrict__ maps,
                 const TileRec *__restrict__ recs, const TileTerm *__restrict__ terms, const unsigned short *__restrict__ ptab, int n,
                 unsigned *__restrict__ sync, unsigned long long basis, int rec_cap, int term_cap, int tab_cap,
                 int tile_bits_max) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long full_bar, desc_bar[2];
    __shared__ ChainRun srun[FH_CHAIN_MAX_RUNS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned raw = smem_u32(smem_raw);
    const unsigned pad = ((raw + 1023u) & ~1023u) - raw;
    unsigned char *tile = smem_raw + pad;
    unsigned char *after = tile + (16u << tile_bits_max);
    TileRec *rec_base = reinterpret_cast<TileRec *>(after);                          // two buffers of rec_cap records
    TileTerm *term_base = reinterpret_cast<TileTerm *>(rec_base + 2 * rec_cap);      // two buffers of term_cap terms
    unsigned short *tab_base = reinterpret_cast<unsigned short *>(term_base + 2 * term_cap);     // two buffers of tab_cap entries
    double2 *ph = reinterpret_cast<double2 *>(tab_base + 2 * tab_cap);
    double2 *buf = reinterpret_cast<double2 *>(tile);
    const unsigned tile_u32 = smem_u32(tile), bar = smem_u32(&full_bar);

    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(smem_u32(&desc_bar[0]), 1);
        mbar_init(smem_u32(&desc_bar[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    smem_copy16(srun, runs, nruns);
    __syncthreads();
    if (threadIdx.x == 0) {
        fetch_run_descriptors(srun[0].tl, recs, terms, ptab, rec_base, term_base, tab_base, smem_u32(&desc_bar[0]));
        prefetch_tensormap(maps + srun[0].map_index);
    }

    unsigned uses = 0;
    int prev_active = 0;                    // CTAs that stored a tile in the previous run
    for (int r = 0; r < nruns; ++r) {
        const ChainRun &R = srun[r];
        const TileLaunch &tl = R.tl;
        const int T = tl.nbits, nsub = tl.nsub;
        const unsigned L = 1u << T, tile_bytes = L * 16u;
        const bool active = blockIdx.x < (unsigned)R.ntiles;
        const unsigned base = (unsigned)tile_base(tl, T, blockIdx.x);
        const CUtensorMap *map = maps + R.map_index;
        const int pb = r & 1;
        if (threadIdx.x == 0 && r + 1 < nruns) {
            // descriptors of the next run into the other buffers (their last readers finished before the barrier that
            // ended run r-1), tensor maps of the next run towards the TMA unit
            const ChainRun &N = srun[r + 1];
            fetch_run_descriptors(N.tl, recs, terms, ptab, rec_base + (pb ^ 1) * rec_cap, term_base + (pb ^ 1) * term_cap,
                                  tab_base + (pb ^ 1) * tab_cap, smem_u32(&desc_bar[pb ^ 1]));
            prefetch_tensormap(maps + N.map_index);
            if (N.map2_index >= 0) prefetch_tensormap(maps + N.map2_index);
        }
        if (active && warp == 0 && !R.init_basis) {
            if (r > 0) {
                if (lane == 0) {
                    const unsigned want = (unsigned)prev_active;
                    while (ld_acquire_gpu(sync + (r - 1)) < want) { }
                    fence_async_all();          // the other CTAs' tile stores -> visible to this CTA's bulk loads
                }
                __syncwarp();
            }
            warp_load_tile(map, R.plan, base, tile_u32, bar, tile_bytes, lane);
        }
        mbar_wait(smem_u32(&desc_bar[pb]), (unsigned)(r >> 1) & 1u);
        if (active) {
            const TileRec *rec = rec_base + pb * rec_cap;
            const TileTerm *tterm = term_base + pb * term_cap;
            const unsigned short *stab = tab_base + pb * tab_cap;
            const unsigned lomask_g = tile_mask(tl, T, 0, 6), himask_g = tile_mask(tl, T, 6, TILE_BITS_CAP);
            const bool swz = R.plan.swizzle != 0;
            if (R.init_basis) {
                // |basis> restricted to this tile: zeros, and 1 at the tile-local index of the basis state if it is here
                const unsigned tmask = lomask_g | himask_g;
                for (unsigned l = threadIdx.x; l < L; l += blockDim.x) buf[l] = make_double2(0.0, 0.0);
                __syncthreads();
                if (threadIdx.x == 0 && ((unsigned)basis & ~tmask) == base) {
                    unsigned l = 0;
#pragma unroll
                    for (int b = 0; b < TILE_BITS_CAP; ++b)
                        if (b < T) l |= (((unsigned)basis >> tl.bits[b]) & 1u) << b;
                    buf[swz ? tslot<true>(l) : l] = make_double2(1.0, 0.0);
                }
                __syncthreads();
            } else {
                mbar_wait(bar, uses & 1u);
                ++uses;
            }
            if (swz) tile_ops<true>(buf, ph, rec, tterm, stab, tl, T, nsub, base, lomask_g, himask_g);
            else tile_ops<false>(buf, ph, rec, tterm, stab, tl, T, nsub, base, lomask_g, himask_g);
            fence_async_smem();
        }
        __syncthreads();                    // ops done (all threads)
        if (active && warp == 0) {
            warp_store_tile(map, R.plan, base, tile_u32, lane);
            if (R.map2_index >= 0) warp_store_tile(maps + R.map2_index, R.plan, base, tile_u32, lane);
            if (r + 1 < nruns) {
                bulk_wait0();               // this lane's stores are complete (global writes performed)
                fence_async_all();
                __syncwarp();
                if (lane == 0) {
                    __threadfence();
                    red_release_gpu_add(sync + r, 1u);
                }
            } else {
                bulk_wait_read0();
            }
        }
        prev_active = min((int)gridDim.x, R.ntiles);
        __syncthreads();                    // tile buffer free for the next load
    }
}

// ----------------------------------------------------------------------------------------------
// launch
// ----------------------------------------------------------------------------------------------
static size_t tile_tma_smem(const TileLaunch &tl, int stages) {
    return 1024 + (size_t)stages * (16ull << tl.nbits) + (size_t)tl.nsub * sizeof(TileRec) + (size_t)tl.nterms * sizeof(TileTerm) +
           (size_t)tl.ptab_words * 2 + 192 * sizeof(double2) + 64;
}

// Four warps per CTA: every thread owns several independent pairs of an op (latency hidden by instruction-level
// parallelism), and the per-op overhead (header fetch, barrier) is paid by four warps instead of sixteen.
int fh_tile_threads(int nbits) {
    static const int env = getenv("FHSIM_TILE_THREADS") ? atoi(getenv("FHSIM_TILE_THREADS")) : 0;
    if (env >= 32 && env <= 128 && (env & (env - 1)) == 0) return env;     // the kernels are built for <= 128 threads
    return nbits >= 8 ? 128 : 64;
}

// -1: the TMA kernels do not take this tile (launch_tile falls back to the register-staged kernel); else the layout
int fh_tile_tma_layout(const unsigned char *bits, int T, int n) {
    if (n > 30 || T < 3 || T > FH_MAX_TILE_BITS || bits[0] != 0) return -1;
    std::vector<std::pair<int, int>> dims;
    TmaPlan plan;
    plan_dims(bits, T, true, dims, plan);
    if (plan.swizzle && plan.map_bits < 6) plan_dims(bits, T, false, dims, plan);
    if (plan.n_extra > 6) return -1;
    return plan.swizzle;
}

void launch_tile(cudaStream_t s, double2 *psi, const TileLaunch &tl, const TileRec *d_recs, const TileTerm *d_terms,
                 const unsigned short *d_ptab, int n) {
    const bool force_ldg = getenv("FHSIM_TILE_LDG") != nullptr;      // read per call: tests flip it at run time
    const TmaEntry *e = force_ldg ? nullptr : tma_entry(psi, n, tl);
    if (!e || !d_ptab) {
        launch_tile_ldg(s, psi, tl, d_recs, d_terms, n);
        return;
    }
    const int nbits = tl.nbits;
    const unsigned long long ntiles = 1ull << (n - nbits);
    int dev = 0, sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
    // one tile per CTA while the tiles fit the machine in one wave of <= 2 CTAs per SM; otherwise persistent CTAs with
    // two tile buffers each (load of the next tile and store of the previous one overlap the op loop)
    int stages = 1;
    unsigned long long grid = ntiles;
    const size_t smem1 = tile_tma_smem(tl, 1), smem2 = tile_tma_smem(tl, 2);
    const size_t sm_budget = 224 * 1024;
    if (smem1 > sm_budget) {
        launch_tile_ldg(s, psi, tl, d_recs, d_terms, n);
        return;
    }
    if (ntiles > (unsigned long long)sm * 2) {
        // more tiles than one wave: persistent CTAs.  Several CTAs per SM with one buffer each (the siblings' op loops
        // hide this CTA's tile traffic) if at least two fit, else one CTA per SM with two buffers (load of tile i+1 /
        // store of tile i-1 under the op loop of tile i), else one buffer
        int per_sm = (int)(sm_budget / (smem1 + 1024));
        if (per_sm > 2) per_sm = 2;                 // registers: two 128-thread CTAs per SM
        if (per_sm >= 2) {
            grid = (unsigned long long)sm * per_sm;
        } else if (smem2 <= sm_budget) {
            stages = 2;
            grid = (unsigned long long)sm;
        } else {
            grid = (unsigned long long)sm;
        }
        if (grid > ntiles) grid = ntiles;
    }
    const size_t smem = stages == 2 ? smem2 : smem1;
    const int threads = fh_tile_threads(nbits);
    static const bool pdl_all = getenv("FHSIM_PDL") != nullptr, pdl_off = getenv("FHSIM_NO_PDL") != nullptr;
    const bool pdl = !pdl_off && (pdl_all || g_fh_tile_pdl_scope > 0);
    ++g_fh_launch_count;
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
    cfg.numAttrs = pdl ? 1 : 0;
    if (pdl) {
        if (e->plan.swizzle) cudaLaunchKernelEx(&cfg, k_tile_tma<true, true>, e->map, tl, e->plan, d_recs, d_terms, d_ptab, n, stages);
        else cudaLaunchKernelEx(&cfg, k_tile_tma<true, false>, e->map, tl, e->plan, d_recs, d_terms, d_ptab, n, stages);
    } else {
        if (e->plan.swizzle) cudaLaunchKernelEx(&cfg, k_tile_tma<false, true>, e->map, tl, e->plan, d_recs, d_terms, d_ptab, n, stages);
        else cudaLaunchKernelEx(&cfg, k_tile_tma<false, false>, e->map, tl, e->plan, d_recs, d_terms, d_ptab, n, stages);
    }
}

int fh_tile_tma_init_device() {
    const int bytes = 224 * 1024;
    FH_CUDA(cudaFuncSetAttribute(k_tile_tma<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    FH_CUDA(cudaFuncSetAttribute(k_tile_tma<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    FH_CUDA(cudaFuncSetAttribute(k_tile_tma<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    FH_CUDA(cudaFuncSetAttribute(k_tile_tma<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    FH_CUDA(cudaFuncSetAttribute(k_tile_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));   // + ~5 KB static
    return FH_OK;
}

// ----------------------------------------------------------------------------------------------
// chain launch (host)
// ----------------------------------------------------------------------------------------------
struct ChainRunHost {
    TileLaunch tl;
    int second_store;     // 1: also store this run's tile to `psi2`
    int init_basis;
};

// Plans one cooperative launch for `nruns` consecutive tile runs on `psi`.  Writes the ChainRun records and the tensor
// maps into host staging memory (`h_runs`, `h_maps`: the caller copies them to `d_runs`, `d_maps` on the stream BEFORE
// the launch -- inside a captured graph that copy is one more memcpy node).  Returns the number of maps used, or -1
// when the chain does not apply (TMA path unavailable, or more tiles than co-resident CTAs).
int plan_tile_chain(int sm, double2 *psi, double2 *psi2, const ChainRunHost *hruns, int nruns, int n, void *h_runs_v,
                    void *h_maps_v, int map_base, int *grid_out, size_t *smem_out, int *rec_cap_out, int *term_cap_out,
                    int *tab_cap_out, int *tbits_out) {
    ChainRun *h_runs = reinterpret_cast<ChainRun *>(h_runs_v);
    CUtensorMap *h_maps = reinterpret_cast<CUtensorMap *>(h_maps_v);
    // measured at 18 qubits (profiles/r02_*): the chain is not faster than one programmatic-dependent launch per run (the
    // run-boundary counter costs what the overlapped launch does), so it is opt-in; the tests run both
    if (getenv("FHSIM_TILE_LDG") || getenv("FHSIM_NO_CHAIN") || !getenv("FHSIM_CHAIN")) return -1;
    int nmaps = 0, grid = 0, rec_cap = 1, term_cap = 1, tab_cap = 8, tbits = 0;
    if (nruns > FH_CHAIN_MAX_RUNS) return -1;
    for (int r = 0; r < nruns; ++r) {
        const TileLaunch &tl = hruns[r].tl;
        const TmaEntry *e = tma_entry(psi, n, tl);
        if (!e) return -1;
        ChainRun cr;
        memset(&cr, 0, sizeof(cr));
        cr.tl = tl;
        cr.plan = e->plan;
        cr.map_index = map_base + nmaps;
        h_maps[nmaps++] = e->map;
        cr.map2_index = -1;
        if (hruns[r].second_store) {
            const TmaEntry *e2 = tma_entry(psi2, n, tl);
            if (!e2) return -1;
            cr.map2_index = map_base + nmaps;
            h_maps[nmaps++] = e2->map;
        }
        cr.init_basis = hruns[r].init_basis;
        cr.ntiles = 1 << (n - tl.nbits);
        if (n - tl.nbits > 12) return -1;
        h_runs[r] = cr;
        grid = cr.ntiles > grid ? cr.ntiles : grid;
        rec_cap = tl.nsub > rec_cap ? tl.nsub : rec_cap;
        term_cap = tl.nterms > term_cap ? tl.nterms : term_cap;
        tab_cap = tl.ptab_words > tab_cap ? tl.ptab_words : tab_cap;
        tbits = tl.nbits > tbits ? tl.nbits : tbits;
    }
    const size_t smem = 1024 + (16ull << tbits) + 2 * (size_t)rec_cap * sizeof(TileRec) + 2 * (size_t)term_cap * sizeof(TileTerm) +
                        2 * (size_t)tab_cap * 2 + 192 * sizeof(double2) + 64;
    if (smem > 200 * 1024) return -1;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tile_chain, fh_tile_threads(tbits), smem) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    if ((long long)per_sm * sm < grid) return -1;
    *grid_out = grid;
    *smem_out = smem;
    *rec_cap_out = rec_cap;
    *term_cap_out = term_cap;
    *tab_cap_out = tab_cap;
    *tbits_out = tbits;
    return nmaps;
}

size_t fh_chain_run_bytes() { return sizeof(ChainRun); }
size_t fh_chain_map_bytes() { return sizeof(CUtensorMap); }

int launch_tile_chain(cudaStream_t s, const void *d_runs, int nruns, const void *d_maps, const TileRec *d_recs,
                      const TileTerm *d_terms, const unsigned short *d_ptab, int n, unsigned *d_sync, unsigned long long basis,
                      int grid, size_t smem, int rec_cap, int term_cap, int tab_cap, int tbits) {
    const ChainRun *runs = reinterpret_cast<const ChainRun *>(d_runs);
    const CUtensorMap *maps = reinterpret_cast<const CUtensorMap *>(d_maps);
    const int threads = fh_tile_threads(tbits);
    void *args[] = {(void *)&runs, (void *)&nruns, (void *)&maps, (void *)&d_recs, (void *)&d_terms, (void *)&d_ptab,
                    (void *)&n, (void *)&d_sync, (void *)&basis, (void *)&rec_cap, (void *)&term_cap, (void *)&tab_cap,
                    (void *)&tbits};
    ++g_fh_launch_count;
    const cudaError_t e = cudaLaunchCooperativeKernel((const void *)k_tile_chain, dim3((unsigned)grid), dim3((unsigned)threads),
                                                      args, smem, s);
    if (e != cudaSuccess) {
        cudaGetLastError();
        --g_fh_launch_count;
        return 0;
    }
    return 1;
}
