// sector_eval.cu -- one whole evaluation (ansatz, W, H, W^dagger, pool screening) on a SECTOR-COMPRESSED state that lives in
// the shared memory of ONE thread-block cluster.
//
// Every circuit of the ADAPT / HVA drivers (models/adapt_vqe.py:325-361, hva.py:273-303) conserves N_up and N_dn, so the
// state never leaves the (N_up, N_dn) sector of its basis state: 3x3 at (5, 4) has 126 x 126 = 15 876 amplitudes (254 KB)
// instead of 2^18 (4 MiB).  That fits the distributed shared memory of a 16-CTA cluster twice (psi and lambda), so the
// 19 latency-bound launches of the full-space path collapse into two:
//   k_sector_eval   ONE cluster.  The state is the matrix Psi[a][b]: in ROW layout a = rank of the up pattern, b = rank of the
//                   down pattern; in COL layout the roles are exchanged.  Row a lives in CTA a % C (row a / C there).  A pair
//                   op restricted to the sector is a PRODUCT of two short lists: the up patterns matching the op's up
//                   pattern bits (with partner rank and sign) times the down patterns matching its down bits -- so there is
//                   no index arithmetic and no table of 2^n anything; the lists (<= d_up + d_dn words per op, built on the
//                   host once per program) are copied into shared memory once, so the op loop makes no global access.  The
//                   owner of the "low" amplitude of a pair reads the partner (DSMEM, ld.shared::cluster, when it lives in
//                   another CTA), applies the 2x2 block and writes both back.  Ops that only touch down orbitals are
//                   CTA-local in ROW layout, up-only ops in COL layout: they end with __syncthreads instead of a cluster
//                   barrier, and the host inserts a transpose (one all-to-all through DSMEM into the spare buffer) when a
//                   run of such ops is ahead.  H psi is a gather over the x-mask groups of the observable in compact
//                   coordinates (partner through the two rank tables, four amplitudes per thread in flight), fused with
//                   <psi|H|psi>; the adjoint part runs W^dagger on lambda in place.
//   k_sector_pool   all SMs: one CTA per pool entry, pairs enumerated from the entry's two lists, psi_s / lambda_s read from
//                   the compressed global copies; fixed-order reduction, last CTA folds entries into outputs.
// Results equal the full-space path to rounding (same arithmetic on the non-zero amplitudes, other summation order).
//
// Measured on B200 (profiles/r02_sector_*): a 16-CTA cluster barrier costs ~950 cycles (555 for 2 CTAs) and an FP64 warp
// instruction 4-6 issue cycles per SM sub-partition, so an op that exchanges amplitudes between CTAs costs ~1 700 cycles and
// a CTA-local one ~1 100: at 3x3 (131 steps) the kernel takes 0.19 ms against 0.17 ms for the 19-launch full-space path,
// at 2x3 0.081 against 0.113 ms.  fh_program_evaluate therefore takes this path by itself only for sectors of <= 2 048
// amplitudes; FHSIM_SECTOR=1 takes it whenever it applies, FHSIM_NO_SECTOR=1 never.
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "common.cuh"
#include "sector_eval.cuh"

extern long long g_fh_launch_count;
extern thread_local int g_fh_tile_pdl_scope;      // > 0 inside fh_program_evaluate: consecutive kernels of the graph use PDL

// Programmatic dependent launch for the small kernels of the sector tail: the next kernel's CTAs become resident while this one
// drains, and wait for its memory right here.  Without the launch attribute both instructions are no-ops.
#define SEC_PDL_PROLOGUE()                                      \
    do {                                                        \
        asm volatile("griddepcontrol.launch_dependents;");      \
        asm volatile("griddepcontrol.wait;" ::: "memory");      \
    } while (0)

template <typename... KArgs, typename... Args>
static void sec_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    // measured on the 18-qubit screening: PDL between these small kernels makes the step SLOWER (0.132 vs 0.124 ms: the early
    // CTAs of the dependent kernel take slots while the producer still runs), so it is opt-in (FHSIM_SECTOR_PDL=1)
    static const bool pdl_off = getenv("FHSIM_NO_PDL") != nullptr || getenv("FHSIM_SECTOR_PDL") == nullptr;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (!pdl_off && g_fh_tile_pdl_scope > 0) ? 1 : 0;
    ++g_fh_launch_count;
    cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

#define SEC_MAX_C 16
#define SEC_THREADS 256
#define SEC_FB 14            // bits of the rank fields of a list entry: rank | partner << SEC_FB | sign << 31
#define SEC_FM 0x3fffu
#define SEC_MAX_D 16384     // patterns per spin

enum { SV_PAIR = 1, SV_DIAG = 2, SV_TRANSPOSE = 3, SV_CHECKPOINT = 4, SV_HAPPLY = 5, SV_STORE = 6 };
enum { SF_DAGGER = 1, SF_COL = 2, SF_CLUSTER_AFTER = 4, SF_WANT_LAMBDA = 8, SF_LOCAL = 16 };

struct __align__(16) SecVOp {      // 96 bytes
    int kind;
    int index;                     // pair: PairOp index; diag: first DiagTerm index
    int nterms;                    // diag
    int flags;
    unsigned listU, listD;         // first word of the op's up / down list
    unsigned short offU[SEC_MAX_C + 1], offD[SEC_MAX_C + 1];      // owner segments (cyclic: owner = rank % C)
    int slot;                      // pair: index of this vop's list slot in shared memory
};
static_assert(sizeof(SecVOp) == 96, "SecVOp layout");

struct __align__(16) SecGroup {    // 32 bytes: TabGroup in compact coordinates (up orbital b -> bit b, down orbital b -> bit 16+b)
    unsigned x;
    int first_class, n_class;
    unsigned live;
    unsigned char pos[4];          // compact bit positions of the x bits; unused = 31 (always 0: half <= 15)
    int kbits;
    unsigned zeta1;                // single-class groups (every hopping group): the class's zeta and value offset inline, so the
    int vofs1;                     // gather needs no class record
};
struct __align__(8) SecClass {
    unsigned zeta;
    int vofs;
};
struct __align__(8) SecTerm {      // 24 bytes: one diagonal term of a vop, dagger already folded into s
    double c, s;
    unsigned zc, pad;
};

struct SecArgs {
    int half, C, logC;
    unsigned d_up, d_dn;
    unsigned S;                    // amplitudes per CTA buffer
    const unsigned short *cfgU, *cfgD, *rankU, *rankD;
    const SecVOp *vops;
    int nvops, nterms_total;
    const int *term_first;         // [nvops]: first SecTerm of a diag vop in the kernel's term array
    const unsigned *term_zc;       // compact z of every DiagTerm of the program
    const unsigned *lists;
    const PairOp *pairs;
    const DiagTerm *dterms;
    const SecGroup *groups;
    const SecClass *classes;
    const double2 *vals;
    const double2 *hdiag;
    int ngroups, nclasses, nvals;
    double2 *chk, *lam_out;
    double *res;
    unsigned basis_ru, basis_rd;
    int layout0;
    unsigned la_max, lb_max;       // list slot geometry: la_max + lb_max words per pair vop
    int npair_vops;
    long long *timeline;           // debugging (FHSIM_SECTOR_TIMELINE): clock64 of CTA 0 at every vop, or NULL
};

// ----------------------------------------------------------------------------------------------
// device helpers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned sec_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned sec_mapa(unsigned addr, unsigned cta) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
    return r;
}
__device__ __forceinline__ double2 sec_ld_cluster(unsigned addr) {
    double2 v;
    asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sec_st_cluster(unsigned addr, double2 v) {
    asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void sec_st_cluster_f64(unsigned addr, double v) {
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void sec_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ double2 sec_cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double sec_flip(double v, unsigned sbit) {     // sbit = 0 or 0x80000000: v or -v
    return __hiloint2double(__double2hiint(v) ^ (int)sbit, __double2loint(v));
}

// fixed-order block sum of two doubles (256 threads); result valid in thread 0
__device__ __forceinline__ void sec_block_sum2(double &a, double &b, double *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, o);
        b += __shfl_down_sync(0xffffffffu, b, o);
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) {
        red[2 * w] = a;
        red[2 * w + 1] = b;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sa = 0.0, sb = 0.0;
        for (int k = 0; k < SEC_THREADS / 32; ++k) {
            sa += red[2 * k];
            sb += red[2 * k + 1];
        }
        a = sa;
        b = sb;
    }
    __syncthreads();
}

// ----------------------------------------------------------------------------------------------
// the cluster kernel
// ----------------------------------------------------------------------------------------------
// predicated loads; `tag` (the vop index) is an unused input that keeps the compiler from merging loads of different vops:
// the asm statements are otherwise pure, so loads of independent pairs / amplitudes can be issued back to back
__device__ __forceinline__ double2 sec_ld_cluster_if(unsigned addr, unsigned pred, int tag) {
    double2 v;
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\tmov.f64 %1, 0d0000000000000000;\n\t"
        "@p ld.shared::cluster.v2.f64 {%0, %1}, [%2];\n\t}"
        : "=d"(v.x), "=d"(v.y)
        : "r"(addr), "r"(pred), "r"(tag));
    return v;
}
__device__ __forceinline__ double2 sec_ld_shared_if(unsigned addr, unsigned pred, int tag) {
    double2 v;
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\tmov.f64 %1, 0d0000000000000000;\n\t"
        "@p ld.shared.v2.f64 {%0, %1}, [%2];\n\t}"
        : "=d"(v.x), "=d"(v.y)
        : "r"(addr), "r"(pred), "r"(tag));
    return v;
}

struct SecMat {
    double2 m00, m01, m10, m11;
};

__device__ __forceinline__ void sec_pair_math(const SecMat &M, unsigned sbit, double2 x, double2 y, double2 &nx, double2 &ny) {
    const double2 t01 = make_double2(sec_flip(M.m01.x, sbit), sec_flip(M.m01.y, sbit));
    const double2 t10 = make_double2(sec_flip(M.m10.x, sbit), sec_flip(M.m10.y, sbit));
    nx.x = M.m00.x * x.x - M.m00.y * x.y + t01.x * y.x - t01.y * y.y;
    nx.y = M.m00.x * x.y + M.m00.y * x.x + t01.x * y.y + t01.y * y.x;
    ny.x = t10.x * x.x - t10.y * x.y + M.m11.x * y.x - M.m11.y * y.y;
    ny.y = t10.x * x.y + t10.y * x.x + M.m11.x * y.y + M.m11.y * y.x;
}

// decode pair p of the product list: shared-memory byte offsets of the two amplitudes, partner CTA, sign bit
__device__ __forceinline__ void sec_pair_decode(const unsigned *LA, const unsigned *LB, unsigned p, unsigned nB, unsigned dB, int logC,
                                                unsigned cmask, unsigned &offi, unsigned &offj, unsigned &ctaj, unsigned &sbit) {
    const unsigned ai = p / nB, bi = p - ai * nB;
    const unsigned ea = LA[ai], eb = LB[bi];
    const unsigned a = ea & SEC_FM, ap = (ea >> SEC_FB) & SEC_FM, b = eb & SEC_FM, bp = (eb >> SEC_FB) & SEC_FM;
    sbit = (ea ^ eb) & 0x80000000u;
    offi = ((a >> logC) * dB + b) * 16u;
    offj = ((ap >> logC) * dB + bp) * 16u;
    ctaj = ap & cmask;
}

// all pairs of one op owned by this CTA, two independent pairs per thread and trip.  LOCAL: the partner is in this CTA
template <bool LOCAL>
__device__ __forceinline__ void sec_pair_apply(unsigned cur_u32, const unsigned *LA, const unsigned *LB, unsigned nA, unsigned nB,
                                               unsigned dB, int logC, unsigned cmask, const SecMat &M, int tag) {
    const unsigned P = nA * nB;
    for (unsigned p0 = threadIdx.x; p0 < P; p0 += 2 * SEC_THREADS) {
        const unsigned p1 = p0 + SEC_THREADS;
        const unsigned has1 = p1 < P ? 1u : 0u;
        unsigned oi0, oj0, cj0, sb0, oi1, oj1, cj1, sb1;
        sec_pair_decode(LA, LB, p0, nB, dB, logC, cmask, oi0, oj0, cj0, sb0);
        sec_pair_decode(LA, LB, has1 ? p1 : p0, nB, dB, logC, cmask, oi1, oj1, cj1, sb1);
        const unsigned aj0 = LOCAL ? cur_u32 + oj0 : sec_mapa(cur_u32 + oj0, cj0);
        const unsigned aj1 = LOCAL ? cur_u32 + oj1 : sec_mapa(cur_u32 + oj1, cj1);
        const double2 x0 = sec_ld_shared_if(cur_u32 + oi0, 1u, tag);
        const double2 y0 = LOCAL ? sec_ld_shared_if(aj0, 1u, tag) : sec_ld_cluster_if(aj0, 1u, tag);
        const double2 x1 = sec_ld_shared_if(cur_u32 + oi1, has1, tag);
        const double2 y1 = LOCAL ? sec_ld_shared_if(aj1, has1, tag) : sec_ld_cluster_if(aj1, has1, tag);
        double2 nx0, ny0, nx1, ny1;
        sec_pair_math(M, sb0, x0, y0, nx0, ny0);
        sec_pair_math(M, sb1, x1, y1, nx1, ny1);
        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(cur_u32 + oi0), "d"(nx0.x), "d"(nx0.y) : "memory");
        if (LOCAL) asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(aj0), "d"(ny0.x), "d"(ny0.y) : "memory");
        else sec_st_cluster(aj0, ny0);
        if (has1) {
            asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(cur_u32 + oi1), "d"(nx1.x), "d"(nx1.y) : "memory");
            if (LOCAL) asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(aj1), "d"(ny1.x), "d"(ny1.y) : "memory");
            else sec_st_cluster(aj1, ny1);
        }
    }
}

__global__ void __launch_bounds__(SEC_THREADS, 1) k_sector_eval(const __grid_constant__ SecArgs A) {
    extern __shared__ __align__(16) unsigned char sm[];
    __shared__ double red[2 * SEC_THREADS / 32];
    __shared__ double cpart[2 * SEC_MAX_C];
    unsigned cta;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta));
    const int C = A.C, logC = A.logC;
    const unsigned cmask = (unsigned)(C - 1);
    const unsigned tid = threadIdx.x;

    // carve
    double2 *bufs = reinterpret_cast<double2 *>(sm);                           // two buffers of S amplitudes
    SecVOp *vops = reinterpret_cast<SecVOp *>(bufs + 2 * (size_t)A.S);
    double *mats = reinterpret_cast<double *>(vops + A.nvops);                 // 8 doubles per vop
    SecTerm *terms = reinterpret_cast<SecTerm *>(mats + 8 * (size_t)A.nvops);
    int *term_first = reinterpret_cast<int *>(terms + ((A.nterms_total + 1) & ~1));
    SecGroup *groups = reinterpret_cast<SecGroup *>(term_first + ((A.nvops + 3) & ~3));
    SecClass *classes = reinterpret_cast<SecClass *>(groups + A.ngroups);
    double2 *vals = reinterpret_cast<double2 *>(classes + ((A.nclasses + 1) & ~1));
    double2 *ptab = vals + A.nvals;                                            // diag phase tables: [R_max][d_max]
    const unsigned dmax = A.d_up > A.d_dn ? A.d_up : A.d_dn;
    const unsigned rmax = (dmax + C - 1) >> logC;
    uint4 *cdesc = reinterpret_cast<uint4 *>(ptab + rmax + dmax);              // per vop, this CTA's view: one LDS.128 per op
    unsigned *lst = reinterpret_cast<unsigned *>(cdesc + A.nvops + 1);         // [npair_vops][la_max + lb_max]
    unsigned short *cfgU = reinterpret_cast<unsigned short *>(lst + (size_t)A.npair_vops * (A.la_max + A.lb_max));
    unsigned short *cfgD = cfgU + ((A.d_up + 7) & ~7u);
    unsigned short *rankU = cfgD + ((A.d_dn + 7) & ~7u);
    unsigned short *rankD = rankU + (1u << A.half);
    const unsigned lw = A.la_max + A.lb_max;

    // ---- one-time copies: descriptors, geometry, theta-dependent payload (already daggered where needed) ----
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(A.vops);
        uint4 *dst = reinterpret_cast<uint4 *>(vops);
        for (unsigned t = tid; t < (unsigned)A.nvops * (sizeof(SecVOp) / 16); t += SEC_THREADS) dst[t] = __ldg(src + t);
        for (unsigned t = tid; t < (unsigned)A.nvops; t += SEC_THREADS) term_first[t] = __ldg(A.term_first + t);
        for (unsigned t = tid; t < A.d_up; t += SEC_THREADS) cfgU[t] = __ldg(A.cfgU + t);
        for (unsigned t = tid; t < A.d_dn; t += SEC_THREADS) cfgD[t] = __ldg(A.cfgD + t);
        for (unsigned t = tid; t < (1u << A.half); t += SEC_THREADS) {
            rankU[t] = __ldg(A.rankU + t);
            rankD[t] = __ldg(A.rankD + t);
        }
        const uint4 *gs = reinterpret_cast<const uint4 *>(A.groups);
        uint4 *gd = reinterpret_cast<uint4 *>(groups);
        for (unsigned t = tid; t < (unsigned)A.ngroups * 2u; t += SEC_THREADS) gd[t] = __ldg(gs + t);
        for (unsigned t = tid; t < (unsigned)A.nclasses; t += SEC_THREADS) classes[t] = A.classes[t];
        for (unsigned t = tid; t < (unsigned)A.nvals; t += SEC_THREADS) vals[t] = A.vals[t];
    }
    __syncthreads();
    // the two lists of every pair vop (this CTA's segment of the distributed one, the whole of the other), one warp per vop:
    // everything the op loop reads is in shared memory from here on -- no global access between the barriers
    for (int v = (int)(tid >> 5); v < A.nvops; v += SEC_THREADS / 32) {
        const SecVOp &op = vops[v];
        if (op.kind != SV_PAIR) continue;
        unsigned *slot = lst + (size_t)op.slot * lw;
        const bool col = (op.flags & SF_COL) != 0;
        const unsigned short *offA = col ? op.offD : op.offU, *offB = col ? op.offU : op.offD;
        const unsigned baseA = col ? op.listD : op.listU, baseB = col ? op.listU : op.listD;
        const unsigned a0 = offA[cta], nA = offA[cta + 1] - a0, nB = offB[C];
        for (unsigned t = tid & 31u; t < nA + nB; t += 32u)
            slot[t < nA ? t : A.la_max + (t - nA)] = __ldg(A.lists + (t < nA ? baseA + a0 + t : baseB + (t - nA)));
    }
    for (unsigned v = tid; v <= (unsigned)A.nvops; v += SEC_THREADS) {
        if (v == (unsigned)A.nvops) {
            cdesc[v] = make_uint4(0u, 0u, 0u, 0u);
            continue;
        }
        const SecVOp &op = vops[v];
        uint4 cd = make_uint4((unsigned)op.kind | ((unsigned)op.flags << 8), 0u, 0u, 0u);
        if (op.kind == SV_PAIR) {
            const bool col = (op.flags & SF_COL) != 0;
            const unsigned short *offA = col ? op.offD : op.offU, *offB = col ? op.offU : op.offD;
            cd.y = (unsigned)(offA[cta + 1] - offA[cta]) | ((unsigned)offB[C] << 16);
            cd.z = (unsigned)op.slot;
            const double *m = A.pairs[op.index].m;
            double *d = mats + 8 * (size_t)v;
            if (op.flags & SF_DAGGER) {          // [[conj m00, conj m10], [conj m01, conj m11]]
                d[0] = m[0]; d[1] = -m[1];
                d[2] = m[4]; d[3] = -m[5];
                d[4] = m[2]; d[5] = -m[3];
                d[6] = m[6]; d[7] = -m[7];
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) d[k] = m[k];
            }
        } else if (op.kind == SV_DIAG) {
            // terms with bits in one species only first (they go into the two factor tables), the others behind them
            const int t0 = term_first[v];
            int w = 0;
            for (int pass = 0; pass < 2; ++pass)
                for (int m = 0; m < op.nterms; ++m) {
                    const unsigned zc = A.term_zc[op.index + m];
                    const bool mixed = (zc & 0xffffu) != 0u && (zc >> 16) != 0u;
                    if (mixed != (pass == 1)) continue;
                    const DiagTerm &t = A.dterms[op.index + m];
                    SecTerm st;
                    st.c = t.c;
                    st.s = (op.flags & SF_DAGGER) ? -t.s : t.s;
                    st.zc = zc;
                    st.pad = 0;
                    terms[t0 + w++] = st;
                    if (pass == 0) cd.y = (unsigned)w;          // number of single-species terms
                }
            cd.z = (unsigned)op.nterms;
            cd.w = (unsigned)t0;
        }
        cdesc[v] = cd;
    }

    // ---- |basis> ----
    unsigned curS = 0, othS = A.S;    // first amplitude of the current / spare buffer in bufs
    int layout = A.layout0;          // 0 ROW (a = up rank), 1 COL (a = down rank)
    for (unsigned q = tid; q < A.S; q += SEC_THREADS) bufs[q] = make_double2(0.0, 0.0);
    __syncthreads();
    if (tid == 0) {
        const unsigned a = layout ? A.basis_rd : A.basis_ru, b = layout ? A.basis_ru : A.basis_rd;
        const unsigned dB = layout ? A.d_up : A.d_dn;
        if ((a & cmask) == cta) bufs[(a >> logC) * dB + b] = make_double2(1.0, 0.0);
    }
    sec_cluster_sync();              // every CTA of the cluster is running and initialised before any remote access

    double e_re = 0.0, e_im = 0.0;
    const unsigned bufs_u32 = sec_smem_u32(bufs);
    uint4 cd = cdesc[0];
    for (int v = 0; v < A.nvops; ++v) {
        const bool tl = A.timeline && cta == 0 && tid == 0;
        if (tl) A.timeline[4 * v] = clock64();
        const uint4 cdn = cdesc[v + 1];          // the next op's descriptor is in registers before this op's barrier
        const unsigned kind = cd.x & 0xffu, flags = cd.x >> 8;
        const unsigned dA = layout ? A.d_dn : A.d_up, dB = layout ? A.d_up : A.d_dn;
        const unsigned R = (dA + C - 1) >> logC;
        const unsigned cur_u32 = bufs_u32 + curS * 16u;
        double2 *cur = bufs + curS, *oth = bufs + othS;
        if (kind == SV_PAIR) {
            const unsigned nA = cd.y & 0xffffu, nB = cd.y >> 16;
            const unsigned *LA = lst + (size_t)cd.z * lw, *LB = LA + A.la_max;
            const double2 *mp = reinterpret_cast<const double2 *>(mats + 8 * (size_t)v);
            SecMat M;
            M.m00 = mp[0];
            M.m01 = mp[1];
            M.m10 = mp[2];
            M.m11 = mp[3];
            if (tl) A.timeline[4 * v + 1] = clock64() + (long long)(M.m00.x == 123.0);
            if (flags & SF_LOCAL) sec_pair_apply<true>(cur_u32, LA, LB, nA, nB, dB, logC, cmask, M, v);
            else sec_pair_apply<false>(cur_u32, LA, LB, nA, nB, dB, logC, cmask, M, v);
        } else if (kind == SV_DIAG) {
            // phase(a, b) = PA[a] * PB[b] * (terms with bits in both species, per amplitude)
            const SecTerm *T = terms + cd.w;
            const int npure = (int)cd.y, nt = (int)cd.z;
            for (unsigned q = tid; q < R + dB; q += SEC_THREADS) {
                const bool isA = q < R;
                const unsigned idx = isA ? ((q << logC) + cta) : (q - R);
                double2 f = make_double2(1.0, 0.0);
                if (!isA || idx < dA) {
                    const bool up_table = isA ? (layout == 0) : (layout != 0);
                    const unsigned part = up_table ? (unsigned)cfgU[idx] : (unsigned)cfgD[idx];
                    for (int mth = 0; mth < npure; ++mth) {
                        const unsigned z = T[mth].zc;
                        const unsigned zmine = up_table ? (z & 0xffffu) : (z >> 16), zother = up_table ? (z >> 16) : (z & 0xffffu);
                        if (zother != 0u) continue;                   // the other table's term
                        if (zmine == 0u && !isA) continue;           // plain phase: counted once, in the row table
                        const double sg = (__popc(part & zmine) & 1) ? T[mth].s : -T[mth].s;
                        f = sec_cmul(f, make_double2(T[mth].c, sg));
                    }
                }
                ptab[q] = f;
            }
            __syncthreads();
            for (unsigned q = tid; q < R * dB; q += SEC_THREADS) {
                const unsigned lr = q / dB, b = q - lr * dB, a = (lr << logC) + cta;
                if (a >= dA) continue;
                double2 f = sec_cmul(ptab[lr], ptab[R + b]);
                if (nt > npure) {
                    const unsigned cu = layout ? (unsigned)cfgU[b] : (unsigned)cfgU[a], cd2 = layout ? (unsigned)cfgD[a] : (unsigned)cfgD[b];
                    const unsigned cfg = cu | (cd2 << 16);
                    for (int mth = npure; mth < nt; ++mth) {
                        const double sg = (__popc(cfg & T[mth].zc) & 1) ? T[mth].s : -T[mth].s;
                        f = sec_cmul(f, make_double2(T[mth].c, sg));
                    }
                }
                cur[q] = sec_cmul(cur[q], f);
            }
        } else if (kind == SV_TRANSPOSE) {
            const unsigned oth_u32 = bufs_u32 + othS * 16u;
            for (unsigned q = tid; q < R * dB; q += SEC_THREADS) {
                const unsigned lr = q / dB, b = q - lr * dB, a = (lr << logC) + cta;
                if (a >= dA) continue;
                sec_st_cluster(sec_mapa(oth_u32 + ((b >> logC) * dA + a) * 16u, b & cmask), cur[q]);
            }
            const unsigned t = curS;
            curS = othS;
            othS = t;
            layout ^= 1;
        } else if (kind == SV_CHECKPOINT || kind == SV_STORE) {
            double2 *out = kind == SV_CHECKPOINT ? A.chk : A.lam_out;
            for (unsigned q = tid; q < R * dB; q += SEC_THREADS) {
                const unsigned lr = q / dB, b = q - lr * dB, a = (lr << logC) + cta;
                if (a >= dA) continue;
                const unsigned r = layout ? b * A.d_dn + a : a * A.d_dn + b;
                out[r] = cur[q];
            }
        } else if (kind == SV_HAPPLY) {
            // gather over the x-mask groups, four amplitudes per thread in flight; groups whose x-mask stays inside the
            // undistributed species read their partner from this CTA's shared memory, the others through DSMEM
            const bool want_lambda = (flags & SF_WANT_LAMBDA) != 0;
            const unsigned nq = R * dB;
            for (unsigned q0 = tid; q0 < nq; q0 += 4 * SEC_THREADS) {
                unsigned cfgk[4], okk[4];
                double2 selfk[4];
                double ar[4], ai[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const unsigned q = q0 + (unsigned)k * SEC_THREADS;
                    const unsigned qq = q < nq ? q : q0;
                    const unsigned lr = qq / dB, b = qq - lr * dB, a = (lr << logC) + cta;
                    okk[k] = (q < nq && a < dA) ? 1u : 0u;
                    const unsigned aa = a < dA ? a : 0u;
                    const unsigned ru = layout ? b : aa, rd = layout ? aa : b;
                    cfgk[k] = (unsigned)cfgU[ru] | ((unsigned)cfgD[rd] << 16);
                    selfk[k] = cur[qq];
                    const double2 hd = __ldg(A.hdiag + (size_t)ru * A.d_dn + rd);
                    ar[k] = hd.x * selfk[k].x - hd.y * selfk[k].y;
                    ai[k] = hd.x * selfk[k].y + hd.y * selfk[k].x;
                }
                for (int g = 0; g < A.ngroups; ++g) {
                    const SecGroup G = groups[g];
                    const bool glocal = (layout ? (G.x >> 16) : (G.x & 0xffffu)) == 0u;
                    double wr[4], wi[4];
                    double2 pv[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const unsigned j = cfgk[k] ^ G.x;
                        const unsigned pat = ((j >> G.pos[0]) & 1u) | (((j >> G.pos[1]) & 1u) << 1) | (((j >> G.pos[2]) & 1u) << 2) |
                                             (((j >> G.pos[3]) & 1u) << 3);
                        const unsigned live = okk[k] & ((G.live >> pat) & 1u);
                        double w_r = 0.0, w_i = 0.0;
                        for (int c = G.first_class; c < G.first_class + G.n_class; ++c) {
                            const double2 w = vals[classes[c].vofs + pat];
                            const bool neg = (__popc(j & classes[c].zeta) & 1) != 0;
                            w_r += neg ? -w.x : w.x;
                            w_i += neg ? -w.y : w.y;
                        }
                        wr[k] = w_r;
                        wi[k] = w_i;
                        const unsigned qu = rankU[j & 0xffffu], qd = rankD[j >> 16];
                        const unsigned pa = layout ? qd : qu, pb = layout ? qu : qd;
                        const unsigned off = (((pa >> logC) * dB + pb) & 0xfffffu) * 16u;
                        pv[k] = glocal ? sec_ld_shared_if(cur_u32 + off, live, v) : sec_ld_cluster_if(sec_mapa(cur_u32 + off, pa & cmask), live, v);
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        ar[k] += wr[k] * pv[k].x - wi[k] * pv[k].y;
                        ai[k] += wr[k] * pv[k].y + wi[k] * pv[k].x;
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (!okk[k]) continue;
                    const unsigned q = q0 + (unsigned)k * SEC_THREADS;
                    if (want_lambda) oth[q] = make_double2(ar[k], ai[k]);
                    e_re += selfk[k].x * ar[k] + selfk[k].y * ai[k];          // conj(self) * (H psi)
                    e_im += selfk[k].x * ai[k] - selfk[k].y * ar[k];
                }
            }
            sec_block_sum2(e_re, e_im, red);
            if (tid == 0) {
                const unsigned base = sec_mapa(sec_smem_u32(cpart), 0u);
                sec_st_cluster_f64(base + (2u * cta) * 8u, e_re);
                sec_st_cluster_f64(base + (2u * cta + 1u) * 8u, e_im);
            }
            if (want_lambda) {
                const unsigned t = curS;
                curS = othS;
                othS = t;
            }
        }
        if (tl) A.timeline[4 * v + 2] = clock64();
        if (flags & SF_CLUSTER_AFTER) sec_cluster_sync();
        else __syncthreads();
        if (tl) A.timeline[4 * v + 3] = clock64();
        if (kind == SV_HAPPLY && cta == 0 && tid == 0) {
            double sr = 0.0, si = 0.0;
            for (int c = 0; c < C; ++c) {
                sr += cpart[2 * c];
                si += cpart[2 * c + 1];
            }
            A.res[0] = sr;
            A.res[1] = si;
        }
        cd = cdn;
    }
    if (A.timeline && cta == 0 && tid == 0) A.timeline[4 * A.nvops] = clock64();
    sec_cluster_sync();              // no CTA leaves while its shared memory may still be addressed
}

// ----------------------------------------------------------------------------------------------
// K3 on the compressed states: one CTA per pool entry
// ----------------------------------------------------------------------------------------------
struct SecPoolEntry {
    unsigned listU, nU, listD, nD;
    double br, bi;
};

__global__ void __launch_bounds__(256) k_sector_pool(const SecPoolEntry *__restrict__ entries, const unsigned *__restrict__ lists,
                                                     int e0, int e1, unsigned d_dn, const double2 *__restrict__ psi,
                                                     const double2 *__restrict__ lam, double *__restrict__ partial,
                                                     const int *__restrict__ out_first, int first_out, int count,
                                                     double *__restrict__ d_out, unsigned *__restrict__ counter) {
    SEC_PDL_PROLOGUE();
    extern __shared__ unsigned sl[];
    __shared__ double red[16];
    __shared__ unsigned is_last;
    for (int e = e0 + (int)blockIdx.x; e < e1; e += (int)gridDim.x) {
        const SecPoolEntry E = entries[e];
        __syncthreads();
        for (unsigned t = threadIdx.x; t < E.nU; t += blockDim.x) sl[t] = __ldg(lists + E.listU + t);
        for (unsigned t = threadIdx.x; t < E.nD; t += blockDim.x) sl[E.nU + t] = __ldg(lists + E.listD + t);
        __syncthreads();
        double acc = 0.0, dummy = 0.0;
        const unsigned P = E.nU * E.nD;
        for (unsigned p = threadIdx.x; p < P; p += blockDim.x) {
            const unsigned ai = p / E.nD, bi = p - ai * E.nD;
            const unsigned ea = sl[ai], eb = sl[E.nU + bi];
            const unsigned i = (ea & SEC_FM) * d_dn + (eb & SEC_FM), j = ((ea >> SEC_FB) & SEC_FM) * d_dn + ((eb >> SEC_FB) & SEC_FM);
            const double2 pi = psi[i], pj = psi[j], li = lam[i], lj = lam[j];
            // Im( conj(li) B pj + conj(lj) conj(B) pi ), sign applied afterwards
            const double2 bpj = make_double2(E.br * pj.x - E.bi * pj.y, E.br * pj.y + E.bi * pj.x);
            const double2 bpi = make_double2(E.br * pi.x + E.bi * pi.y, E.br * pi.y - E.bi * pi.x);
            const double im = (li.x * bpj.y - li.y * bpj.x) + (lj.x * bpi.y - lj.y * bpi.x);
            acc += ((ea ^ eb) & 0x80000000u) ? -im : im;
        }
        sec_block_sum2(acc, dummy, red);
        if (threadIdx.x == 0) partial[e] = 2.0 * acc;
    }
    // the last CTA folds entries into outputs (fixed order)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1u) ? 1u : 0u;
    __syncthreads();
    if (is_last) {
        __threadfence();
        for (int o = first_out + (int)threadIdx.x; o < first_out + count; o += (int)blockDim.x) {
            double s = 0.0;
            for (int e = out_first[o]; e < out_first[o + 1]; ++e) s += __ldcg(partial + e);
            d_out[o] = s;
        }
        if (threadIdx.x == 0) *counter = 0u;
    }
}

// ----------------------------------------------------------------------------------------------
// host: geometry, caches, plan
// ----------------------------------------------------------------------------------------------
struct SecGeomHost {
    int n = 0, half = 0, n_up = 0, n_dn = 0;
    // index bits of the up / down orbitals, ascending; orbital k of a species is bit k (up) / 16 + k (down) of a compact mask.
    // Standard layout (even wires = up): up = odd index bits, down = even index bits; a slab of a sharded state has its own.
    std::vector<int> upbits, dnbits;
    void standard(int nq) {
        n = nq;
        half = nq / 2;
        upbits.clear();
        dnbits.clear();
        for (int b = 0; b < half; ++b) {
            upbits.push_back(2 * b + 1);
            dnbits.push_back(2 * b);
        }
    }
    void from_masks(int nq, u64 upmask, u64 dnmask) {
        n = nq;
        upbits.clear();
        dnbits.clear();
        for (int b = 0; b < nq; ++b) {
            if (upmask >> b & 1ull) upbits.push_back(b);
            else if (dnmask >> b & 1ull) dnbits.push_back(b);
        }
        half = (int)std::max(upbits.size(), dnbits.size());
    }
    std::vector<unsigned short> cfgU, cfgD, rankU, rankD;
    unsigned short *d_cfgU = nullptr, *d_cfgD = nullptr, *d_rankU = nullptr, *d_rankD = nullptr;
    unsigned d_up() const { return (unsigned)cfgU.size(); }
    unsigned d_dn() const { return (unsigned)cfgD.size(); }
    void release() {
        cudaFree(d_cfgU); cudaFree(d_cfgD); cudaFree(d_rankU); cudaFree(d_rankD);
        d_cfgU = d_cfgD = d_rankU = d_rankD = nullptr;
    }
};

static void sec_patterns(int nbits, int count, std::vector<unsigned short> &cfg, std::vector<unsigned short> &rank) {
    rank.assign((size_t)1 << nbits, 0xffffu);
    cfg.clear();
    for (unsigned pat = 0; pat < (1u << nbits); ++pat)
        if (__builtin_popcount(pat) == count) {
            rank[pat] = (unsigned short)cfg.size();
            cfg.push_back((unsigned short)pat);
        }
}

// full-index mask -> compact packed mask (up orbital b = index bit 2b+1 -> bit b; down orbital b = index bit 2b -> bit 16+b)
static unsigned sec_compact(u64 m, const SecGeomHost &G) {
    unsigned out = 0;
    for (size_t k = 0; k < G.upbits.size(); ++k)
        if (m >> G.upbits[k] & 1ull) out |= 1u << k;
    for (size_t k = 0; k < G.dnbits.size(); ++k)
        if (m >> G.dnbits[k] & 1ull) out |= 1u << (16 + k);
    return out;
}
static int sec_compact_pos(int p, const SecGeomHost &G) {
    for (size_t k = 0; k < G.upbits.size(); ++k)
        if (G.upbits[k] == p) return (int)k;
    for (size_t k = 0; k < G.dnbits.size(); ++k)
        if (G.dnbits[k] == p) return 16 + (int)k;
    return 31;
}

template <typename T>
static int sec_upload(T **d, const std::vector<T> &v) {
    *d = nullptr;
    if (v.empty()) return FH_OK;
    FH_CUDA(cudaMalloc(d, sizeof(T) * v.size()));
    FH_CUDA(cudaMemcpy(*d, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
    return FH_OK;
}

// does (x, fixmask, fixval) map the sector to itself?  x fully pinned, as many 1 -> 0 as 0 -> 1 flips per species
static bool sec_pattern_conserves(u64 x, u64 fixmask, u64 fixval, u64 upmask, u64 dnmask) {
    if (x & ~fixmask) return false;
    const int up1 = __builtin_popcountll(fixval & x & upmask), up = __builtin_popcountll(x & upmask);
    const int dn1 = __builtin_popcountll(fixval & x & dnmask), dn = __builtin_popcountll(x & dnmask);
    return 2 * up1 == up && 2 * dn1 == dn;
}

// the two lists of a pattern-pinned pair (x, fixmask, fixval, zeta) in compact coordinates; entry = rank | partner << SEC_FB | sign << 31
static void sec_build_lists(const SecGeomHost &G, unsigned xc, unsigned fmc, unsigned fvc, unsigned ztc, std::vector<unsigned> &LU,
                            std::vector<unsigned> &LD) {
    LU.clear();
    LD.clear();
    const unsigned xu = xc & 0xffffu, xd = xc >> 16, fmu = fmc & 0xffffu, fmd = fmc >> 16, fvu = fvc & 0xffffu, fvd = fvc >> 16,
                   zu = ztc & 0xffffu, zd = ztc >> 16;
    for (unsigned r = 0; r < G.d_up(); ++r) {
        const unsigned u = G.cfgU[r];
        if ((u & fmu) != fvu) continue;
        const unsigned pr = G.rankU[u ^ xu];
        LU.push_back(r | (pr << SEC_FB) | ((unsigned)(__builtin_popcount(u & zu) & 1) << 31));
    }
    for (unsigned r = 0; r < G.d_dn(); ++r) {
        const unsigned d = G.cfgD[r];
        if ((d & fmd) != fvd) continue;
        const unsigned pr = G.rankD[d ^ xd];
        LD.push_back(r | (pr << SEC_FB) | ((unsigned)(__builtin_popcount(d & zd) & 1) << 31));
    }
}

// ---- per-table cache: groups in compact coordinates + the diagonal on the sector ----
struct SecTableCache {
    u64 table_uid = 0;
    int n_up = -1, n_dn = -1;
    int ngroups = 0, nclasses = 0, nvals = 0;
    SecGroup *d_groups = nullptr;
    SecClass *d_classes = nullptr;
    double2 *d_vals = nullptr, *d_hdiag = nullptr;
    void release() {
        cudaFree(d_groups); cudaFree(d_classes); cudaFree(d_vals); cudaFree(d_hdiag);
        d_groups = nullptr; d_classes = nullptr; d_vals = d_hdiag = nullptr;
    }
};

// every x != 0 group tabulated (<= 4 x bits) and every live pattern keeps (N_up, N_dn)?
static bool sec_table_conserves(const fh_table *tab, u64 upmask, u64 dnmask) {
    for (const TabGroup &g : tab->groups) {
        if (g.x == 0) continue;
        if (g.kbits == 0) return false;
        for (int pat = 0; pat < (1 << g.kbits); ++pat) {
            if (!((g.live >> pat) & 1u)) continue;
            int dup = 0, ddn = 0;
            for (int b = 0; b < g.kbits; ++b) {
                const u64 bit = 1ull << g.pos[b];
                const int was = (pat >> b) & 1;
                if (bit & upmask) dup += 1 - 2 * was;
                if (bit & dnmask) ddn += 1 - 2 * was;
            }
            if (dup != 0 || ddn != 0) return false;
        }
    }
    return true;
}

static int sec_build_table(const fh_table *tab, const SecGeomHost &G, u64 upmask, u64 dnmask, SecTableCache &T, bool *ok) {
    *ok = false;
    std::vector<SecGroup> groups;
    std::vector<SecClass> classes;
    std::vector<double2> hdiag((size_t)G.d_up() * G.d_dn(), make_double2(0.0, 0.0));
    for (const TabGroup &g : tab->groups) {
        if (g.x == 0) continue;                   // the diagonal comes from tab->diag_terms below
        if (g.kbits == 0) return FH_OK;           // more than 4 x bits: not tabulated
        // every live pattern must keep (N_up, N_dn)
        for (int pat = 0; pat < (1 << g.kbits); ++pat) {
            if (!((g.live >> pat) & 1u)) continue;
            int dup = 0, ddn = 0;
            for (int b = 0; b < g.kbits; ++b) {
                const u64 bit = 1ull << g.pos[b];
                const int was = (pat >> b) & 1;
                if (bit & upmask) dup += 1 - 2 * was;
                if (bit & dnmask) ddn += 1 - 2 * was;
            }
            if (dup != 0 || ddn != 0) return FH_OK;
        }
        SecGroup s;
        memset(&s, 0, sizeof(s));
        s.x = sec_compact(g.x, G);
        s.first_class = (int)classes.size();
        s.n_class = g.n_class;
        s.live = g.live;
        s.kbits = g.kbits;
        for (int b = 0; b < 4; ++b) s.pos[b] = b < g.kbits ? (unsigned char)sec_compact_pos(g.pos[b], G) : 31;
        s.zeta1 = sec_compact(tab->classes[g.first_class].zeta, G);
        s.vofs1 = tab->classes[g.first_class].vofs;
        // bits 8..9 of kbits: 1 = x touches up orbitals only, 2 = down orbitals only (the partner keeps the other rank)
        if ((s.x >> 16) == 0u) s.kbits |= 1 << 8;
        else if ((s.x & 0xffffu) == 0u) s.kbits |= 2 << 8;
        for (int c = g.first_class; c < g.first_class + g.n_class; ++c) {
            SecClass sc;
            sc.zeta = sec_compact(tab->classes[c].zeta, G);
            sc.vofs = tab->classes[c].vofs;
            classes.push_back(sc);
        }
        groups.push_back(s);
    }
    for (const TabTerm &t : tab->diag_terms) {
        const unsigned zc = sec_compact(t.z, G);
        for (unsigned ru = 0; ru < G.d_up(); ++ru)
            for (unsigned rd = 0; rd < G.d_dn(); ++rd) {
                const unsigned cfg = (unsigned)G.cfgU[ru] | ((unsigned)G.cfgD[rd] << 16);
                double2 &h = hdiag[(size_t)ru * G.d_dn() + rd];
                if (__builtin_popcount(cfg & zc) & 1) {
                    h.x -= t.dr;
                    h.y -= t.di;
                } else {
                    h.x += t.dr;
                    h.y += t.di;
                }
            }
    }
    T.release();
    T.ngroups = (int)groups.size();
    T.nclasses = (int)classes.size();
    T.nvals = (int)tab->vals.size();
    FH_TRY(sec_upload(&T.d_groups, groups));
    FH_TRY(sec_upload(&T.d_classes, classes));
    FH_TRY(sec_upload(&T.d_vals, tab->vals));
    FH_TRY(sec_upload(&T.d_hdiag, hdiag));
    T.table_uid = tab->uid;
    T.n_up = G.n_up;
    T.n_dn = G.n_dn;
    *ok = true;
    return FH_OK;
}

// ---- per-pool cache ----
struct SecPoolCache {
    u64 pool_uid = 0, upmask = 0, dnmask = 0;
    int n_up = -1, n_dn = -1;
    SecPoolEntry *d_entries = nullptr;
    unsigned *d_lists = nullptr, *d_counter = nullptr;
    double *d_partial = nullptr;
    unsigned max_words = 0;
    void release() {
        cudaFree(d_entries); cudaFree(d_lists); cudaFree(d_counter); cudaFree(d_partial);
        d_entries = nullptr; d_lists = nullptr; d_counter = nullptr; d_partial = nullptr;
    }
};

static int sec_build_pool(const fh_pool *pool, const SecGeomHost &G, u64 upmask, u64 dnmask, SecPoolCache &Pc, bool *ok) {
    *ok = false;
    std::vector<SecPoolEntry> ents;
    std::vector<unsigned> lists, LU, LD;
    unsigned max_words = 0;
    for (const PoolEntry &e : pool->entries) {
        if (!sec_pattern_conserves(e.x, e.fixmask, e.fixval, upmask, dnmask)) return FH_OK;
        sec_build_lists(G, sec_compact(e.x, G), sec_compact(e.fixmask, G), sec_compact(e.fixval, G), sec_compact(e.zeta, G), LU, LD);
        SecPoolEntry s;
        s.listU = (unsigned)lists.size();
        s.nU = (unsigned)LU.size();
        lists.insert(lists.end(), LU.begin(), LU.end());
        s.listD = (unsigned)lists.size();
        s.nD = (unsigned)LD.size();
        lists.insert(lists.end(), LD.begin(), LD.end());
        s.br = e.br;
        s.bi = e.bi;
        ents.push_back(s);
        max_words = std::max(max_words, s.nU + s.nD);
    }
    Pc.release();
    FH_TRY(sec_upload(&Pc.d_entries, ents));
    if (lists.empty()) lists.push_back(0u);
    FH_TRY(sec_upload(&Pc.d_lists, lists));
    FH_CUDA(cudaMalloc(&Pc.d_counter, sizeof(unsigned)));
    FH_CUDA(cudaMemset(Pc.d_counter, 0, sizeof(unsigned)));
    FH_CUDA(cudaMalloc(&Pc.d_partial, sizeof(double) * std::max<size_t>(1, ents.size())));
    Pc.max_words = max_words;
    Pc.pool_uid = pool->uid;
    Pc.n_up = G.n_up;
    Pc.n_dn = G.n_dn;
    *ok = true;
    return FH_OK;
}

// caches live as long as their table / pool handle; keyed by handle uid.  Handles are not thread-safe, but these maps are
// shared by all of them: every function that touches one holds g_sec_mutex (recursive: prepare calls build / forget)
static std::recursive_mutex g_sec_mutex;
#define SEC_LOCK std::lock_guard<std::recursive_mutex> sec_lock_guard_(g_sec_mutex)
static std::map<u64, SecTableCache> g_sec_tables;
static std::map<u64, SecPoolCache> g_sec_pools;
void fh_sector_forget_table(u64 uid) {
    SEC_LOCK;
    auto it = g_sec_tables.find(uid);
    if (it != g_sec_tables.end()) {
        it->second.release();
        g_sec_tables.erase(it);
    }
}
void fh_sector_forget_pool(u64 uid) {
    SEC_LOCK;
    auto it = g_sec_pools.find(uid);
    if (it != g_sec_pools.end()) {
        it->second.release();
        g_sec_pools.erase(it);
    }
}

// ---- the plan of one program ----
struct fh_sector_plan {
    bool eligible = false;
    // key
    int n_up = -1, n_dn = -1, pool_flat = -1, has_pool = 0, C = 0;
    u64 table_uid = 0, pool_uid = 0, max_dim = 0;
    int prefix_flat = -1;
    SecGeomHost G;
    SecVOp *d_vops = nullptr;
    int *d_term_first = nullptr;
    unsigned *d_term_zc = nullptr, *d_lists = nullptr;
    double2 *d_chk = nullptr, *d_lam = nullptr;
    int nvops = 0, nterms_total = 0, layout0 = 0, npair_vops = 0;
    long long *d_timeline = nullptr;
    unsigned la_max = 1, lb_max = 1, S = 0;
    size_t smem = 0;
    int n_transposes = 0, n_remote = 0;
    void release() {
        G.release();
        cudaFree(d_timeline);
        d_timeline = nullptr;
        cudaFree(d_vops); cudaFree(d_term_first); cudaFree(d_term_zc); cudaFree(d_lists); cudaFree(d_chk); cudaFree(d_lam);
        d_vops = nullptr; d_term_first = nullptr; d_term_zc = nullptr; d_lists = nullptr; d_chk = d_lam = nullptr;
    }
};

void fh_sector_plan_free(fh_sector_plan *plan) {
    if (!plan) return;
    plan->release();
    delete plan;
}

static int g_sec_cluster_dev[64];     // usable cluster size per device + 1 (0: not probed yet)
static int sec_cluster_size(size_t smem) {
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (g_sec_cluster_dev[dev]) return g_sec_cluster_dev[dev] - 1;
    int g_sec_cluster = 0;
    struct Remember {
        int *slot, *val;
        ~Remember() { *slot = *val + 1; }
    } remember{&g_sec_cluster_dev[dev], &g_sec_cluster};
    if (cudaFuncSetAttribute(k_sector_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return 0;
    cudaFuncSetAttribute(k_sector_eval, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    int want = 16;
    if (const char *env = getenv("FHSIM_SECTOR_CLUSTER")) want = atoi(env);
    for (int c = want; c >= 2; c >>= 1) {
        if (c & (c - 1)) continue;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)c);
        cfg.blockDim = dim3(SEC_THREADS);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)c;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int nclusters = 0;
        if (cudaOccupancyMaxActiveClusters(&nclusters, k_sector_eval, &cfg) == cudaSuccess && nclusters >= 1) {
            g_sec_cluster = c;
            break;
        }
    }
    cudaGetLastError();
    return g_sec_cluster;
}

struct SecLogical {
    int kind, index, nterms, dagger;     // kind: SV_*
    unsigned xc;                         // pair: compact x
};

// Build (or reuse) the plan.  Returns FH_OK with plan->eligible telling whether the sector path applies.
int fh_sector_prepare(fh_sector_plan **slot, fh_ctx *ctx, int n, u64 basis, const std::vector<PairOp> &pairs,
                      const std::vector<DiagOp> &diagops, const std::vector<DiagTerm> &dterms, const std::vector<SecFlatOp> &flat,
                      int pool_flat, const fh_table *tab, const fh_pool *pool, u64 max_dim, int prefix_flat) {
    SEC_LOCK;
    // prefix_flat >= 0: only the flat ops [0, prefix_flat) and a checkpoint of the compressed state (no observable, no pool):
    // the ansatz part of a screening whose tail runs as dense sector blocks
    if (!*slot) *slot = new fh_sector_plan();
    fh_sector_plan *P = *slot;
    const int half = n / 2;
    u64 upmask = 0, dnmask = 0;
    for (int b = 0; b < half; ++b) {
        upmask |= 1ull << (2 * b + 1);
        dnmask |= 1ull << (2 * b);
    }
    const int n_up = __builtin_popcountll(basis & upmask), n_dn = __builtin_popcountll(basis & dnmask);
    const bool same = P->n_up == n_up && P->n_dn == n_dn && P->pool_flat == pool_flat && P->has_pool == (pool ? 1 : 0) &&
                      P->table_uid == tab->uid && P->pool_uid == (pool ? pool->uid : 0) && P->max_dim == max_dim &&
                      P->prefix_flat == prefix_flat;
    if (same) return FH_OK;
    P->release();
    P->eligible = false;
    P->n_up = n_up;
    P->n_dn = n_dn;
    P->pool_flat = pool_flat;
    P->has_pool = pool ? 1 : 0;
    P->table_uid = tab->uid;
    P->pool_uid = pool ? pool->uid : 0;
    P->max_dim = max_dim;
    P->prefix_flat = prefix_flat;
    if ((n & 1) || half < 1 || half > 15) return FH_OK;

    SecGeomHost &G = P->G;
    G.standard(n);
    G.n_up = n_up;
    G.n_dn = n_dn;
    sec_patterns(half, n_up, G.cfgU, G.rankU);
    sec_patterns(half, n_dn, G.cfgD, G.rankD);
    const unsigned d_up = G.d_up(), d_dn = G.d_dn();
    if (d_up > SEC_MAX_D || d_dn > SEC_MAX_D) return FH_OK;
    const u64 dim = (u64)d_up * d_dn;
    if (max_dim && dim > max_dim) return FH_OK;
    if (dim * 32 > (u64)SEC_MAX_C * 180 * 1024) return FH_OK;      // psi + lambda cannot fit a 16-CTA cluster

    // ops must map the sector to itself
    for (const SecFlatOp &f : flat)
        if (f.type == 1) {
            const PairOp &op = pairs[f.index];
            if (!sec_pattern_conserves(op.x, op.fixmask, op.fixval, upmask, dnmask)) return FH_OK;
        }

    // logical sequence
    std::vector<SecLogical> seq;
    auto push_op = [&](const SecFlatOp &f, int dagger) {
        SecLogical L;
        L.dagger = dagger;
        if (f.type == 1) {
            L.kind = SV_PAIR;
            L.index = f.index;
            L.nterms = 0;
            L.xc = sec_compact(pairs[f.index].x, P->G);
        } else {
            L.kind = SV_DIAG;
            L.index = diagops[f.index].first;
            L.nterms = diagops[f.index].count;
            L.xc = 0;
        }
        seq.push_back(L);
    };
    const int nflat = (int)flat.size();
    const bool prefix = prefix_flat >= 0;
    if (prefix) {
        for (int k = 0; k < prefix_flat && k < nflat; ++k) push_op(flat[k], 0);
        seq.push_back({SV_CHECKPOINT, 0, 0, 0, 0});
    } else {
        for (int k = 0; k < nflat; ++k) {
            if (pool && k == pool_flat) seq.push_back({SV_CHECKPOINT, 0, 0, 0, 0});
            push_op(flat[k], 0);
        }
        if (pool && pool_flat == nflat) seq.push_back({SV_CHECKPOINT, 0, 0, 0, 0});
        seq.push_back({SV_HAPPLY, 0, 0, 0, 0});
        if (pool) {
            for (int k = nflat - 1; k >= pool_flat; --k) push_op(flat[k], 1);
            seq.push_back({SV_STORE, 0, 0, 0, 0});
        }
    }

    // usable cluster size decides the distribution
    // (smem is known only after the plan; probe with the maximum the kernel may ask for)
    const int C = sec_cluster_size(200 * 1024);
    if (C < 2) return FH_OK;
    P->C = C;
    int logC = 0;
    while ((1 << logC) < C) ++logC;

    // layouts and transposes: an op whose x touches only down orbitals is CTA-local in ROW layout, only up orbitals in COL
    auto pure = [&](const SecLogical &L) -> int {        // 0 none / mixed / not a pair, 1 wants ROW, 2 wants COL
        if (L.kind != SV_PAIR) return 0;
        const unsigned xu = L.xc & 0xffffu, xd = L.xc >> 16;
        if (xu == 0u && xd != 0u) return 1;
        if (xd == 0u && xu != 0u) return 2;
        return 0;
    };
    int layout = 0;
    for (const SecLogical &L : seq) {
        const int w = pure(L);
        if (w) {
            layout = w - 1;
            break;
        }
    }
    P->layout0 = layout;
    static const bool no_transpose = getenv("FHSIM_SECTOR_NO_TRANSPOSE") != nullptr;
    std::vector<SecLogical> seq2;
    std::vector<int> lay_of;
    for (size_t k = 0; k < seq.size(); ++k) {
        const int w = pure(seq[k]);
        if (w && w - 1 != layout && !no_transpose) {
            // run of ops wanting the other layout before one wants this layout again
            int run = 0;
            for (size_t q = k; q < seq.size(); ++q) {
                const int wq = pure(seq[q]);
                if (wq == w) ++run;
                else if (wq != 0) break;
            }
            if (run >= 3) {
                seq2.push_back({SV_TRANSPOSE, 0, 0, 0, 0});
                lay_of.push_back(layout);
                layout ^= 1;
                P->n_transposes++;
            }
        }
        seq2.push_back(seq[k]);
        lay_of.push_back(layout);
    }

    // vops + lists
    const int nv = (int)seq2.size();
    std::vector<SecVOp> vops((size_t)nv);
    std::vector<int> term_first((size_t)nv, 0);
    std::vector<unsigned> lists;
    std::vector<char> remote((size_t)nv, 0);
    std::map<int, std::pair<unsigned, unsigned>> list_of_pair;       // pair index -> (listU, listD) (forward and dagger share)
    std::map<int, std::vector<unsigned short>> offs_of_pair;
    unsigned la_max = 1, lb_max = 1;
    int nterms_total = 0, npair_vops = 0;
    std::vector<unsigned> LU, LD;
    for (int v = 0; v < nv; ++v) {
        const SecLogical &L = seq2[v];
        SecVOp &o = vops[v];
        memset(&o, 0, sizeof(o));
        o.kind = L.kind;
        o.index = L.index;
        o.nterms = L.nterms;
        o.flags = (L.dagger ? SF_DAGGER : 0) | (lay_of[v] ? SF_COL : 0);
        if (L.kind == SV_PAIR) {
            const PairOp &op = pairs[L.index];
            auto it = list_of_pair.find(L.index);
            if (it == list_of_pair.end()) {
                sec_build_lists(G, sec_compact(op.x, G), sec_compact(op.fixmask, G), sec_compact(op.fixval, G), sec_compact(op.zeta, G),
                                LU, LD);
                // group by owner (rank % C), stable
                std::vector<unsigned short> offs(2 * (SEC_MAX_C + 1), 0);
                auto grouped = [&](const std::vector<unsigned> &Lx, unsigned short *off) {
                    std::vector<unsigned> out;
                    for (int c = 0; c < C; ++c) {
                        off[c] = (unsigned short)out.size();
                        for (unsigned e : Lx)
                            if ((int)((e & SEC_FM) & (unsigned)(C - 1)) == c) out.push_back(e);
                    }
                    for (int c = C; c <= SEC_MAX_C; ++c) off[c] = (unsigned short)out.size();
                    return out;
                };
                const std::vector<unsigned> gu = grouped(LU, offs.data()), gd = grouped(LD, offs.data() + SEC_MAX_C + 1);
                const unsigned bu = (unsigned)lists.size();
                lists.insert(lists.end(), gu.begin(), gu.end());
                const unsigned bd = (unsigned)lists.size();
                lists.insert(lists.end(), gd.begin(), gd.end());
                it = list_of_pair.emplace(L.index, std::make_pair(bu, bd)).first;
                offs_of_pair[L.index] = offs;
            }
            o.listU = it->second.first;
            o.listD = it->second.second;
            o.slot = npair_vops++;
            const std::vector<unsigned short> &offs = offs_of_pair[L.index];
            memcpy(o.offU, offs.data(), sizeof(o.offU));
            memcpy(o.offD, offs.data() + SEC_MAX_C + 1, sizeof(o.offD));
            const unsigned short *offA = lay_of[v] ? o.offD : o.offU, *offB = lay_of[v] ? o.offU : o.offD;
            for (int c = 0; c < C; ++c) la_max = std::max<unsigned>(la_max, offA[c + 1] - offA[c]);
            lb_max = std::max<unsigned>(lb_max, offB[C]);
            const int w = pure(L);
            remote[v] = !(w && w - 1 == lay_of[v]);
            if (remote[v]) P->n_remote++;
            else o.flags |= SF_LOCAL;
        } else if (L.kind == SV_DIAG) {
            term_first[v] = nterms_total;
            nterms_total += L.nterms;
        } else if (L.kind == SV_TRANSPOSE || L.kind == SV_HAPPLY) {
            remote[v] = 1;
            if (L.kind == SV_HAPPLY && pool) o.flags |= SF_WANT_LAMBDA;
        }
    }
    for (int v = 0; v < nv; ++v)
        if (remote[v] || (v + 1 < nv && remote[v + 1])) vops[v].flags |= SF_CLUSTER_AFTER;

    // shared memory
    const unsigned dmax = std::max(d_up, d_dn), rmax = (dmax + C - 1) >> logC;
    const unsigned S = std::max(((d_up + C - 1) >> logC) * d_dn, ((d_dn + C - 1) >> logC) * d_up);
    SecTableCache Tnone;                     // prefix plans read no observable
    SecTableCache &T = prefix ? Tnone : g_sec_tables[tab->uid];
    if (!prefix && (T.table_uid != tab->uid || T.n_up != n_up || T.n_dn != n_dn)) {
        bool ok = false;
        FH_TRY(sec_build_table(tab, G, upmask, dnmask, T, &ok));
        if (!ok) {
            fh_sector_forget_table(tab->uid);
            return FH_OK;
        }
    }
    if (pool && !prefix) {
        SecPoolCache &Pc = g_sec_pools[pool->uid];
        if (Pc.pool_uid != pool->uid || Pc.n_up != n_up || Pc.n_dn != n_dn) {
            bool ok = false;
            FH_TRY(sec_build_pool(pool, G, upmask, dnmask, Pc, &ok));
            if (!ok) {
                fh_sector_forget_pool(pool->uid);
                return FH_OK;
            }
        }
    }
    size_t smem = (size_t)S * 32 + (size_t)nv * sizeof(SecVOp) + (size_t)nv * 64 + (size_t)((nterms_total + 1) & ~1) * sizeof(SecTerm) +
                  (size_t)((nv + 3) & ~3) * 4 + (size_t)T.ngroups * sizeof(SecGroup) + (size_t)((T.nclasses + 1) & ~1) * sizeof(SecClass) +
                  (size_t)T.nvals * 16 + (size_t)(rmax + dmax) * 16 + (size_t)(nv + 1) * 16 + (size_t)npair_vops * (la_max + lb_max) * 4 +
                  (size_t)(((d_up + 7) & ~7u) + ((d_dn + 7) & ~7u) + 2 * (1u << half)) * 2 + 64;
    if (smem > 200 * 1024) return FH_OK;

    std::vector<unsigned> zc(std::max<size_t>(1, dterms.size()), 0u);
    for (size_t m = 0; m < dterms.size(); ++m) zc[m] = sec_compact(dterms[m].z, G);
    if (lists.empty()) lists.push_back(0u);
    FH_TRY(sec_upload(&G.d_cfgU, G.cfgU));
    FH_TRY(sec_upload(&G.d_cfgD, G.cfgD));
    FH_TRY(sec_upload(&G.d_rankU, G.rankU));
    FH_TRY(sec_upload(&G.d_rankD, G.rankD));
    FH_TRY(sec_upload(&P->d_vops, vops));
    FH_TRY(sec_upload(&P->d_term_first, term_first));
    FH_TRY(sec_upload(&P->d_term_zc, zc));
    FH_TRY(sec_upload(&P->d_lists, lists));
    FH_CUDA(cudaMalloc(&P->d_chk, sizeof(double2) * dim));
    FH_CUDA(cudaMalloc(&P->d_lam, sizeof(double2) * dim));
    P->nvops = nv;
    P->nterms_total = nterms_total;
    P->npair_vops = npair_vops;
    P->la_max = la_max;
    P->lb_max = lb_max;
    P->S = S;
    P->smem = smem;
    P->eligible = true;
    (void)ctx;
    return FH_OK;
}

bool fh_sector_plan_eligible(const fh_sector_plan *plan) { return plan && plan->eligible; }

void fh_sector_plan_describe(const fh_sector_plan *plan, int *cluster, u64 *dim, int *nvops, int *transposes, int *remote_ops,
                             size_t *smem) {
    if (cluster) *cluster = plan ? plan->C : 0;
    if (dim) *dim = plan ? (u64)plan->G.d_up() * plan->G.d_dn() : 0;
    if (nvops) *nvops = plan ? plan->nvops : 0;
    if (transposes) *transposes = plan ? plan->n_transposes : 0;
    if (remote_ops) *remote_ops = plan ? plan->n_remote : 0;
    if (smem) *smem = plan ? plan->smem : 0;
}

// Enqueue the evaluation: E -> d_res[0..1]; pool outputs -> d_pool_out[first .. first+count)
int fh_sector_enqueue(fh_sector_plan *P, fh_ctx *ctx, u64 basis, const PairOp *d_pairs, const DiagTerm *d_dterms, double *d_res,
                      const fh_pool *pool, int pool_first, int pool_count, double *d_pool_out, double2 *chk_override) {
    SEC_LOCK;
    const SecGeomHost &G = P->G;
    const SecTableCache Tnone;
    const SecTableCache &T = P->prefix_flat >= 0 ? Tnone : g_sec_tables[P->table_uid];
    SecArgs A;
    memset(&A, 0, sizeof(A));
    A.half = G.half;
    A.C = P->C;
    A.logC = 0;
    while ((1 << A.logC) < P->C) ++A.logC;
    A.d_up = G.d_up();
    A.d_dn = G.d_dn();
    A.S = P->S;
    A.cfgU = G.d_cfgU;
    A.cfgD = G.d_cfgD;
    A.rankU = G.d_rankU;
    A.rankD = G.d_rankD;
    A.vops = P->d_vops;
    A.nvops = P->nvops;
    A.nterms_total = P->nterms_total;
    A.term_first = P->d_term_first;
    A.term_zc = P->d_term_zc;
    A.lists = P->d_lists;
    A.pairs = d_pairs;
    A.dterms = d_dterms;
    A.groups = T.d_groups;
    A.classes = T.d_classes;
    A.vals = T.d_vals;
    A.hdiag = T.d_hdiag;
    A.ngroups = T.ngroups;
    A.nclasses = T.nclasses;
    A.nvals = T.nvals;
    A.chk = chk_override ? chk_override : P->d_chk;
    A.lam_out = P->d_lam;
    A.res = d_res;
    u64 upc = 0, dnc = 0;
    for (int b = 0; b < G.half; ++b) {
        if (basis >> (2 * b + 1) & 1ull) upc |= 1ull << b;
        if (basis >> (2 * b) & 1ull) dnc |= 1ull << b;
    }
    A.basis_ru = G.rankU[upc];
    A.basis_rd = G.rankD[dnc];
    A.layout0 = P->layout0;
    A.la_max = P->la_max;
    A.lb_max = P->lb_max;
    A.npair_vops = P->npair_vops;
    static const bool want_timeline = getenv("FHSIM_SECTOR_TIMELINE") != nullptr;
    if (want_timeline && !P->d_timeline) cudaMalloc(&P->d_timeline, sizeof(long long) * (size_t)(4 * P->nvops + 4));
    A.timeline = want_timeline ? P->d_timeline : nullptr;

    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)P->C);
    cfg.blockDim = dim3(SEC_THREADS);
    cfg.dynamicSmemBytes = P->smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)P->C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ++g_fh_launch_count;
    FH_CUDA(cudaLaunchKernelEx(&cfg, k_sector_eval, A));
    if (pool && pool_count > 0 && P->prefix_flat < 0) {
        const SecPoolCache &Pc = g_sec_pools[pool->uid];
        const int e0 = pool->out_first[pool_first], e1 = pool->out_first[pool_first + pool_count];
        int grid = e1 - e0;
        if (grid > ctx->sm_count * 4) grid = ctx->sm_count * 4;
        if (grid < 1) grid = 1;
        ++g_fh_launch_count;
        k_sector_pool<<<grid, 256, sizeof(unsigned) * std::max(1u, Pc.max_words), ctx->stream>>>(
            Pc.d_entries, Pc.d_lists, e0, e1, G.d_dn(), P->d_chk, P->d_lam, Pc.d_partial, pool->d_out_first, pool_first, pool_count,
            d_pool_out, Pc.d_counter);
    }
    FH_CUDA(cudaGetLastError());
    if (A.timeline) {
        // debugging only (use with FHSIM_NO_GRAPH=1): cycles of CTA 0 per vop of the launch above
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(ctx->stream, &cs);
        if (cs == cudaStreamCaptureStatusNone) {
            std::vector<long long> t((size_t)4 * P->nvops + 4);
            cudaStreamSynchronize(ctx->stream);
            cudaMemcpy(t.data(), P->d_timeline, sizeof(long long) * t.size(), cudaMemcpyDeviceToHost);
            std::vector<SecVOp> hv((size_t)P->nvops);
            cudaMemcpy(hv.data(), P->d_vops, sizeof(SecVOp) * hv.size(), cudaMemcpyDeviceToHost);
            fprintf(stderr, "sector timeline (%d vops, total %lld cycles; per vop: total [setup/work/barrier]):", P->nvops,
                    t[4 * (size_t)P->nvops] - t[0]);
            for (int v = 0; v < P->nvops; ++v) {
                const long long *q = &t[4 * (size_t)v];
                if (hv[v].kind == SV_PAIR)
                    fprintf(stderr, " %c%s%lld[%lld/%lld/%lld]", "?PDTCHS"[hv[v].kind], (hv[v].flags & SF_CLUSTER_AFTER) ? "*" : "",
                            q[4] - q[0], q[1] - q[0], q[2] - q[1], q[3] - q[2]);
                else
                    fprintf(stderr, " %c%s%lld[%lld/%lld]", "?PDTCHS"[hv[v].kind], (hv[v].flags & SF_CLUSTER_AFTER) ? "*" : "",
                            q[4] - q[0], q[2] - q[0], q[3] - q[2]);
            }
            fprintf(stderr, "\n");
        }
    }
    return FH_OK;
}


// ----------------------------------------------------------------------------------------------
// K3 on compressed copies of FULL-SPACE states (the 19-launch path keeps its tile kernels for the circuit, but screens the
// pool in the sector): psi_s and lambda_s are gathered into rank order (k_sector_compress2), then k_sector_pool.
// 3x3: 324 x 1 225 pairs instead of 324 x 2^15; 3x4: the two compressed vectors are 13.7 MB each and stay in L2, while the
// full-space kernel streams 4 * 2^24 B per gradient from HBM.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sector_compress2(const unsigned *__restrict__ depU, const unsigned *__restrict__ depD,
                                                          unsigned d_dn, unsigned dim, const double2 *__restrict__ a,
                                                          const double2 *__restrict__ b, double2 *__restrict__ ac,
                                                          double2 *__restrict__ bc) {
    SEC_PDL_PROLOGUE();
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned r = blockIdx.x * blockDim.x + threadIdx.x; r < dim; r += stride) {
        const unsigned ru = r / d_dn, rd = r - ru * d_dn;
        const unsigned i = __ldg(depU + ru) | __ldg(depD + rd);          // full index of sector state r
        ac[r] = a[i];
        bc[r] = b[i];
    }
}

struct fh_sector_pool_plan {
    bool eligible = false;           // K3 in the sector
    bool table_ok = false;           // K2 in the sector as well (standard species layout, observable in compact form)
    int n_up = -1, n_dn = -1, n = 0, half = 0;
    u64 table_uid = 0, pool_uid = 0, upmask = 0, dnmask = 0;
    unsigned d_up = 0, d_dn = 0;
    unsigned *d_depU = nullptr, *d_depD = nullptr;      // index bits of every up / down pattern, rank order
    double2 *d_psi = nullptr, *d_lam = nullptr;         // K3: compressed psi_s, lambda_s
    double2 *d_in = nullptr, *d_out = nullptr;          // K2: compressed psi, H psi
    unsigned short *d_cfgU = nullptr, *d_cfgD = nullptr, *d_rankU = nullptr, *d_rankD = nullptr;
    double *d_k2_partials = nullptr;
    unsigned *d_k2_counter = nullptr;
    void release() {
        cudaFree(d_depU); cudaFree(d_depD); cudaFree(d_psi); cudaFree(d_lam); cudaFree(d_in); cudaFree(d_out);
        cudaFree(d_cfgU); cudaFree(d_cfgD); cudaFree(d_rankU); cudaFree(d_rankD); cudaFree(d_k2_partials); cudaFree(d_k2_counter);
        d_depU = d_depD = nullptr; d_psi = d_lam = d_in = d_out = nullptr;
        d_cfgU = d_cfgD = d_rankU = d_rankD = nullptr; d_k2_partials = nullptr; d_k2_counter = nullptr;
    }
};

void fh_sector_pool_plan_free(fh_sector_pool_plan *plan) {
    if (!plan) return;
    plan->release();
    delete plan;
}
bool fh_sector_pool_plan_eligible(const fh_sector_pool_plan *plan) { return plan && plan->eligible; }
bool fh_sector_pool_plan_table_ok(const fh_sector_pool_plan *plan) { return plan && plan->table_ok; }

// upmask / dnmask: the index bits of the up / down orbitals (together all n bits); 0, 0 = the standard layout
int fh_sector_pool_prepare(fh_sector_pool_plan **slot, fh_ctx *ctx, int n, u64 upmask, u64 dnmask, int n_up, int n_dn,
                           const std::vector<PairOp> &pairs, const std::vector<SecFlatOp> &flat, const fh_table *tab,
                           const fh_pool *pool) {
    SEC_LOCK;
    if (!*slot) *slot = new fh_sector_pool_plan();
    fh_sector_pool_plan *P = *slot;
    if (upmask == 0 && dnmask == 0)
        for (int b = 0; b < n; ++b) ((b & 1) ? upmask : dnmask) |= 1ull << b;
    const u64 pool_uid = pool ? pool->uid : 0;
    // a plan built with a pool also serves the calls without one (drivers alternate screening and training evaluations)
    if (P->n_up == n_up && P->n_dn == n_dn && P->table_uid == tab->uid && (P->pool_uid == pool_uid || !pool) && P->upmask == upmask &&
        P->dnmask == dnmask && P->n == n)
        return FH_OK;
    P->release();
    P->eligible = false;
    P->table_ok = false;
    P->n = n;
    P->n_up = n_up;
    P->n_dn = n_dn;
    P->table_uid = tab->uid;
    P->pool_uid = pool_uid;
    P->upmask = upmask;
    P->dnmask = dnmask;
    const u64 full = n >= 64 ? ~0ull : ((1ull << n) - 1ull);
    if (n < 2 || n > 31 || (upmask & dnmask) || (upmask | dnmask) != full) return FH_OK;
    SecGeomHost G;
    G.from_masks(n, upmask, dnmask);
    const int nbu = (int)G.upbits.size(), nbd = (int)G.dnbits.size();
    if (nbu < 1 || nbd < 1 || nbu > 16 || nbd > 16 || n_up < 0 || n_up > nbu || n_dn < 0 || n_dn > nbd) return FH_OK;
    // psi_s and lambda_s stay in the sector when every op, the observable and the pool map it to itself
    for (const SecFlatOp &f : flat)
        if (f.type == 1) {
            const PairOp &op = pairs[f.index];
            if (!sec_pattern_conserves(op.x, op.fixmask, op.fixval, upmask, dnmask)) return FH_OK;
        }
    if (!sec_table_conserves(tab, upmask, dnmask)) return FH_OK;
    G.n_up = n_up;
    G.n_dn = n_dn;
    sec_patterns(nbu, n_up, G.cfgU, G.rankU);
    sec_patterns(nbd, n_dn, G.cfgD, G.rankD);
    if (G.d_up() > SEC_MAX_D || G.d_dn() > SEC_MAX_D || (u64)G.d_up() * G.d_dn() >= (1ull << 32)) return FH_OK;
    bool pool_ok = false;
    if (pool) {
        SecPoolCache &Pc = g_sec_pools[pool->uid];
        pool_ok = true;
        if (Pc.pool_uid != pool->uid || Pc.n_up != n_up || Pc.n_dn != n_dn || Pc.upmask != upmask || Pc.dnmask != dnmask) {
            FH_TRY(sec_build_pool(pool, G, upmask, dnmask, Pc, &pool_ok));
            if (!pool_ok) fh_sector_forget_pool(pool->uid);
            Pc.upmask = upmask;
            Pc.dnmask = dnmask;
        }
    }
    // K2 in the sector: the observable in compact coordinates (standard species layout, real table handle only)
    bool standard = !(n & 1) && n / 2 <= 15;
    for (int b = 0; b < n && standard; ++b) standard = ((upmask >> b) & 1ull) == (u64)(b & 1);
    bool table_ok = false;
    if (standard && tab->uid != ~0ull && !getenv("FHSIM_NO_SECTOR_K2")) {
        SecTableCache &T = g_sec_tables[tab->uid];
        table_ok = true;
        if (T.table_uid != tab->uid || T.n_up != n_up || T.n_dn != n_dn) {
            SecGeomHost Gs = G;
            Gs.standard(n);
            FH_TRY(sec_build_table(tab, Gs, upmask, dnmask, T, &table_ok));
            if (!table_ok) fh_sector_forget_table(tab->uid);
        }
    }
    if (!pool_ok && !table_ok) return FH_OK;
    const u64 dim = (u64)G.d_up() * G.d_dn();
    std::vector<unsigned> depU(G.d_up()), depD(G.d_dn());
    for (unsigned r = 0; r < G.d_up(); ++r) {
        unsigned dep = 0;
        for (int b = 0; b < nbu; ++b)
            if (G.cfgU[r] >> b & 1u) dep |= 1u << G.upbits[b];
        depU[r] = dep;
    }
    for (unsigned r = 0; r < G.d_dn(); ++r) {
        unsigned dep = 0;
        for (int b = 0; b < nbd; ++b)
            if (G.cfgD[r] >> b & 1u) dep |= 1u << G.dnbits[b];
        depD[r] = dep;
    }
    FH_TRY(sec_upload(&P->d_depU, depU));
    FH_TRY(sec_upload(&P->d_depD, depD));
    if (pool_ok) {
        FH_CUDA(cudaMalloc(&P->d_psi, sizeof(double2) * dim));
        FH_CUDA(cudaMalloc(&P->d_lam, sizeof(double2) * dim));
    }
    if (table_ok) {
        FH_CUDA(cudaMalloc(&P->d_in, sizeof(double2) * dim));
        FH_CUDA(cudaMalloc(&P->d_out, sizeof(double2) * dim));
        FH_TRY(sec_upload(&P->d_cfgU, G.cfgU));
        FH_TRY(sec_upload(&P->d_cfgD, G.cfgD));
        FH_TRY(sec_upload(&P->d_rankU, G.rankU));
        FH_TRY(sec_upload(&P->d_rankD, G.rankD));
        FH_CUDA(cudaMalloc(&P->d_k2_partials, sizeof(double) * 2 * 4096));
        FH_CUDA(cudaMalloc(&P->d_k2_counter, sizeof(unsigned)));
        FH_CUDA(cudaMemset(P->d_k2_counter, 0, sizeof(unsigned)));
    }
    P->half = G.half;
    P->d_up = G.d_up();
    P->d_dn = G.d_dn();
    P->eligible = pool_ok;
    P->table_ok = table_ok;
    (void)ctx;
    return FH_OK;
}

// ---- K2 on the compressed state: lam_c = H psi_c, E = <psi|H|psi>; optionally scattered back into a full-space vector ----
__global__ void __launch_bounds__(256) k_sector_compress1(const unsigned *__restrict__ depU, const unsigned *__restrict__ depD,
                                                          unsigned d_dn, unsigned dim, const double2 *__restrict__ a,
                                                          double2 *__restrict__ ac) {
    SEC_PDL_PROLOGUE();
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned r = blockIdx.x * blockDim.x + threadIdx.x; r < dim; r += stride) {
        const unsigned ru = r / d_dn, rd = r - ru * d_dn;
        ac[r] = a[__ldg(depU + ru) | __ldg(depD + rd)];
    }
}
__global__ void __launch_bounds__(256) k_sector_scatter1(const unsigned *__restrict__ depU, const unsigned *__restrict__ depD,
                                                         unsigned d_dn, unsigned dim, const double2 *__restrict__ ac,
                                                         double2 *__restrict__ a) {
    SEC_PDL_PROLOGUE();
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned r = blockIdx.x * blockDim.x + threadIdx.x; r < dim; r += stride) {
        const unsigned ru = r / d_dn, rd = r - ru * d_dn;
        a[__ldg(depU + ru) | __ldg(depD + rd)] = ac[r];
    }
}

// gather over the x-mask groups in compact coordinates (as the H step of k_sector_eval, on global compressed vectors).  Four
// lanes share an amplitude and split its groups (g = lane, lane + 4, ...): four times the parallelism and a quarter of the
// dependent L2 round trips per thread; the four partial sums are folded by shuffles in a fixed order.  Per-CTA energy
// partials, folded by the last CTA in slot order.
// SPLIT lanes share an amplitude: 4 for small sectors (parallelism, shorter dependent chains), 1 for large ones (a warp is then
// 32 consecutive ranks applying the same group: the partner loads of an up-hop are one contiguous 512-byte run)
template <bool REAL, int SPLIT>
__global__ void __launch_bounds__(256) k_sector_happly(const SecGroup *__restrict__ groups, int ngroups, const SecClass *__restrict__ classes,
                                                       const double2 *__restrict__ vals, const double2 *__restrict__ hdiag,
                                                       const unsigned short *__restrict__ cfgU, const unsigned short *__restrict__ cfgD,
                                                       const unsigned short *__restrict__ rankU, const unsigned short *__restrict__ rankD,
                                                       unsigned d_dn, unsigned dim, const double2 *__restrict__ in,
                                                       double2 *__restrict__ out, double *__restrict__ partials,
                                                       unsigned *__restrict__ counter, double *__restrict__ result) {
    SEC_PDL_PROLOGUE();
    __shared__ double red[16];
    __shared__ unsigned is_last;
    double e_re = 0.0, e_im = 0.0;
    const unsigned sub = threadIdx.x & (SPLIT - 1);
    const unsigned per_cta = 256 / SPLIT;
    const unsigned stride = gridDim.x * per_cta;
    const unsigned rounds = (dim + stride - 1) / stride;          // uniform trip count: the shuffles need whole warps
    for (unsigned it = 0; it < rounds; ++it) {
        const unsigned r = it * stride + blockIdx.x * per_cta + (threadIdx.x / SPLIT);
        const bool ok = r < dim;
        const unsigned rr = ok ? r : 0u;
        const unsigned ru = rr / d_dn, rd = rr - ru * d_dn;
        const unsigned cfg = (unsigned)__ldg(cfgU + ru) | ((unsigned)__ldg(cfgD + rd) << 16);
        double ar = 0.0, ai = 0.0;
        for (int g = (int)sub; g < ngroups; g += SPLIT) {
            const uint4 g0 = __ldg(reinterpret_cast<const uint4 *>(groups + g));          // x, first_class, n_class, live
            const uint4 g1 = __ldg(reinterpret_cast<const uint4 *>(groups + g) + 1);      // pos[4], kbits, zeta1, vofs1
            const unsigned gp = g1.x;
            const unsigned j = cfg ^ g0.x;
            const unsigned pat = ((j >> (gp & 0xffu)) & 1u) | (((j >> ((gp >> 8) & 0xffu)) & 1u) << 1) |
                                 (((j >> ((gp >> 16) & 0xffu)) & 1u) << 2) | (((j >> (gp >> 24)) & 1u) << 3);
            if (!((g0.w >> pat) & 1u)) continue;
            double wr = 0.0, wi = 0.0;
            if (g0.z == 1u) {               // one class: zeta and the table offset sit in the group record
                const double2 w = __ldg(vals + g1.w + pat);
                const bool neg = (__popc(j & g1.z) & 1) != 0;
                wr = neg ? -w.x : w.x;
                if (!REAL) wi = neg ? -w.y : w.y;
            } else {
                for (int c = (int)g0.y; c < (int)(g0.y + g0.z); ++c) {
                    const uint2 cl = __ldg(reinterpret_cast<const uint2 *>(classes + c));      // zeta, vofs
                    const double2 w = __ldg(vals + cl.y + pat);
                    const bool neg = (__popc(j & cl.x) & 1) != 0;
                    wr += neg ? -w.x : w.x;
                    if (!REAL) wi += neg ? -w.y : w.y;
                }
            }
            // partner rank: a single-species x-mask keeps the other species' rank (one table lookup instead of two)
            const unsigned sp = (g1.y >> 8) & 3u;
            const unsigned qu = sp == 2u ? ru : (unsigned)__ldg(rankU + (j & 0xffffu));
            const unsigned qd = sp == 1u ? rd : (unsigned)__ldg(rankD + (j >> 16));
            const double2 pv = in[qu * d_dn + qd];
            if (REAL) {
                ar += wr * pv.x;
                ai += wr * pv.y;
            } else {
                ar += wr * pv.x - wi * pv.y;
                ai += wr * pv.y + wi * pv.x;
            }
        }
#pragma unroll
        for (int o = 1; o < SPLIT; o <<= 1) {
            ar += __shfl_xor_sync(0xffffffffu, ar, o);
            ai += __shfl_xor_sync(0xffffffffu, ai, o);
        }
        if (ok && sub == 0) {
            const double2 self = in[r];
            const double2 hd = __ldg(hdiag + r);
            ar += hd.x * self.x - hd.y * self.y;
            ai += hd.x * self.y + hd.y * self.x;
            out[r] = make_double2(ar, ai);
            e_re += self.x * ar + self.y * ai;
            e_im += self.x * ai - self.y * ar;
        }
    }
    sec_block_sum2(e_re, e_im, red);
    if (threadIdx.x == 0) {
        partials[2 * blockIdx.x] = e_re;
        partials[2 * blockIdx.x + 1] = e_im;
        __threadfence();
        is_last = (atomicAdd(counter, 1u) == gridDim.x - 1u) ? 1u : 0u;
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double r = 0.0, im = 0.0;
        for (unsigned t = threadIdx.x; t < gridDim.x; t += blockDim.x) {
            r += __ldcg(partials + 2 * t);
            im += __ldcg(partials + 2 * t + 1);
        }
        sec_block_sum2(r, im, red);
        if (threadIdx.x == 0) {
            result[0] = r;
            result[1] = im;
            *counter = 0u;
        }
    }
}

template <bool REAL>
static void sec_launch_happly(cudaStream_t st, const SecTableCache &T, const fh_sector_pool_plan *P, unsigned dim, const double2 *in,
                              double2 *out, double *d_result) {
    const bool split = dim < (1u << 17);
    const unsigned per_cta = split ? 64u : 256u;
    unsigned hgrid = (dim + per_cta - 1u) / per_cta;
    if (hgrid > 4096u) hgrid = 4096u;                 // the energy partial array
    if (hgrid < 1u) hgrid = 1u;
    if (split)
        sec_launch(k_sector_happly<REAL, 4>, dim3(hgrid), dim3(256), 0, st, (const SecGroup *)T.d_groups, T.ngroups, (const SecClass *)T.d_classes,
                   (const double2 *)T.d_vals, (const double2 *)T.d_hdiag, (const unsigned short *)P->d_cfgU, (const unsigned short *)P->d_cfgD,
                   (const unsigned short *)P->d_rankU, (const unsigned short *)P->d_rankD, P->d_dn, dim, in, out, P->d_k2_partials,
                   P->d_k2_counter, d_result);
    else
        sec_launch(k_sector_happly<REAL, 1>, dim3(hgrid), dim3(256), 0, st, (const SecGroup *)T.d_groups, T.ngroups, (const SecClass *)T.d_classes,
                   (const double2 *)T.d_vals, (const double2 *)T.d_hdiag, (const unsigned short *)P->d_cfgU, (const unsigned short *)P->d_cfgD,
                   (const unsigned short *)P->d_rankU, (const unsigned short *)P->d_rankD, P->d_dn, dim, in, out, P->d_k2_partials,
                   P->d_k2_counter, d_result);
}

// out (full space, may be NULL) <- H in; E -> d_result[0..1].  `in` must be confined to the plan's sector.
int fh_sector_table_enqueue(fh_sector_pool_plan *P, fh_ctx *ctx, const fh_table *tab, const double2 *in, double2 *out,
                            double *d_result) {
    SEC_LOCK;
    const SecTableCache &T = g_sec_tables[tab->uid];
    const unsigned dim = P->d_up * P->d_dn;
    unsigned cgrid = (dim + 255u) / 256u;
    if (cgrid > (unsigned)ctx->sm_count * 8u) cgrid = (unsigned)ctx->sm_count * 8u;
    ++g_fh_launch_count;
    k_sector_compress1<<<cgrid, 256, 0, ctx->stream>>>(P->d_depU, P->d_depD, P->d_dn, dim, in, P->d_in);
    if (tab->all_real) sec_launch_happly<true>(ctx->stream, T, P, dim, P->d_in, P->d_out, d_result);
    else sec_launch_happly<false>(ctx->stream, T, P, dim, P->d_in, P->d_out, d_result);
    if (out) {
        FH_CUDA(cudaMemsetAsync(out, 0, sizeof(double2) << P->n, ctx->stream));
        ++g_fh_launch_count;
        k_sector_scatter1<<<cgrid, 256, 0, ctx->stream>>>(P->d_depU, P->d_depD, P->d_dn, dim, P->d_out, out);
    }
    FH_CUDA(cudaGetLastError());
    return FH_OK;
}

static bool g_sec_pool_attr[64];
// pool outputs o in [first, first+count) of full-space states psi / lam -> d_pool_out[o]
int fh_sector_pool_enqueue(fh_sector_pool_plan *P, fh_ctx *ctx, const double2 *psi, const double2 *lam, const fh_pool *pool,
                           int pool_first, int pool_count, double *d_pool_out) {
    SEC_LOCK;
    if (pool_count <= 0) return FH_OK;
    const SecPoolCache &Pc = g_sec_pools[pool->uid];
    const unsigned dim = P->d_up * P->d_dn;
    unsigned cgrid = (dim + 255u) / 256u;
    if (cgrid > (unsigned)ctx->sm_count * 8u) cgrid = (unsigned)ctx->sm_count * 8u;
    ++g_fh_launch_count;
    k_sector_compress2<<<cgrid, 256, 0, ctx->stream>>>(P->d_depU, P->d_depD, P->d_dn, dim, psi, lam, P->d_psi, P->d_lam);
    const int e0 = pool->out_first[pool_first], e1 = pool->out_first[pool_first + pool_count];
    int grid = e1 - e0;
    if (grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
    if (grid < 1) grid = 1;
    const size_t smem = sizeof(unsigned) * std::max(1u, Pc.max_words);
    if (smem > 48 * 1024) {
        int dev = 0;
        cudaGetDevice(&dev);
        dev &= 63;
        if (!g_sec_pool_attr[dev]) {
            FH_CUDA(cudaFuncSetAttribute(k_sector_pool, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
            g_sec_pool_attr[dev] = true;
        }
    }
    ++g_fh_launch_count;
    k_sector_pool<<<grid, 256, smem, ctx->stream>>>(Pc.d_entries, Pc.d_lists, e0, e1, P->d_dn, P->d_psi, P->d_lam, Pc.d_partial,
                                                    pool->d_out_first, pool_first, pool_count, d_pool_out, Pc.d_counter);
    FH_CUDA(cudaGetLastError());
    return FH_OK;
}

// ---- the same screening as a stand-alone call on caller-owned full-space states --------------------------------
static std::map<u64, fh_sector_pool_plan *> g_sec_pool_plans;      // per pool handle (freed with the pool)
void fh_sector_forget_pool_plan(u64 uid) {
    SEC_LOCK;
    auto it = g_sec_pool_plans.find(uid);
    if (it != g_sec_pool_plans.end()) {
        fh_sector_pool_plan_free(it->second);
        g_sec_pool_plans.erase(it);
    }
}

static int pool_gradients_sector_impl(const char *who, const fh_pool *pool, const fh_state *psi, const fh_state *lambda, u64 upmask,
                                      u64 dnmask, int n_up, int n_dn, int first, int count, double *out) {
    SEC_LOCK;
    FH_REQUIRE(pool && psi && lambda, "%s: NULL argument", who);
    FH_REQUIRE(psi->n == pool->n && lambda->n == pool->n, "%s: qubit count mismatch", who);
    FH_REQUIRE(first >= 0 && count >= 0 && first + count <= pool->n_out, "%s: range [%d, %d) outside pool of %d", who, first,
               first + count, pool->n_out);
    if (count == 0) return FH_OK;
    fh_ctx *ctx = pool->ctx;
    FH_CUDA(cudaSetDevice(ctx->device));
    fh_sector_pool_plan *&P = g_sec_pool_plans[pool->uid];
    fh_table dummy;          // no ops, no observable to check: the caller vouches for the states
    dummy.uid = ~0ull;
    dummy.n = pool->n;
    FH_TRY(fh_sector_pool_prepare(&P, ctx, pool->n, upmask, dnmask, n_up, n_dn, std::vector<PairOp>(), std::vector<SecFlatOp>(),
                                  &dummy, pool));
    FH_REQUIRE(P->eligible, "%s: masks / particle numbers invalid, an entry does not conserve (N_up, N_dn) = (%d, %d), or the sector "
               "is too large (> 16 384 patterns per spin)", who, n_up, n_dn);
    FH_TRY(fh_sector_pool_enqueue(P, ctx, psi->d, lambda->d, pool, first, count, pool->d_out));
    if (!out) return FH_OK;          // enqueue only (timing)
    FH_CUDA(cudaMemcpyAsync(pool->h_out, pool->d_out + first, sizeof(double) * count, cudaMemcpyDeviceToHost, ctx->stream));
    FH_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(out, pool->h_out, sizeof(double) * count);
    return FH_OK;
}

extern "C" int fh_pool_gradients_sector(const fh_pool *pool, const fh_state *psi, const fh_state *lambda, int n_up, int n_dn,
                                        int first, int count, double *out) {
    return pool_gradients_sector_impl("fh_pool_gradients_sector", pool, psi, lambda, 0, 0, n_up, n_dn, first, count, out);
}

extern "C" int fh_pool_gradients_sector_masks(const fh_pool *pool, const fh_state *psi, const fh_state *lambda, uint64_t up_mask,
                                              uint64_t dn_mask, int n_up, int n_dn, int first, int count, double *out) {
    FH_REQUIRE(up_mask != 0 || dn_mask != 0, "fh_pool_gradients_sector_masks: both masks empty");
    return pool_gradients_sector_impl("fh_pool_gradients_sector_masks", pool, psi, lambda, up_mask, dn_mask, n_up, n_dn, first, count,
                                      out);
}

// ---- K2 in the sector as a stand-alone call --------------------------------------------------------------------
static std::map<u64, fh_sector_pool_plan *> g_sec_table_plans;      // per table handle (freed with the table)
void fh_sector_forget_table_plan(u64 uid) {
    SEC_LOCK;
    auto it = g_sec_table_plans.find(uid);
    if (it != g_sec_table_plans.end()) {
        fh_sector_pool_plan_free(it->second);
        g_sec_table_plans.erase(it);
    }
}

extern "C" int fh_apply_table_sector(const fh_table *tab, const fh_state *in, fh_state *out, int n_up, int n_dn, double *e_re,
                                     double *e_im) {
    SEC_LOCK;
    FH_REQUIRE(tab && in, "fh_apply_table_sector: NULL argument");
    FH_REQUIRE(in->n == tab->n && (!out || out->n == tab->n), "fh_apply_table_sector: qubit count mismatch");
    FH_REQUIRE(!out || out->d != in->d, "fh_apply_table_sector: in and out must differ");
    fh_ctx *ctx = tab->ctx;
    FH_CUDA(cudaSetDevice(ctx->device));
    fh_sector_pool_plan *&P = g_sec_table_plans[tab->uid];
    FH_TRY(fh_sector_pool_prepare(&P, ctx, tab->n, 0, 0, n_up, n_dn, std::vector<PairOp>(), std::vector<SecFlatOp>(), tab, nullptr));
    FH_REQUIRE(P->table_ok, "fh_apply_table_sector: the table does not conserve (N_up, N_dn) = (%d, %d), has a term group with more "
               "than four X/Y factors, or the qubit count is odd / above 30", n_up, n_dn);
    FH_TRY(fh_sector_table_enqueue(P, ctx, tab, in->d, out ? out->d : nullptr, ctx->d_result));
    if (!e_re && !e_im) return FH_OK;        // enqueue only (timing)
    FH_CUDA(cudaMemcpyAsync(ctx->h_result, ctx->d_result, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FH_CUDA(cudaStreamSynchronize(ctx->stream));
    if (e_re) *e_re = ctx->h_result[0];
    if (e_im) *e_im = ctx->h_result[1];
    return FH_OK;
}

// ----------------------------------------------------------------------------------------------
// Dense tail: a trailing run of FIXED single-species ops (the basis change W of the ADAPT / HVA drivers) as two dense
// sector transforms.  In the blocked orbital order (all up, then all down) an up-only op is U (x) 1; the interleaved
// Jordan-Wigner order differs from it by the diagonal sign S(u, d) = (-1)^#{(i, j): up orbital i and down orbital j occupied,
// j <= i}, so on the sector matrix Psi[ru][rd]
//        W Psi = S o ( U_up (S o Psi) U_dn^T ),        W^dagger Psi = S o ( U_up^dagger (S o Psi) conj(U_dn) ).
// U_up / U_dn (D_up x D_up, D_dn x D_dn) are accumulated on the host once per program from the ops' up / down lists; every op
// is checked for the sign structure the factorisation needs (its other-species zeta bits must be exactly the crossing mask).
// Then the whole tail of a screening -- W, H, W^dagger, K3 -- runs on compressed vectors: 2 + 1 + 2 + 1 small launches.
// ----------------------------------------------------------------------------------------------
struct SecDense {
    bool ok = false;
    int first_flat = 0;                       // the tail = flat ops [first_flat, end)
    double2 *d_Uup = nullptr, *d_UdnT = nullptr;        // W:        A = U_up,        B = U_dn^T
    double2 *d_UupH = nullptr, *d_UdnC = nullptr;       // W^dagger: A = U_up^dagger, B = conj(U_dn)
    unsigned short *d_pp = nullptr;           // per down pattern: prefix parities (sign S = parity(cfgU & pp))
    double2 *d_t0 = nullptr, *d_t1 = nullptr; // compressed work vectors
    void release() {
        cudaFree(d_Uup); cudaFree(d_UdnT); cudaFree(d_UupH); cudaFree(d_UdnC); cudaFree(d_pp); cudaFree(d_t0); cudaFree(d_t1);
        d_Uup = d_UdnT = d_UupH = d_UdnC = d_t0 = d_t1 = nullptr; d_pp = nullptr;
        ok = false;
    }
};

// C[M x N] = A[M x K] * B'[K x N];  SIGN_B: B'[k][n] = S(k, n) B[k][n] (rows of B are up ranks);  SIGN_C: C is multiplied by
// S(m, n) on store (rows of C are up ranks).  16 x 16 output tile per CTA; K in chunks of 64 whose 2 x 1024 elements are all
// requested before the first is used (one memory latency per chunk: the matrices are a few hundred KB and this runs right
// after other kernels, so the loop is latency-, not flop-bound).
#define SEC_GK 64
#define SEC_GCH 2            // K chunks whose loads are all issued before the first use (K <= 128: one memory latency per GEMM)
// GATHER: B is a FULL-SPACE state read through the sector index (depU[k] | depD[n]) -- the compress step fused into the first
// GEMM of the tail; the row-block-0 CTAs also write the compressed copy Bc (K3 needs psi_s).
template <bool SIGN_B, bool SIGN_C, bool GATHER>
__global__ void __launch_bounds__(256) k_sector_gemm(const double2 *__restrict__ A, const double2 *__restrict__ B, double2 *__restrict__ C,
                                                     int M, int N, int K, const unsigned short *__restrict__ cfgU,
                                                     const unsigned short *__restrict__ pp, const unsigned *__restrict__ depU,
                                                     const unsigned *__restrict__ depD, double2 *__restrict__ Bc) {
    SEC_PDL_PROLOGUE();
    __shared__ double2 sa[16][SEC_GK + 1], sb[SEC_GK][17];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * 16, n0 = blockIdx.x * 16;
    const int m = m0 + ty, n = n0 + tx;
    const unsigned ppn = (n < N && SIGN_C) ? (unsigned)__ldg(pp + n) : 0u;
    double cr = 0.0, ci = 0.0;
    for (int kbase = 0; kbase < K; kbase += SEC_GK * SEC_GCH) {
        double2 ra[SEC_GCH][4], rb[SEC_GCH][4];
#pragma unroll
        for (int ch = 0; ch < SEC_GCH; ++ch) {
            const int k0 = kbase + ch * SEC_GK;
#pragma unroll
            for (int q = 0; q < 4; ++q) {          // A tile: 16 rows x 64 k, element e = q * 256 + tid -> (row e / 64, k e % 64)
                const int e = q * 256 + (int)threadIdx.x, row = e >> 6, kk = e & 63;
                ra[ch][q] = (m0 + row < M && k0 + kk < K) ? A[(size_t)(m0 + row) * K + k0 + kk] : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {          // B tile: 64 k x 16 columns, element e -> (k e / 16, column e % 16)
                const int e = q * 256 + (int)threadIdx.x, kk = e >> 4, col = e & 15;
                const bool ok = k0 + kk < K && n0 + col < N;
                double2 v = make_double2(0.0, 0.0);
                if (ok) {
                    if (GATHER) {
                        v = B[__ldg(depU + k0 + kk) | __ldg(depD + n0 + col)];
                        if (blockIdx.y == 0) Bc[(size_t)(k0 + kk) * N + n0 + col] = v;
                    } else {
                        v = B[(size_t)(k0 + kk) * N + n0 + col];
                    }
                    if (SIGN_B && (__popc((unsigned)__ldg(cfgU + k0 + kk) & (unsigned)__ldg(pp + n0 + col)) & 1)) v = make_double2(-v.x, -v.y);
                }
                rb[ch][q] = v;
            }
        }
#pragma unroll
        for (int ch = 0; ch < SEC_GCH; ++ch) {
            if (kbase + ch * SEC_GK >= K) break;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int e = q * 256 + (int)threadIdx.x;
                sa[e >> 6][e & 63] = ra[ch][q];
                sb[e >> 4][e & 15] = rb[ch][q];
            }
            __syncthreads();
#pragma unroll 16
            for (int k = 0; k < SEC_GK; ++k) {
                const double2 a = sa[ty][k], bb = sb[k][tx];
                cr += a.x * bb.x - a.y * bb.y;
                ci += a.x * bb.y + a.y * bb.x;
            }
            __syncthreads();
        }
    }
    if (m < M && n < N) {
        if (SIGN_C && (__popc((unsigned)__ldg(cfgU + m) & ppn) & 1)) {
            cr = -cr;
            ci = -ci;
        }
        C[(size_t)m * N + n] = make_double2(cr, ci);
    }
}

// out = W in (dagger: W^dagger in) on compressed vectors; tmp is a third vector
static void sec_dense_apply(const fh_sector_pool_plan *P, const SecDense &D, cudaStream_t s, const double2 *in, double2 *tmp, double2 *out,
                            bool dagger, double2 *gather_copy);

struct fh_sector_dense {
    SecDense D;
    int tail_item = -1, n_up = -1, n_dn = -1;
    u64 pool_plan_key = 0;
    bool built = false;
};
void fh_sector_dense_free(fh_sector_dense *d) {
    if (!d) return;
    d->D.release();
    delete d;
}
bool fh_sector_dense_ok(const fh_sector_dense *d) { return d && d->D.ok; }
int fh_sector_dense_tail_item(const fh_sector_dense *d) { return d ? d->tail_item : -1; }
int fh_sector_dense_first_flat(const fh_sector_dense *d) { return d ? d->D.first_flat : -1; }

typedef std::vector<double2> SecMatH;          // row-major square matrix
static void sec_mat_rows_pair(SecMatH &U, unsigned d, const std::vector<unsigned> &L, const double m[8]) {
    // rows r (pattern side) and r' of U <- the 2x2 block, with the op's own-species sign
    for (unsigned e : L) {
        const unsigned r = e & SEC_FM, rp = (e >> SEC_FB) & SEC_FM;
        const double sg = (e >> 31) ? -1.0 : 1.0;
        const double2 m00 = make_double2(m[0], m[1]), m01 = make_double2(sg * m[2], sg * m[3]), m10 = make_double2(sg * m[4], sg * m[5]),
                      m11 = make_double2(m[6], m[7]);
        double2 *a = &U[(size_t)r * d], *b = &U[(size_t)rp * d];
        for (unsigned c = 0; c < d; ++c) {
            const double2 x = a[c], y = b[c];
            a[c] = make_double2(m00.x * x.x - m00.y * x.y + m01.x * y.x - m01.y * y.y, m00.x * x.y + m00.y * x.x + m01.x * y.y + m01.y * y.x);
            b[c] = make_double2(m10.x * x.x - m10.y * x.y + m11.x * y.x - m11.y * y.y, m10.x * x.y + m10.y * x.x + m11.x * y.y + m11.y * y.x);
        }
    }
}

// Build (or reuse) the dense tail of a program.  min_item: the tail may not begin before this item; exact: it must begin
// exactly there (screening: the pool position).  Leaves D.ok = false when the program has no such tail.
int fh_sector_dense_prepare(fh_sector_dense **slot, const fh_sector_pool_plan *P, int n, const std::vector<PairOp> &pairs,
                            const std::vector<DiagOp> &diagops, const std::vector<DiagTerm> &dterms, const std::vector<SecFlatOp> &flat,
                            const std::vector<int> &item_flat_first, int min_item, bool exact) {
    if (!*slot) *slot = new fh_sector_dense();
    fh_sector_dense *X = *slot;
    const int n_items = (int)item_flat_first.size() - 1;
    if (!P || !P->table_ok || getenv("FHSIM_NO_SECTOR_DENSE")) {
        X->D.release();
        X->built = false;
        return FH_OK;
    }
    SecGeomHost G;
    G.standard(n);
    const int half = G.half;
    const unsigned upc = (1u << half) - 1u;
    // per flat op: 1 up-only, 2 down-only, 0 not eligible
    auto classify = [&](const SecFlatOp &f) -> int {
        if (f.type == 1) {
            const PairOp &op = pairs[f.index];
            if (op.kind != 0) return 0;
            const unsigned xc = sec_compact(op.x, G), zc = sec_compact(op.zeta, G), fmc = sec_compact(op.fixmask, G);
            const unsigned xu = xc & 0xffffu, xd = xc >> 16;
            if ((xu != 0u) == (xd != 0u)) return 0;
            if (xu ? (fmc >> 16) != 0u : (fmc & 0xffffu) != 0u) return 0;       // pattern conditioned on the other species
            if (xu) {          // other-species zeta = down orbitals j with an odd number of flipped up orbitals i >= j
                unsigned need = 0;
                for (int j = 0; j < half; ++j)
                    if (__builtin_popcount(xu & (upc & ~((1u << j) - 1u))) & 1) need |= 1u << j;
                return (zc >> 16) == need ? 1 : 0;
            }
            unsigned need = 0;  // up orbitals i with an odd number of flipped down orbitals j <= i
            for (int i = 0; i < half; ++i)
                if (__builtin_popcount(xd & ((2u << i) - 1u)) & 1) need |= 1u << i;
            return (zc & 0xffffu) == need ? 2 : 0;
        }
        const DiagOp &dop = diagops[f.index];
        if (dop.param >= 0) return 0;
        for (int m = dop.first; m < dop.first + dop.count; ++m) {
            const unsigned zc = sec_compact(dterms[m].z, G);
            if ((zc & 0xffffu) && (zc >> 16)) return 0;
        }
        return 3;               // diagonal, every term inside one species
    };
    int i_min = n_items;
    while (i_min > 0) {
        bool ok = true;
        for (int k = item_flat_first[i_min - 1]; k < item_flat_first[i_min] && ok; ++k) ok = classify(flat[k]) != 0;
        if (!ok) break;
        --i_min;
    }
    int tail_item = i_min < min_item ? min_item : i_min;
    if (exact && tail_item != min_item) tail_item = -1;
    const u64 key = ((u64)P->table_uid << 20) ^ ((u64)(unsigned)P->n_up << 8) ^ (u64)(unsigned)P->n_dn;
    if (X->built && X->tail_item == tail_item && X->pool_plan_key == key) return FH_OK;
    X->D.release();
    X->built = true;
    X->tail_item = tail_item;
    X->pool_plan_key = key;
    if (tail_item < 0) return FH_OK;
    const unsigned d_up = P->d_up, d_dn = P->d_dn;
    if (d_up > 512 || d_dn > 512) return FH_OK;           // dense blocks only while they are small (3x3: 126 x 126)
    G.n_up = P->n_up;
    G.n_dn = P->n_dn;
    sec_patterns(half, P->n_up, G.cfgU, G.rankU);
    sec_patterns(half, P->n_dn, G.cfgD, G.rankD);
    SecMatH Uu((size_t)d_up * d_up, make_double2(0.0, 0.0)), Ud((size_t)d_dn * d_dn, make_double2(0.0, 0.0));
    for (unsigned r = 0; r < d_up; ++r) Uu[(size_t)r * d_up + r].x = 1.0;
    for (unsigned r = 0; r < d_dn; ++r) Ud[(size_t)r * d_dn + r].x = 1.0;
    std::vector<unsigned> LU, LD;
    for (int k = item_flat_first[tail_item]; k < item_flat_first[n_items]; ++k) {
        const SecFlatOp &f = flat[k];
        const int cls = classify(f);
        if (cls == 1 || cls == 2) {
            const PairOp &op = pairs[f.index];
            sec_build_lists(G, sec_compact(op.x, G), sec_compact(op.fixmask, G), sec_compact(op.fixval, G), sec_compact(op.zeta, G), LU, LD);
            if (cls == 1) sec_mat_rows_pair(Uu, d_up, LU, op.m);
            else sec_mat_rows_pair(Ud, d_dn, LD, op.m);
        } else {
            const DiagOp &dop = diagops[f.index];
            for (int m = dop.first; m < dop.first + dop.count; ++m) {
                const unsigned zc = sec_compact(dterms[m].z, G);
                const bool up = (zc >> 16) == 0u;             // z = 0 (a plain phase) goes to the up block
                const unsigned zs = up ? (zc & 0xffffu) : (zc >> 16);
                SecMatH &U = up ? Uu : Ud;
                const unsigned d = up ? d_up : d_dn;
                const double c = cos(dterms[m].angle), sn = sin(dterms[m].angle);
                for (unsigned r = 0; r < d; ++r) {
                    const unsigned cfg = up ? G.cfgU[r] : G.cfgD[r];
                    const double2 ph = make_double2(c, (__builtin_popcount(cfg & zs) & 1) ? sn : -sn);      // exp(-i a sigma)
                    double2 *row = &U[(size_t)r * d];
                    for (unsigned q = 0; q < d; ++q) row[q] = make_double2(row[q].x * ph.x - row[q].y * ph.y, row[q].x * ph.y + row[q].y * ph.x);
                }
            }
        }
    }
    // device forms: W: A = U_up, B = U_dn^T;  W^dagger: A = U_up^dagger, B = conj(U_dn)
    SecMatH UuH((size_t)d_up * d_up), UdT((size_t)d_dn * d_dn), UdC((size_t)d_dn * d_dn);
    for (unsigned a = 0; a < d_up; ++a)
        for (unsigned b = 0; b < d_up; ++b) UuH[(size_t)a * d_up + b] = make_double2(Uu[(size_t)b * d_up + a].x, -Uu[(size_t)b * d_up + a].y);
    for (unsigned a = 0; a < d_dn; ++a)
        for (unsigned b = 0; b < d_dn; ++b) {
            UdT[(size_t)a * d_dn + b] = Ud[(size_t)b * d_dn + a];
            UdC[(size_t)a * d_dn + b] = make_double2(Ud[(size_t)a * d_dn + b].x, -Ud[(size_t)a * d_dn + b].y);
        }
    std::vector<unsigned short> pp(d_dn);
    for (unsigned r = 0; r < d_dn; ++r) {
        unsigned v = 0;
        for (int i = 0; i < half; ++i)
            if (__builtin_popcount((unsigned)G.cfgD[r] & ((2u << i) - 1u)) & 1) v |= 1u << i;
        pp[r] = (unsigned short)v;
    }
    SecDense &D = X->D;
    FH_TRY(sec_upload(&D.d_Uup, Uu));
    FH_TRY(sec_upload(&D.d_UdnT, UdT));
    FH_TRY(sec_upload(&D.d_UupH, UuH));
    FH_TRY(sec_upload(&D.d_UdnC, UdC));
    FH_TRY(sec_upload(&D.d_pp, pp));
    const size_t dim = (size_t)d_up * d_dn;
    FH_CUDA(cudaMalloc(&D.d_t0, sizeof(double2) * dim));
    FH_CUDA(cudaMalloc(&D.d_t1, sizeof(double2) * dim));
    D.first_flat = item_flat_first[tail_item];
    D.ok = true;
    return FH_OK;
}

static void sec_dense_apply(const fh_sector_pool_plan *P, const SecDense &D, cudaStream_t s, const double2 *in, double2 *tmp, double2 *out,
                            bool dagger, double2 *gather_copy) {
    const int du = (int)P->d_up, dd = (int)P->d_dn;
    const dim3 grid((dd + 15) / 16, (du + 15) / 16);
    // tmp = A (S o in);  out = S o (tmp B).  gather_copy != NULL: `in` is a full-space state, its compressed copy is written there
    const unsigned *nou = nullptr;
    double2 *nod = nullptr;
    if (gather_copy)
        sec_launch(k_sector_gemm<true, false, true>, grid, dim3(256), 0, s, (const double2 *)(dagger ? D.d_UupH : D.d_Uup), in, tmp, du, dd, du,
                   (const unsigned short *)P->d_cfgU, (const unsigned short *)D.d_pp, (const unsigned *)P->d_depU, (const unsigned *)P->d_depD,
                   gather_copy);
    else
        sec_launch(k_sector_gemm<true, false, false>, grid, dim3(256), 0, s, (const double2 *)(dagger ? D.d_UupH : D.d_Uup), in, tmp, du, dd, du,
                   (const unsigned short *)P->d_cfgU, (const unsigned short *)D.d_pp, nou, nou, nod);
    sec_launch(k_sector_gemm<false, true, false>, grid, dim3(256), 0, s, (const double2 *)tmp, (const double2 *)(dagger ? D.d_UdnC : D.d_UdnT), out,
               du, dd, dd, (const unsigned short *)P->d_cfgU, (const unsigned short *)D.d_pp, nou, nou, nod);
}

double2 *fh_sector_dense_psi_buffer(fh_sector_pool_plan *P, const fh_pool *pool) { return (pool && P->d_psi) ? P->d_psi : P->d_in; }

// psi_full == NULL: the compressed psi_s is already in fh_sector_dense_psi_buffer() (written by the cluster kernel)
int fh_sector_dense_enqueue(fh_sector_dense *X, fh_sector_pool_plan *P, fh_ctx *ctx, const fh_table *tab, const double2 *psi_full,
                            double *d_result, const fh_pool *pool, int pool_first, int pool_count, double *d_pool_out) {
    SEC_LOCK;
    const SecDense &D = X->D;
    const SecTableCache &T = g_sec_tables[tab->uid];
    const unsigned dim = P->d_up * P->d_dn;
    unsigned cgrid = (dim + 255u) / 256u;
    if (cgrid > (unsigned)ctx->sm_count * 8u) cgrid = (unsigned)ctx->sm_count * 8u;
    // psi_s (compressed) lives in the K3 buffer when a pool is screened, else in the K2 input buffer
    double2 *psi_s = (pool && P->d_psi) ? P->d_psi : P->d_in;
    (void)cgrid;
    if (psi_full) sec_dense_apply(P, D, ctx->stream, psi_full, D.d_t0, D.d_t1, false, psi_s);   // phi = W psi_s (t1); psi_s compressed on the way
    else sec_dense_apply(P, D, ctx->stream, psi_s, D.d_t0, D.d_t1, false, nullptr);
    if (tab->all_real) sec_launch_happly<true>(ctx->stream, T, P, dim, D.d_t1, P->d_out, d_result);          // H phi (d_out)
    else sec_launch_happly<false>(ctx->stream, T, P, dim, D.d_t1, P->d_out, d_result);
    if (pool && pool_count > 0) {
        sec_dense_apply(P, D, ctx->stream, P->d_out, D.d_t0, P->d_lam, true, nullptr); // lambda_s = W^dagger H phi
        const SecPoolCache &Pc = g_sec_pools[pool->uid];
        const int e0 = pool->out_first[pool_first], e1 = pool->out_first[pool_first + pool_count];
        int grid = e1 - e0;
        if (grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
        if (grid < 1) grid = 1;
        sec_launch(k_sector_pool, dim3(grid), dim3(256), sizeof(unsigned) * std::max(1u, Pc.max_words), ctx->stream,
                   (const SecPoolEntry *)Pc.d_entries, (const unsigned *)Pc.d_lists, e0, e1, P->d_dn, (const double2 *)P->d_psi,
                   (const double2 *)P->d_lam, Pc.d_partial, (const int *)pool->d_out_first, pool_first, pool_count, d_pool_out, Pc.d_counter);
    }
    FH_CUDA(cudaGetLastError());
    return FH_OK;
}
