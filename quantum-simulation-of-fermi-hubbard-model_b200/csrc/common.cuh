// Shared types for libfhsim (sm_100a).  Host structs mirror the device descriptors 1:1.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/fhsim.h"

typedef unsigned long long u64;

// ----------------------------------------------------------------------------------------------
// error plumbing: thread-local message, no exceptions across the ABI
// ----------------------------------------------------------------------------------------------
void fh_set_error(const char *fmt, ...);

#define FH_CUDA(call)                                                                           \
    do {                                                                                        \
        cudaError_t _e = (call);                                                                \
        if (_e != cudaSuccess) {                                                                \
            fh_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return FH_ECUDA;                                                                    \
        }                                                                                       \
    } while (0)

#define FH_REQUIRE(cond, ...)              \
    do {                                   \
        if (!(cond)) {                     \
            fh_set_error(__VA_ARGS__);     \
            return FH_EINVAL;              \
        }                                  \
    } while (0)

#define FH_TRY(call)                \
    do {                            \
        int _rc = (call);           \
        if (_rc != FH_OK) return _rc; \
    } while (0)

// ----------------------------------------------------------------------------------------------
// device descriptors (plain data, uploaded verbatim)
// ----------------------------------------------------------------------------------------------
#define FH_MAX_FIX 16      // max bits in a pair op's fixmask
#define FH_MAX_TILE_BITS 13
#define FH_TILE_MAX_SUB 96      // ops per fused tile kernel (descriptors live in shared memory)
#define FH_TILE_MAX_TERMS 256   // diagonal terms per fused tile kernel

// 2x2 op on index pairs (i, i^x) selected by (i & fixmask) == fixval, sign s = parity(i & zeta).
struct __align__(16) PairOp {
    u64 x, fixmask, fixval, zeta;       // 32
    double m[8];                        // m00 m01 m10 m11 (re,im)  64
    double bhat[2];                     // normalised generator element (rotation ops)  16
    double gscale;                      // d(angle)/d(theta); 0 for fixed ops  8
    int param;                          // parameter index or -1
    int kind;                           // 0 fixed, 1 rotation
    unsigned char npos;                 // popcount(fixmask)
    unsigned char pos[FH_MAX_FIX];      // ascending bit positions of fixmask
    unsigned char pad[7];
};

struct DiagTerm {
    u64 z;
    double angle;    // current angle (theta*coef or fixed)
    double c, s;     // cos(angle), sin(angle)
    double coef;     // d(angle)/d(theta) (0 for fixed)
};

struct DiagOp {
    int first, count;   // range in the DiagTerm array
    int param;
    int pad;
};

// one entry of a fused tile kernel's op list (host bookkeeping)
struct TileSub {
    int type;        // 1 pair, 2 diag
    int index;       // index into PairOp / DiagOp arrays
    int lpivot;      // diag: tile-local offset of its terms
    unsigned int xlocal;   // x in tile-local coordinates (pair)
};

// geometry of one fused tile kernel: passed BY VALUE as a kernel argument (constant bank, no latency)
struct TileLaunch {
    int nbits;                          // T
    int nsub;                           // records in this run
    int first_rec;                      // offset into the record array of the chosen direction
    int first_term, nterms;             // range in the tile-term array of the chosen direction
    int ptab_first, ptab_words;         // range in the pair-table array of the chosen direction (16-bit entries)
    unsigned char bits[16];             // ascending global bit positions of the tile
};

// ready-to-execute record of one op inside a tile (built on the host, one array per direction)
struct __align__(16) TileRec {          // 160 bytes = 10 x 16 B; the kernels read it as uint4 / double2 words
    // word 0
    unsigned fixmask_out, fixval_out;   // pattern bits outside the tile: uniform per tile
    unsigned zeta, xlocal;
    // word 1
    unsigned lfixval;                   // pattern bits inside the tile, in tile-local coordinates
    int type;                           // 1 pair (complex matrix), 3 pair (real matrix), 6 pair (real diagonal), 2 diag
    int nlfix, term_off;                // term_off: diag ops: first term; pair ops: first word of the op in the run's pair table
    // word 2: bit-insertion masks (1<<p)-1 of the (at most 4) pattern bits inside the tile, ascending;
    // unused slots hold 0xffffffff (insertion is then a no-op)
    unsigned lowmask[4];
    // word 3: nterms (diag); seg = index of this op among the tile's parametrised ops in execution order of this
    // direction, or -1 (used by the fused adjoint sweep, which runs the dagger records)
    int nterms, seg;                    // nterms: diag ops: term count; pair ops: slot image of the x-mask (TMA layout)
    unsigned zeta_local;                // in-tile bits of zeta in tile-local coordinates (sign without an index load)
    int reps;                           // pair ops: pair-table entries per thread = ceil(pairs in a tile / threads per CTA)
    // words 4..7
    double m[8];
    // word 8: normalised generator element of a rotation op (gradient of the adjoint sweep)
    double bhat[2];
    // word 9 (TMA tile kernels): register-fused runs.  run_len >= 2 on the FIRST record of a run of Givens-like ops that all
    // live inside the three tile-local bits run_bits (b0 | b1 << 8 | b2 << 16): the run is applied to 8-amplitude groups
    // held in registers, one shared-memory round trip and one barrier for the whole run.  sub_i / x3: pattern side and
    // x-mask of THIS op in the 3-bit coordinates of its run.
    int run_len;
    unsigned run_bits;
    unsigned char sub_i, x3, run_pad[6];
};

struct __align__(16) TileTerm {         // 48 bytes
    u64 z;
    double angle, c, s;
    double coef;                        // d(angle)/d(theta) (0 for fixed terms)
    unsigned zlocal;                    // in-tile bits of z in tile-local coordinates (sign without a global index)
    unsigned pad;
};

struct TileOp {                         // host bookkeeping of one tile run
    int nbits;
    int first_sub, nsub;
    int pad;                            // running count of diagonal terms while the tile is open
    unsigned char bits[16];
    int first_rec_fwd, first_rec_dag;   // offsets into the two record arrays
    int first_term_fwd, first_term_dag, nterms;
    int n_param_subs;                   // parametrised ops inside this tile
    int first_ptab_fwd, first_ptab_dag, ptab_words;
};

struct TabTerm {
    u64 z;
    double dr, di;    // coefficient * i^k  (weight of psi[i^x] is sum_m d_m (-1)^popcount((i^x)&z_m))
};

// K2 device form.  Inside one x-mask group the terms are split into classes by zeta = z & ~x; within a class
// the weight only depends on the bits of j = i^x at the (at most 4) x positions, so it is tabulated:
//     w_g(j) = sum_c (-1)^popcount(j & zeta_c) * V_c[pattern(j)],   pattern = bits of j at pos[0..kbits)
// Groups with x = 0 or more than 4 x bits use kbits = 0 and one class per term (V = the coefficient).
struct __align__(16) TabGroup {     // 32 bytes
    u64 x;
    int first_class, n_class;
    unsigned char pos[4];           // ascending bit positions of x; unused slots = 63 (that bit of j is always 0)
    int kbits;
    unsigned rpat;                  // 8 x 4 bits: pattern(j ^ (r << 8)) = pattern(j) ^ ((rpat >> 4r) & 15)   (k_apply_table4)
    unsigned live;                  // bit p set: some class has a non-zero table entry for pattern p
};

struct __align__(16) TabClass {     // 16 bytes
    u64 zeta;
    int vofs;                       // first entry of this class's table in the V array
    int pad;
};

struct PoolEntry {
    u64 x, fixmask, fixval, zeta;
    double br, bi;
    int out;
    unsigned char npos;
    unsigned char pos[FH_MAX_FIX];
    unsigned char pad[3];
};

// K3, tiled form.  A *pass* fixes a set of T index bits; every pool entry whose x-mask lies inside those bits is
// served from psi / lambda tiles staged in shared memory, so one read of both states feeds many gradients.
#define FH_POOL_PASS_MAX_RECS 256
struct __align__(16) PoolTileRec {      // 64 bytes
    unsigned fixmask_out, fixval_out;   // pattern bits outside the tile (uniform per tile)
    unsigned zeta, xlocal;
    unsigned lfixval;                   // in-tile pattern, tile-local coordinates
    int nlfix;                          // number of in-tile pattern bits (<= 4)
    int entry;                          // index into the pool's entry array (partials row)
    int pad;
    unsigned lowmask[4];                // bit-insertion masks of the in-tile pattern bits (ascending; unused = ~0)
    double br, bi;
};

struct PoolPass {
    int nbits, first_rec, nrec, pad;
    unsigned char bits[16];
};

// ----------------------------------------------------------------------------------------------
// host objects behind the opaque handles
// ----------------------------------------------------------------------------------------------
struct fh_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    // scratch for reductions
    double *d_partials = nullptr;     // 2 * FH_MAX_PARTIALS doubles
    unsigned int *d_counter = nullptr;
    double *d_result = nullptr;       // small device result buffer (64 doubles)
    double *h_result = nullptr;       // pinned mirror
    void *d_diag = nullptr;           // factor tables of the standalone diagonal kernel (fh_diag_scratch_bytes())
    void *d_flush = nullptr;
    size_t flush_bytes = 0;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    // state-sized scratch buffers handed back by destroyed programs (fh_ctx_scratch_get / _put): cudaFree of a
    // multi-MiB block costs 4-5 ms and a driver compiles a new program every ADAPT epoch
    std::vector<std::pair<size_t, void *>> scratch_cache;
    size_t scratch_cached_bytes = 0;
};

#define FH_SCRATCH_CACHE_MAX_BYTES ((size_t)1 << 30)      // blocks above 1 GiB in total are freed, not cached
#define FH_SCRATCH_CACHE_MAX_BLOCKS 12
int fh_ctx_scratch_get(fh_ctx *ctx, size_t bytes, void **out);      // exact-size reuse, else cudaMalloc
void fh_ctx_scratch_put(fh_ctx *ctx, size_t bytes, void *ptr);      // caller has synchronised the stream

#define FH_MAX_PARTIALS 4096

// Every handle carries a process-wide unique id (never reused): cached CUDA graphs key on (address, uid), so a handle
// that was freed and re-allocated at the same host address is never mistaken for the one a graph was captured with.
u64 fh_next_uid();

struct fh_state {
    fh_ctx *ctx;
    int n;
    u64 dim;
    double2 *d;
    bool owned;
    u64 uid = fh_next_uid();
};

struct fh_table_tiles;      // table_tile.cu: cover of the x-mask groups by shared-memory tile passes (n >= 22), or NULL
struct fh_table {
    u64 uid = fh_next_uid();
    fh_table_tiles *tiles = nullptr;
    fh_ctx *ctx;
    int n;
    int n_terms, n_groups;
    bool all_real;
    TabGroup *d_groups;
    TabClass *d_classes;
    double2 *d_vals;
    double2 *d_diag = nullptr;      // diagonal part D(i) = DA[i & 4095] + DB[(i >> 12) & 4095] + DC[i >> 24], or NULL
    std::vector<TabGroup> groups;
    std::vector<TabClass> classes;
    std::vector<double2> vals;
    std::vector<TabTerm> terms;     // host copy in table order (grouped by x-mask)
    std::vector<TabTerm> diag_terms;      // the x = 0 terms (the diagonal of the observable), for the sector path
};

struct fh_pool {
    u64 uid = fh_next_uid();
    fh_ctx *ctx;
    int n;
    int n_entries, n_out;
    int chunks;                 // row width of the partials array (slots per entry)
    int kchunks;                // blocks per entry of the per-entry kernel (k_pool)
    int narrow;                 // 1: every entry pins <= 4 bits (k_pool32 applies)
    PoolEntry *d_entries;
    int *d_out_first;           // [n_out+1] entry ranges per output (entries sorted by out)
    double *d_partials;         // [n_entries * chunks]
    double *d_out;              // [n_out]
    double *h_out;              // pinned [n_out]
    std::vector<PoolEntry> entries;
    std::vector<int> out_first;
    // tiled scan: passes + their records; entries not covered by a pass go through k_pool via `rest`
    int tile_bits = 0, tile_grid = 0;
    std::vector<PoolPass> passes;
    std::vector<PoolTileRec> tile_recs;
    std::vector<int> rest;
    PoolPass *d_passes = nullptr;
    PoolTileRec *d_tile_recs = nullptr;
    int *d_rest = nullptr;
};

// ----------------------------------------------------------------------------------------------
// kernel launch wrappers (kernels.cu)
// ----------------------------------------------------------------------------------------------
void launch_pair(cudaStream_t s, int sm, double2 *psi, const PairOp *d_op, int n, int nfix, int dagger);
// adjoint step: grad partial of Ghat at (psi, lam), then apply op^dagger to both
void launch_pair_adjoint(cudaStream_t s, int sm, double2 *psi, double2 *lam, const PairOp *d_op, int n, int nfix,
                         double *d_partials, int max_blocks, int *blocks_used);
void launch_diag(cudaStream_t s, int sm, double2 *psi, const DiagTerm *d_terms, int nterms, int n, int dagger,
                 void *scratch = nullptr);
size_t fh_diag_scratch_bytes();
void launch_diag_adjoint(cudaStream_t s, int sm, double2 *psi, double2 *lam, const DiagTerm *d_terms, int nterms, int n,
                         double *d_partials, int max_blocks, int *blocks_used);
void launch_tile(cudaStream_t s, double2 *psi, const TileLaunch &tl, const TileRec *d_recs, const TileTerm *d_terms,
                 const unsigned short *d_ptab, int n);
// geometry the TMA tile kernels will use for a tile bit set: -1 TMA path not applicable, 0 linear shared-memory layout,
// 1 hardware 128-byte swizzle (slot = l ^ ((l >> 3) & 7)); threads per CTA of the tile kernels
int fh_tile_tma_layout(const unsigned char *bits, int nbits, int n);
int fh_tile_threads(int nbits);
// nl consecutive tile runs in one cooperative launch (grid barriers between runs); returns 0 if unavailable
int launch_tile_multi(cudaStream_t s, int sm, double2 *psi, const TileLaunch *d_tls, int nl, int max_bits, int min_bits,
                      const TileRec *d_recs, const TileTerm *d_terms, int n);
// fused adjoint step of a whole tile run (dagger records): gradient partials of every parametrised op into
// gpart[(seg_base + rec.seg) * FH_GRAD_BLOCKS + block], then the inverse op on psi AND lam
void launch_tile_adjoint(cudaStream_t s, double2 *psi, double2 *lam, const TileLaunch &tl, const TileRec *d_recs,
                         const TileTerm *d_terms, int n, double *d_gpart, int seg_base);
#define FH_GRAD_BLOCKS 256     // partial blocks per parametrised op in the adjoint sweep
#define FH_TILE_ADJOINT_MAX_BITS 12   // two tiles + index table must fit in shared memory
// mode 0: expectation only; 1: out = H in; 2: out += H in.   Result (re, im of <in|H|in>) -> d_result[0..1]
void launch_apply_table(cudaStream_t s, int sm, const fh_table *tab, const double2 *in, double2 *out, int mode,
                        double *d_partials, double *d_result);
int fh_table_plan_tiles(fh_table *tab);                 // called once at upload (host plan + device descriptors)
void fh_table_tiles_free(fh_table_tiles *t);
bool launch_apply_table_tiles(cudaStream_t s, int sm, const fh_table *tab, const double2 *in, double2 *out, int mode,
                              double *d_partials, double *d_result);
void launch_pool(cudaStream_t s, const PoolEntry *entries, int first_entry, int n_entries, int chunks, int n,
                 const double2 *psi, const double2 *lam, double *d_partials, const int *entry_ids = nullptr,
                 int e0 = 0, int e1 = 0x7fffffff, int row = 0, int narrow = 0);   // narrow: every pattern pins <= 4 bits
// tiled scan of the entries covered by passes; only entries in [e0, e1) are evaluated
void launch_pool_tiles(cudaStream_t s, const PoolPass *d_passes, int npasses, const PoolTileRec *d_recs, int tile_bits,
                       int grid_x, int chunks, int n, const double2 *psi, const double2 *lam, double *d_partials, int e0,
                       int e1);
void launch_pool_finalize(cudaStream_t s, const double *d_partials, const int *d_out_first, int chunks, int first_out,
                          int count, double *d_out);
void launch_inner(cudaStream_t s, int sm, const double2 *a, const double2 *b, u64 dim, double *d_partials,
                  double *d_result);
void launch_sum_segments(cudaStream_t s, const double *d_partials, const int *d_first, int nseg, double *d_out);
void launch_sum_strided(cudaStream_t s, const double *d_partials, int stride, int nseg, double *d_out);
void launch_set_basis(cudaStream_t s, double2 *psi, u64 dim, u64 index);
void launch_flush(cudaStream_t s, void *buf, size_t bytes);
void launch_swap_bits(cudaStream_t s, int sm, const double2 *src, double2 *dst, int n, int npairs, const int *a,
                      const int *b);
// Lanczos helpers
void launch_axpby(cudaStream_t s, int sm, double2 *y, double a, const double2 *x, double b, u64 dim);          // y = a*x + b*y
void launch_lanczos_update(cudaStream_t s, int sm, double2 *w, const double2 *v, const double2 *vprev, double alpha,
                           double beta, u64 dim);                                                           // w -= alpha v + beta vprev
void launch_scale(cudaStream_t s, int sm, double2 *y, double a, u64 dim);
void launch_sector_random(cudaStream_t s, int sm, double2 *v, int n, int n_up, int n_dn, u64 seed);
void launch_caxpy(cudaStream_t s, int sm, double2 *y, double ar, double ai, const double2 *x, u64 dim);      // y += (ar + i ai) x

int fh_alloc_check(void *p, const char *what);
// opt every kernel that needs more than 48 KB of dynamic shared memory in on the CURRENT device (the attribute is
// per device; called from fh_ctx_create)
int fh_kernels_init_device();
