// Compiled circuits: ordered pair/diag ops, fused shared-memory tile runs, and the one-call
// evaluation (forward, observables, adjoint-gradient sweep, pool screening) replayed as a CUDA graph.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "common.cuh"
#include "sector_eval.cuh"

int fh_fill_pair(PairOp *op, int n, u64 x, u64 fixmask, u64 fixval, u64 zeta, const double m[8]);
int fh_enqueue_apply_table(const fh_table *tab, const double2 *in, double2 *out, int result_slot);
int fh_enqueue_pool(const fh_pool *pool, const double2 *psi, const double2 *lam, int first, int count,
                    double *d_out_override);

// tile_tma.cu: several consecutive tile runs in one cooperative launch
struct ChainRunHost {
    TileLaunch tl;
    int second_store;
    int init_basis;
};
int plan_tile_chain(int sm, double2 *psi, double2 *psi2, const ChainRunHost *hruns, int nruns, int n, void *h_runs,
                    void *h_maps, int map_base, int *grid_out, size_t *smem_out, int *rec_cap_out, int *term_cap_out,
                    int *tab_cap_out, int *tbits_out);
int launch_tile_chain(cudaStream_t s, const void *d_runs, int nruns, const void *d_maps, const TileRec *d_recs,
                      const TileTerm *d_terms, const unsigned short *d_ptab, int n, unsigned *d_sync, unsigned long long basis,
                      int grid, size_t smem, int rec_cap, int term_cap, int tab_cap, int tbits);
size_t fh_chain_run_bytes();
size_t fh_chain_map_bytes();

#define FH_MAX_RESULT_TABLES 8
#define FH_MAX_OVERLAPS 8

struct Item {
    int type;    // 1 pair, 2 diag, 3 tile
    int index;   // into pairs / diagops / tiles
};

struct EvalKey {
    u64 basis;
    int n_tables, n_overlaps, want_grads;
    const void *tables[FH_MAX_RESULT_TABLES];
    const void *targets[FH_MAX_OVERLAPS];
    const void *pool;
    int pool_pos, pool_first, pool_count;
    const void *state_out;
    int sector, sector_pool;     // 1: the sector-resident path (sector_eval.cu) / K3 on sector-compressed copies was planned
    int sector_k2, sector_dense; // 1: K2 of tables[0] on the sector-compressed state / the whole tail (W, H, W^dagger, K3) there
    int sector_prefix, pad_prefix;      // 1: the ops before the dense tail run in the cluster kernel (no full-space state at all)
    // unique ids of the same handles: a freed handle whose address is reused by a new one gets a new id, so the graph
    // (which bakes in the device pointers behind the handles) is re-captured instead of replayed on freed memory
    u64 table_uid[FH_MAX_RESULT_TABLES], target_uid[FH_MAX_OVERLAPS], pool_uid, state_out_uid;
    bool operator==(const EvalKey &o) const { return memcmp(this, &o, sizeof(EvalKey)) == 0; }
};

struct fh_program {
    fh_ctx *ctx = nullptr;
    int n = 0, n_params = 0;
    bool finalized = false, in_tile = false;
    std::vector<PairOp> pairs;
    std::vector<DiagOp> diagops;
    std::vector<DiagTerm> dterms;
    std::vector<TileOp> tiles;
    std::vector<TileSub> subs;
    std::vector<Item> items;
    // device mirrors + pinned staging of the theta-dependent payload
    PairOp *d_pairs = nullptr, *h_pairs = nullptr;
    DiagTerm *d_dterms = nullptr, *h_dterms = nullptr;
    DiagOp *d_diagops = nullptr;
    // fused-tile records, one array per direction (host-built, theta-dependent parts refreshed per call)
    std::vector<TileRec> recs_fwd, recs_dag;
    std::vector<TileTerm> tterms_fwd, tterms_dag;
    std::vector<int> pair_rec_fwd, pair_rec_dag;       // pair index -> record index (or -1)
    std::vector<int> dterm_tt_fwd, dterm_tt_dag;       // diag term index -> tile-term index (or -1)
    // static per-thread pair tables of the TMA tile kernels (theta-independent: uploaded once at finalize)
    std::vector<unsigned short> ptab_fwd, ptab_dag;
    unsigned short *d_ptab_fwd = nullptr, *d_ptab_dag = nullptr;
    TileRec *d_recs_fwd = nullptr, *d_recs_dag = nullptr, *h_recs_fwd = nullptr, *h_recs_dag = nullptr;
    TileTerm *d_tterms_fwd = nullptr, *d_tterms_dag = nullptr, *h_tterms_fwd = nullptr, *h_tterms_dag = nullptr;
    // the theta-dependent payload above lives in ONE pinned arena mirrored by ONE device arena (a single copy node
    // at the head of the captured graph); the typed pointers are views into them
    unsigned char *h_arena = nullptr, *d_arena = nullptr;
    size_t arena_bytes = 0;
    // chain descriptors (ChainRun records + tensor maps) behind the payload in the same arenas: region 0 belongs to the
    // captured evaluation graph (uploaded by the graph's first copy node), region 1 to immediate-mode runs
    size_t chain_off[2] = {0, 0}, chain_maps_off[2] = {0, 0};
    int chain_cap_runs = 0, chain_cap_maps = 0;
    int chain_used_runs[2] = {0, 0}, chain_used_maps[2] = {0, 0};
    int chain_region = 1;                 // region the next chain is planned into
    unsigned *d_sync = nullptr;           // [2 * chain_cap_runs] run-boundary counters of the chains
    // static launch descriptors of every tile run (forward order / reverse order of the dagger runs): the argument
    // arrays of the cooperative multi-run kernel
    TileLaunch *d_tl_fwd = nullptr, *d_tl_dag_rev = nullptr;
    // workspaces
    double2 *d_psi = nullptr, *d_lam = nullptr, *d_chk = nullptr;
    // adjoint-gradient partials: one segment per parametrised op processed
    double *d_gpart = nullptr, *d_gseg = nullptr, *h_gseg = nullptr;
    int *d_gfirst = nullptr;
    int n_param_ops = 0;
    // results (pinned)
    // results: [0, 64) scalars (expvals, overlaps) | [64, 64 + segs) gradient segments | pool outputs; ONE D2H copy
    double *h_res = nullptr, *d_res = nullptr;
    int res_segs = 0, res_pool_cap = 0;
    // graph cache
    bool have_graph = false;
    EvalKey key;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    std::vector<int> seg_param;       // per processed segment: parameter index
    std::vector<double> seg_scale;    // per processed segment: 2*gscale
    int n_segments = 0;
    // sector-resident evaluation (sector_eval.cu): ops in execution order, first flat op of every item, the plan
    std::vector<SecFlatOp> flat;
    std::vector<int> item_flat_first;
    fh_sector_plan *sec = nullptr;
    fh_sector_pool_plan *sec_pool = nullptr;
    fh_sector_dense *sec_dense = nullptr;
    fh_sector_plan *sec_prefix = nullptr;
    bool last_sector = false, last_sector_pool = false, last_sector_k2 = false, last_sector_dense = false, last_sector_prefix = false;
    // measurement
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double last_ms = 0.0;
    int last_launches = 0;
};

extern long long g_fh_launch_count;
extern thread_local int g_fh_tile_pdl_scope;      // kernels.cu: tile launches inside evaluate use PDL

static inline int popcnt(u64 v) { return __builtin_popcountll(v); }

// ----------------------------------------------------------------------------------------------
extern "C" int fh_program_create(fh_ctx *ctx, int n_qubits, int n_params, fh_program **out) {
    FH_REQUIRE(ctx && out, "fh_program_create: NULL argument");
    FH_REQUIRE(n_qubits >= 1 && n_qubits <= 33, "fh_program_create: n_qubits=%d outside [1, 33]", n_qubits);
    FH_REQUIRE(n_params >= 0, "fh_program_create: negative parameter count");
    fh_program *p = new (std::nothrow) fh_program();
    if (!p) return FH_ENOMEM;
    p->ctx = ctx;
    p->n = n_qubits;
    p->n_params = n_params;
    *out = p;
    return FH_OK;
}

static void drop_graph(fh_program *p) {
    if (p->exec) cudaGraphExecDestroy(p->exec);
    if (p->graph) cudaGraphDestroy(p->graph);
    p->exec = nullptr;
    p->graph = nullptr;
    p->have_graph = false;
}

extern "C" int fh_program_destroy(fh_program *p) {
    if (!p) return FH_OK;
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    drop_graph(p);
    fh_sector_plan_free(p->sec);
    p->sec = nullptr;
    fh_sector_pool_plan_free(p->sec_pool);
    p->sec_pool = nullptr;
    fh_sector_dense_free(p->sec_dense);
    p->sec_dense = nullptr;
    fh_sector_plan_free(p->sec_prefix);
    p->sec_prefix = nullptr;
    if (p->ev0) cudaEventDestroy(p->ev0);
    if (p->ev1) cudaEventDestroy(p->ev1);
    cudaFree(p->d_tl_fwd);
    cudaFree(p->d_tl_dag_rev);
    cudaFree(p->d_arena);
    cudaFreeHost(p->h_arena);
    cudaFree(p->d_diagops);
    // the three state-sized scratch vectors go back to the context (measured: cudaFree of 3 x 4 MiB = 11-15 ms,
    // i.e. more than 30 optimiser iterations of the 18-qubit driver, once per ADAPT epoch)
    const size_t state_bytes = sizeof(double2) << p->n;
    fh_ctx_scratch_put(p->ctx, state_bytes, p->d_psi);
    fh_ctx_scratch_put(p->ctx, state_bytes, p->d_lam);
    fh_ctx_scratch_put(p->ctx, state_bytes, p->d_chk);
    cudaFree(p->d_sync);
    cudaFree(p->d_ptab_fwd);
    cudaFree(p->d_ptab_dag);
    cudaFree(p->d_gpart);
    cudaFree(p->d_gfirst);
    cudaFree(p->d_res);
    cudaFreeHost(p->h_res);
    delete p;
    return FH_OK;
}

extern "C" int fh_program_add_pair(fh_program *p, uint64_t x, uint64_t fixmask, uint64_t fixval, uint64_t zeta, int kind,
                                   int param, double scale, double bhat_re, double bhat_im, const double m[8]) {
    FH_REQUIRE(p, "fh_program_add_pair: program is NULL");
    FH_REQUIRE(!p->finalized, "fh_program_add_pair: program already finalized");
    FH_REQUIRE(kind == 0 || kind == 1, "fh_program_add_pair: kind must be 0 (fixed) or 1 (rotation)");
    PairOp op;
    FH_TRY(fh_fill_pair(&op, p->n, x, fixmask, fixval, zeta, kind == 0 ? m : nullptr));
    op.kind = kind;
    if (kind == 0) {
        FH_REQUIRE(m != nullptr, "fh_program_add_pair: fixed op needs a matrix");
        op.param = -1;
    } else {
        FH_REQUIRE(param >= -1 && param < p->n_params, "fh_program_add_pair: parameter index %d out of range", param);
        const double nb = bhat_re * bhat_re + bhat_im * bhat_im;
        FH_REQUIRE(fabs(nb - 1.0) < 1e-9, "fh_program_add_pair: bhat must have unit modulus");
        op.param = param;
        op.gscale = scale;
        op.bhat[0] = bhat_re;
        op.bhat[1] = bhat_im;
    }
    if (p->in_tile)
        FH_REQUIRE(p->tiles.back().nsub < FH_TILE_MAX_SUB, "fh_program_add_pair: more than %d ops in one tile",
                   FH_TILE_MAX_SUB);
    const int idx = (int)p->pairs.size();
    p->pairs.push_back(op);
    if (p->in_tile) {
        TileOp &t = p->tiles.back();
        // x must lie inside the tile bits
        unsigned xlocal = 0;
        int lpivot = -1;
        u64 rest = x;
        for (int b = 0; b < t.nbits; ++b) {
            if (x >> t.bits[b] & 1) {
                xlocal |= 1u << b;
                lpivot = b;
                rest &= ~(1ull << t.bits[b]);
            }
        }
        FH_REQUIRE(rest == 0, "fh_program_add_pair: x-mask 0x%llx not inside the open tile", (u64)x);
        u64 tmask = 0;
        for (int b = 0; b < t.nbits; ++b) tmask |= 1ull << t.bits[b];
        FH_REQUIRE(popcnt(fixmask & tmask) <= 4, "fh_program_add_pair: more than 4 pattern bits inside a tile");
        TileSub s;
        s.type = 1;
        s.index = idx;
        s.lpivot = lpivot;
        s.xlocal = xlocal;
        p->subs.push_back(s);
        t.nsub++;
    } else {
        p->items.push_back({1, idx});
    }
    return FH_OK;
}

extern "C" int fh_program_add_diag(fh_program *p, int n_terms, const uint64_t *z, const double *coef, int param) {
    FH_REQUIRE(p, "fh_program_add_diag: program is NULL");
    FH_REQUIRE(!p->finalized, "fh_program_add_diag: program already finalized");
    FH_REQUIRE(n_terms > 0 && z && coef, "fh_program_add_diag: need at least one term");
    FH_REQUIRE(n_terms <= 512, "fh_program_add_diag: at most 512 terms per diagonal op (got %d)", n_terms);
    FH_REQUIRE(param >= -1 && param < p->n_params, "fh_program_add_diag: parameter index %d out of range", param);
    const u64 full = (1ull << p->n) - 1ull;
    DiagOp d;
    d.first = (int)p->dterms.size();
    d.count = n_terms;
    d.param = param;
    d.pad = 0;
    for (int m = 0; m < n_terms; ++m) {
        // z == 0 is a plain phase exp(-i angle): it arises when a sharded state folds the rank bits of a Z string away
        FH_REQUIRE((z[m] & ~full) == 0, "fh_program_add_diag: bad z-mask in term %d", m);
        DiagTerm t;
        t.z = z[m];
        t.coef = param >= 0 ? coef[m] : 0.0;
        t.angle = param >= 0 ? 0.0 : coef[m];
        t.c = cos(t.angle);
        t.s = sin(t.angle);
        p->dterms.push_back(t);
    }
    const int idx = (int)p->diagops.size();
    p->diagops.push_back(d);
    if (p->in_tile) {
        TileOp &t = p->tiles.back();
        FH_REQUIRE(t.nsub < FH_TILE_MAX_SUB, "fh_program_add_diag: more than %d ops in one tile", FH_TILE_MAX_SUB);
        FH_REQUIRE(t.pad + n_terms <= FH_TILE_MAX_TERMS, "fh_program_add_diag: more than %d diagonal terms in one tile",
                   FH_TILE_MAX_TERMS);
        TileSub s;
        s.type = 2;
        s.index = idx;
        s.lpivot = t.pad;          // tile-local offset of this op's terms
        s.xlocal = 0;
        p->subs.push_back(s);
        t.pad += n_terms;
        t.nsub++;
    } else {
        p->items.push_back({2, idx});
    }
    return FH_OK;
}

extern "C" int fh_program_begin_tile(fh_program *p, int n_bits, const int32_t *bits) {
    FH_REQUIRE(p && bits, "fh_program_begin_tile: NULL argument");
    FH_REQUIRE(!p->finalized && !p->in_tile, "fh_program_begin_tile: program finalized or tile already open");
    FH_REQUIRE(n_bits >= 1 && n_bits <= FH_MAX_TILE_BITS && n_bits <= p->n, "fh_program_begin_tile: n_bits=%d outside [1, %d]",
               n_bits, FH_MAX_TILE_BITS);
    FH_REQUIRE(p->n <= 32, "fh_program_begin_tile: tiles support at most 32 qubits");
    TileOp t;
    memset(&t, 0, sizeof(t));
    t.nbits = n_bits;
    t.first_sub = (int)p->subs.size();
    t.nsub = 0;
    for (int b = 0; b < n_bits; ++b) {
        FH_REQUIRE(bits[b] >= 0 && bits[b] < p->n && (b == 0 || bits[b] > bits[b - 1]),
                   "fh_program_begin_tile: bits must be ascending and < n_qubits");
        t.bits[b] = (unsigned char)bits[b];
    }
    p->tiles.push_back(t);
    p->in_tile = true;
    return FH_OK;
}

extern "C" int fh_program_end_tile(fh_program *p) {
    FH_REQUIRE(p && p->in_tile, "fh_program_end_tile: no open tile");
    p->in_tile = false;
    if (p->tiles.back().nsub == 0) {
        p->tiles.pop_back();
        return FH_OK;
    }
    p->items.push_back({3, (int)p->tiles.size() - 1});
    return FH_OK;
}

// ----------------------------------------------------------------------------------------------
// fused-tile records (host): everything k_tile needs per op, in execution order, for both directions
// ----------------------------------------------------------------------------------------------
static void set_rec_matrix(TileRec &r, const double m[8], bool dagger) {
    if (!dagger) {
        for (int k = 0; k < 8; ++k) r.m[k] = m[k];
    } else {   // conjugate transpose
        r.m[0] = m[0]; r.m[1] = -m[1];
        r.m[2] = m[4]; r.m[3] = -m[5];
        r.m[4] = m[2]; r.m[5] = -m[3];
        r.m[6] = m[6]; r.m[7] = -m[7];
    }
    const bool real = (r.m[1] == 0.0 && r.m[3] == 0.0 && r.m[5] == 0.0 && r.m[7] == 0.0);
    const bool rdiag = (r.m[1] == 0.0 && r.m[7] == 0.0);       // every rotation exp(-i a G) has a real diagonal
    r.type = real ? 3 : (rdiag ? 6 : 1);
}

static void set_tile_term(TileTerm &tt, const DiagTerm &t, bool dagger) {
    tt.z = t.z;
    tt.angle = dagger ? -t.angle : t.angle;
    tt.c = t.c;
    tt.s = dagger ? -t.s : t.s;
    tt.coef = t.coef;
    tt.zlocal = 0;
    tt.pad = 0;
}

static void build_tile_records(fh_program *p) {
    p->pair_rec_fwd.assign(p->pairs.size(), -1);
    p->pair_rec_dag.assign(p->pairs.size(), -1);
    p->dterm_tt_fwd.assign(p->dterms.size(), -1);
    p->dterm_tt_dag.assign(p->dterms.size(), -1);
    for (auto &t : p->tiles) {
        unsigned tilemask = 0;
        for (int b = 0; b < t.nbits; ++b) tilemask |= 1u << t.bits[b];
        const int layout = fh_tile_tma_layout(t.bits, t.nbits, p->n), threads = fh_tile_threads(t.nbits);
        for (int dir = 0; dir < 2; ++dir) {
            std::vector<TileRec> &recs = dir ? p->recs_dag : p->recs_fwd;
            std::vector<TileTerm> &tts = dir ? p->tterms_dag : p->tterms_fwd;
            std::vector<unsigned short> &ptab = dir ? p->ptab_dag : p->ptab_fwd;
            const size_t tab_start = ptab.size();
            (dir ? t.first_ptab_dag : t.first_ptab_fwd) = (int)tab_start;
            (dir ? t.first_rec_dag : t.first_rec_fwd) = (int)recs.size();
            (dir ? t.first_term_dag : t.first_term_fwd) = (int)tts.size();
            int term_off = 0, seg = 0;
            for (int k = 0; k < t.nsub; ++k) {
                const TileSub &sub = p->subs[t.first_sub + (dir ? t.nsub - 1 - k : k)];
                TileRec r;
                memset(&r, 0, sizeof(r));
                r.seg = -1;
                if (sub.type == 1) {
                    const PairOp &op = p->pairs[sub.index];
                    set_rec_matrix(r, op.m, dir != 0);
                    r.bhat[0] = op.bhat[0];
                    r.bhat[1] = op.bhat[1];
                    if (op.kind == 1 && op.param >= 0) r.seg = seg++;
                    const unsigned fm = (unsigned)op.fixmask, fv = (unsigned)op.fixval;
                    r.fixmask_out = fm & ~tilemask;
                    r.fixval_out = fv & ~tilemask;
                    r.zeta = (unsigned)op.zeta;
                    r.xlocal = sub.xlocal;
                    int nl = 0;
                    unsigned lv = 0;
                    for (int q = 0; q < 4; ++q) r.lowmask[q] = 0xffffffffu;
                    for (int b = 0; b < t.nbits; ++b)
                        if (fm >> t.bits[b] & 1u) {
                            if (nl < 4) r.lowmask[nl] = (1u << b) - 1u;
                            ++nl;
                            lv |= ((fv >> t.bits[b]) & 1u) << b;
                        }
                    r.nlfix = nl;
                    r.lfixval = lv;
                    for (int b = 0; b < t.nbits; ++b)
                        if (op.zeta >> t.bits[b] & 1ull) r.zeta_local |= 1u << b;
                    (dir ? p->pair_rec_dag : p->pair_rec_fwd)[sub.index] = (int)recs.size();
                    if (layout >= 0) {
                        // per-thread pair table: 16-bit entry [rep * threads + tid] = slot of the pattern side | in-tile sign
                        // parity << 13 | valid << 14 for pair number rep * threads + tid of this op; partner slot = slot ^ xs
                        auto slot = [&](unsigned l) { return layout ? (l ^ ((l >> 3) & 7u)) : l; };
                        const unsigned npairs = (1u << t.nbits) >> nl;
                        const unsigned reps = (npairs + (unsigned)threads - 1u) / (unsigned)threads;
                        r.term_off = (int)(ptab.size() - tab_start);
                        r.reps = (int)reps;
                        r.run_pad[0] = (npairs == reps * (unsigned)threads) ? 1 : 0;      // table without invalid entries
                        r.nterms = (int)slot(sub.xlocal);
                        for (unsigned k = 0; k < reps * (unsigned)threads; ++k) {
                            unsigned short word = 0;
                            if (k < npairs) {
                                unsigned il = k;
                                for (int q = 0; q < 4; ++q) il = ((il & ~r.lowmask[q]) << 1) | (il & r.lowmask[q]);
                                il |= lv;
                                const unsigned par = (unsigned)__builtin_popcount(il & r.zeta_local) & 1u;
                                word = (unsigned short)(slot(il) | (par << 13) | (1u << 14));
                            }
                            ptab.push_back(word);
                        }
                    }
                } else {
                    const DiagOp &d = p->diagops[sub.index];
                    r.type = 2;
                    r.term_off = term_off;
                    r.nterms = d.count;
                    if (d.param >= 0) r.seg = seg++;
                    // terms whose in-tile z bits straddle the two halves of the tile (local bits 0..5 | 6..) first: the TMA
                    // kernel evaluates those per amplitude and folds all the others into two phase tables
                    unsigned lomask = 0, himask = 0;
                    for (int b = 0; b < t.nbits; ++b) (b < 6 ? lomask : himask) |= 1u << t.bits[b];
                    int nstr = 0;
                    for (int pass = 0; pass < 2; ++pass)
                        for (int m = 0; m < d.count; ++m) {
                            const u64 z = p->dterms[d.first + m].z;
                            const bool straddles = (z & lomask) != 0 && (z & himask) != 0;
                            if (straddles != (pass == 0)) continue;
                            TileTerm tt;
                            set_tile_term(tt, p->dterms[d.first + m], dir != 0);
                            for (int b = 0; b < t.nbits; ++b)
                                if (tt.z >> t.bits[b] & 1ull) tt.zlocal |= 1u << b;
                            (dir ? p->dterm_tt_dag : p->dterm_tt_fwd)[d.first + m] = (int)tts.size();
                            tts.push_back(tt);
                            nstr += straddles;
                        }
                    r.reps = nstr;
                    term_off += d.count;
                }
                recs.push_back(r);
            }
            t.nterms = term_off;
            t.n_param_subs = seg;
            t.ptab_words = (int)(ptab.size() - tab_start);
            // register-fused runs (TMA kernels): consecutive Givens-like ops (x = two in-tile bits, pattern pins exactly
            // those two, nothing outside the tile) whose bits stay inside one set of three tile-local bits
            // Measured (profiles/r02_tile_ab.md): a fused run costs as much as its ops one by one (the op loop is bound by the
            // instruction stream of each warp, not by the shared-memory round trips), so runs are opt-in: FHSIM_RUNS=1
            if (layout >= 0 && getenv("FHSIM_RUNS")) {
                const size_t r0 = recs.size() - (size_t)t.nsub;
                auto givens_like = [&](const TileRec &r, unsigned &bits2) {
                    if (r.type == 2 || r.nlfix != 2 || r.fixmask_out != 0u || __builtin_popcount(r.xlocal) != 2) return false;
                    unsigned pinned = 0;
                    for (int q = 0; q < 2; ++q) pinned |= r.lowmask[q] + 1u;          // lowmask = (1 << b) - 1
                    if (pinned != r.xlocal || __builtin_popcount(r.lfixval & r.xlocal) != 1) return false;     // pattern 01 <-> 10
                    bits2 = r.xlocal;
                    return true;
                };
                size_t i = 0;
                while (i < (size_t)t.nsub) {
                    unsigned b2 = 0;
                    if (!givens_like(recs[r0 + i], b2)) {
                        ++i;
                        continue;
                    }
                    unsigned B = b2;
                    size_t j = i + 1;
                    while (j < (size_t)t.nsub && j - i < 8) {
                        unsigned c2 = 0;
                        if (!givens_like(recs[r0 + j], c2) || __builtin_popcount(B | c2) > 3) break;
                        B |= c2;
                        ++j;
                    }
                    if (j - i >= 2) {
                        for (int b = 0; b < t.nbits && __builtin_popcount(B) < 3; ++b) B |= 1u << b;     // pad to 3 bits
                        int pos[3], np = 0;
                        for (int b = 0; b < t.nbits; ++b)
                            if (B >> b & 1u) pos[np++] = b;
                        for (size_t k = i; k < j; ++k) {
                            TileRec &r = recs[r0 + k];
                            r.run_len = k == i ? (int)(j - i) : 0;
                            r.run_bits = (unsigned)pos[0] | ((unsigned)pos[1] << 8) | ((unsigned)pos[2] << 16);
                            unsigned si = 0, x3 = 0;
                            for (int q = 0; q < 3; ++q) {
                                if (r.lfixval >> pos[q] & 1u) si |= 1u << q;
                                if (r.xlocal >> pos[q] & 1u) x3 |= 1u << q;
                            }
                            r.sub_i = (unsigned char)si;
                            r.x3 = (unsigned char)x3;
                        }
                    }
                    i = j;
                }
            }
        }
    }
}

static TileLaunch make_tile_launch(const TileOp &t, int dagger) {
    TileLaunch tl;
    memset(&tl, 0, sizeof(tl));
    tl.nbits = t.nbits;
    tl.nsub = t.nsub;
    tl.first_rec = dagger ? t.first_rec_dag : t.first_rec_fwd;
    tl.first_term = dagger ? t.first_term_dag : t.first_term_fwd;
    tl.nterms = t.nterms;
    tl.ptab_first = dagger ? t.first_ptab_dag : t.first_ptab_fwd;
    tl.ptab_words = t.ptab_words;
    memcpy(tl.bits, t.bits, 16);
    return tl;
}

template <typename T>
static int upload_vec(T **dptr, const std::vector<T> &v, cudaStream_t s) {
    if (v.empty()) return FH_OK;
    FH_CUDA(cudaMalloc(dptr, sizeof(T) * v.size()));
    FH_CUDA(cudaMemcpyAsync(*dptr, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice, s));
    return FH_OK;
}

extern "C" int fh_program_finalize(fh_program *p) {
    FH_REQUIRE(p, "fh_program_finalize: program is NULL");
    FH_REQUIRE(!p->finalized && !p->in_tile, "fh_program_finalize: already finalized or tile still open");
    fh_ctx *ctx = p->ctx;
    FH_CUDA(cudaSetDevice(ctx->device));
    FH_TRY(upload_vec(&p->d_diagops, p->diagops, ctx->stream));
    build_tile_records(p);
    {
        // carve the payload arena: [pairs | dterms | recs_fwd | recs_dag | tterms_fwd | tterms_dag], 256-byte aligned
        auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
        const size_t sz[6] = {sizeof(PairOp) * p->pairs.size(),       sizeof(DiagTerm) * p->dterms.size(),
                              sizeof(TileRec) * p->recs_fwd.size(),   sizeof(TileRec) * p->recs_dag.size(),
                              sizeof(TileTerm) * p->tterms_fwd.size(), sizeof(TileTerm) * p->tterms_dag.size()};
        size_t off[6], total = 0;
        for (int k = 0; k < 6; ++k) {
            off[k] = total;
            total += align(sz[k]);
        }
        p->arena_bytes = total;           // theta-dependent payload; the chain regions follow
        p->chain_cap_runs = 2 * (int)p->tiles.size() + 2;
        p->chain_cap_maps = p->chain_cap_runs + 4;
        if (!p->tiles.empty()) {
            for (int reg = 0; reg < 2; ++reg) {
                p->chain_off[reg] = total;
                total += align(fh_chain_run_bytes() * (size_t)p->chain_cap_runs);
                p->chain_maps_off[reg] = total;
                total += align(fh_chain_map_bytes() * (size_t)p->chain_cap_maps);
            }
            FH_CUDA(cudaMalloc(&p->d_sync, sizeof(unsigned) * 2 * (size_t)p->chain_cap_runs));
            FH_CUDA(cudaMemsetAsync(p->d_sync, 0, sizeof(unsigned) * 2 * (size_t)p->chain_cap_runs, ctx->stream));
        }
        if (total) {
            FH_CUDA(cudaMalloc(&p->d_arena, total));
            FH_CUDA(cudaMallocHost(&p->h_arena, total));
            memset(p->h_arena, 0, total);
        }
        p->h_pairs = reinterpret_cast<PairOp *>(p->h_arena + off[0]);
        p->d_pairs = reinterpret_cast<PairOp *>(p->d_arena + off[0]);
        p->h_dterms = reinterpret_cast<DiagTerm *>(p->h_arena + off[1]);
        p->d_dterms = reinterpret_cast<DiagTerm *>(p->d_arena + off[1]);
        p->h_recs_fwd = reinterpret_cast<TileRec *>(p->h_arena + off[2]);
        p->d_recs_fwd = reinterpret_cast<TileRec *>(p->d_arena + off[2]);
        p->h_recs_dag = reinterpret_cast<TileRec *>(p->h_arena + off[3]);
        p->d_recs_dag = reinterpret_cast<TileRec *>(p->d_arena + off[3]);
        p->h_tterms_fwd = reinterpret_cast<TileTerm *>(p->h_arena + off[4]);
        p->d_tterms_fwd = reinterpret_cast<TileTerm *>(p->d_arena + off[4]);
        p->h_tterms_dag = reinterpret_cast<TileTerm *>(p->h_arena + off[5]);
        p->d_tterms_dag = reinterpret_cast<TileTerm *>(p->d_arena + off[5]);
        if (sz[0]) memcpy(p->h_pairs, p->pairs.data(), sz[0]);
        if (sz[1]) memcpy(p->h_dterms, p->dterms.data(), sz[1]);
        if (sz[2]) memcpy(p->h_recs_fwd, p->recs_fwd.data(), sz[2]);
        if (sz[3]) memcpy(p->h_recs_dag, p->recs_dag.data(), sz[3]);
        if (sz[4]) memcpy(p->h_tterms_fwd, p->tterms_fwd.data(), sz[4]);
        if (sz[5]) memcpy(p->h_tterms_dag, p->tterms_dag.data(), sz[5]);
        if (p->arena_bytes)
            FH_CUDA(cudaMemcpyAsync(p->d_arena, p->h_arena, p->arena_bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    FH_TRY(upload_vec(&p->d_ptab_fwd, p->ptab_fwd, ctx->stream));
    FH_TRY(upload_vec(&p->d_ptab_dag, p->ptab_dag, ctx->stream));
    if (!p->tiles.empty()) {
        const size_t nt = p->tiles.size();
        std::vector<TileLaunch> fwd(nt), dag_rev(nt);
        for (size_t k = 0; k < nt; ++k) {
            fwd[k] = make_tile_launch(p->tiles[k], 0);
            dag_rev[nt - 1 - k] = make_tile_launch(p->tiles[k], 1);
        }
        FH_TRY(upload_vec(&p->d_tl_fwd, fwd, ctx->stream));
        FH_TRY(upload_vec(&p->d_tl_dag_rev, dag_rev, ctx->stream));
    }
    int n_param_ops = 0;
    for (auto &op : p->pairs) n_param_ops += (op.kind == 1 && op.param >= 0);
    for (auto &d : p->diagops) n_param_ops += (d.param >= 0);
    p->n_param_ops = n_param_ops;
    const int segs = n_param_ops > 0 ? n_param_ops : 1;
    FH_CUDA(cudaMalloc(&p->d_gpart, sizeof(double) * (size_t)segs * FH_GRAD_BLOCKS));
    FH_CUDA(cudaMalloc(&p->d_gfirst, sizeof(int) * (segs + 1)));
    p->res_segs = segs;
    p->res_pool_cap = 0;
    FH_CUDA(cudaMalloc(&p->d_res, sizeof(double) * (64 + segs)));
    FH_CUDA(cudaMallocHost(&p->h_res, sizeof(double) * (64 + segs)));
    p->d_gseg = p->d_res + 64;
    p->h_gseg = p->h_res + 64;
    FH_CUDA(cudaStreamSynchronize(ctx->stream));
    // ops in execution order (tiles expanded) for the sector-resident path
    for (const Item &it : p->items) {
        p->item_flat_first.push_back((int)p->flat.size());
        if (it.type == 3) {
            const TileOp &t = p->tiles[it.index];
            for (int s = t.first_sub; s < t.first_sub + t.nsub; ++s) p->flat.push_back({p->subs[s].type, p->subs[s].index});
        } else {
            p->flat.push_back({it.type, it.index});
        }
    }
    p->item_flat_first.push_back((int)p->flat.size());
    p->finalized = true;
    return FH_OK;
}

extern "C" int fh_program_info(const fh_program *p, int *n_ops, int *n_launches) {
    FH_REQUIRE(p, "fh_program_info: program is NULL");
    if (n_ops) *n_ops = (int)(p->pairs.size() + p->diagops.size());
    if (n_launches) *n_launches = (int)p->items.size();
    return FH_OK;
}

// ----------------------------------------------------------------------------------------------
// theta -> payload (host, double precision), staged in pinned memory and uploaded on the stream
// ----------------------------------------------------------------------------------------------
static void refresh_payload(fh_program *p, const double *thetas) {
    for (size_t k = 0; k < p->pairs.size(); ++k) {
        PairOp &op = p->h_pairs[k];
        if (op.kind != 1) continue;
        const double a = op.param >= 0 ? op.gscale * thetas[op.param] : op.gscale;
        const double c = cos(a), s = sin(a), br = op.bhat[0], bi = op.bhat[1];
        op.m[0] = c;       op.m[1] = 0.0;
        op.m[2] = s * bi;  op.m[3] = -s * br;     // -i s bhat
        op.m[4] = -s * bi; op.m[5] = -s * br;     // -i s conj(bhat)
        op.m[6] = c;       op.m[7] = 0.0;
        if (p->pair_rec_fwd[k] >= 0) {
            set_rec_matrix(p->h_recs_fwd[p->pair_rec_fwd[k]], op.m, false);
            set_rec_matrix(p->h_recs_dag[p->pair_rec_dag[k]], op.m, true);
        }
    }
    for (auto &d : p->diagops) {
        if (d.param < 0) continue;
        for (int m = d.first; m < d.first + d.count; ++m) {
            DiagTerm &t = p->h_dterms[m];
            t.angle = thetas[d.param] * t.coef;
            t.c = cos(t.angle);
            t.s = sin(t.angle);
            if (p->dterm_tt_fwd[m] >= 0) {
                TileTerm &f = p->h_tterms_fwd[p->dterm_tt_fwd[m]], &g = p->h_tterms_dag[p->dterm_tt_dag[m]];
                const unsigned zl = f.zlocal;         // geometry: fixed at finalize
                set_tile_term(f, t, false);
                set_tile_term(g, t, true);
                f.zlocal = g.zlocal = zl;
            }
        }
    }
}

// copy the pinned payload to the device (these become the first nodes of the captured graph)
static int enqueue_payload_upload(fh_program *p, bool with_graph_chains = false) {
    size_t bytes = p->arena_bytes;
    if (with_graph_chains && p->chain_cap_runs > 0 && !p->tiles.empty()) {
        // region 0 sits right behind the payload: one copy node moves both.  The host side of the region is filled
        // while the evaluation is being captured (nothing executes before the graph is launched).
        bytes = p->chain_maps_off[0] + fh_chain_map_bytes() * (size_t)p->chain_cap_maps;
        FH_CUDA(cudaMemsetAsync(p->d_sync, 0, sizeof(unsigned) * (size_t)p->chain_cap_runs, p->ctx->stream));
    }
    if (bytes) FH_CUDA(cudaMemcpyAsync(p->d_arena, p->h_arena, bytes, cudaMemcpyHostToDevice, p->ctx->stream));
    return FH_OK;
}

static int upload_payload(fh_program *p, const double *thetas, int n_thetas) {
    FH_REQUIRE(n_thetas == p->n_params, "program expects %d parameters, got %d", p->n_params, n_thetas);
    FH_REQUIRE(n_thetas == 0 || thetas, "thetas is NULL");
    refresh_payload(p, thetas);
    return enqueue_payload_upload(p);
}

static bool item_has_param(const fh_program *p, const Item &it) {
    if (it.type == 1) return p->pairs[it.index].kind == 1 && p->pairs[it.index].param >= 0;
    if (it.type == 2) return p->diagops[it.index].param >= 0;
    const TileOp &t = p->tiles[it.index];
    for (int s = t.first_sub; s < t.first_sub + t.nsub; ++s) {
        const TileSub &sub = p->subs[s];
        if (sub.type == 1 && p->pairs[sub.index].kind == 1 && p->pairs[sub.index].param >= 0) return true;
        if (sub.type == 2 && p->diagops[sub.index].param >= 0) return true;
    }
    return false;
}

static void apply_item(fh_program *p, const Item &it, double2 *st, int dagger) {
    fh_ctx *ctx = p->ctx;
    if (it.type == 1) {
        launch_pair(ctx->stream, ctx->sm_count, st, p->d_pairs + it.index, p->n, p->pairs[it.index].npos, dagger);
    } else if (it.type == 2) {
        const DiagOp &d = p->diagops[it.index];
        launch_diag(ctx->stream, ctx->sm_count, st, p->d_dterms + d.first, d.count, p->n, dagger, ctx->d_diag);
    } else {
        const TileOp &t = p->tiles[it.index];
        const TileLaunch tl = make_tile_launch(t, dagger);
        launch_tile(ctx->stream, st, tl, dagger ? p->d_recs_dag : p->d_recs_fwd,
                    dagger ? p->d_tterms_dag : p->d_tterms_fwd, dagger ? p->d_ptab_dag : p->d_ptab_fwd, p->n);
    }
}

// One cooperative launch for the consecutive tile items [k0, k0 + run) (positions in travel order) of the range, or 0.
static int try_chain(fh_program *p, int lo, int hi, int k0, int run, double2 *st, int dagger, double2 *chk, int chk_pos,
                     bool init_basis, u64 basis) {
    fh_ctx *ctx = p->ctx;
    const int reg = p->chain_region;
    if (p->tiles.empty() || p->chain_used_runs[reg] + run > p->chain_cap_runs) return 0;
    std::vector<ChainRunHost> hr((size_t)run);
    int n_second = 0;
    for (int r = 0; r < run; ++r) {
        const int idx = dagger ? hi - 1 - (k0 + r) : lo + k0 + r;
        const TileOp &t = p->tiles[p->items[idx].index];
        hr[r].tl = make_tile_launch(t, dagger);
        hr[r].second_store = (!dagger && chk && idx + 1 == chk_pos) ? 1 : 0;
        hr[r].init_basis = (init_basis && r == 0) ? 1 : 0;
        n_second += hr[r].second_store;
    }
    if (p->chain_used_maps[reg] + run + n_second > p->chain_cap_maps) return 0;
    unsigned char *h_runs = p->h_arena + p->chain_off[reg] + fh_chain_run_bytes() * (size_t)p->chain_used_runs[reg];
    unsigned char *h_maps = p->h_arena + p->chain_maps_off[reg];
    int grid = 0, rec_cap = 0, term_cap = 0, tab_cap = 0, tbits = 0;
    size_t smem = 0;
    const int nmaps = plan_tile_chain(ctx->sm_count, st, chk, hr.data(), run, p->n, h_runs,
                                      h_maps + fh_chain_map_bytes() * (size_t)p->chain_used_maps[reg], p->chain_used_maps[reg],
                                      &grid, &smem, &rec_cap, &term_cap, &tab_cap, &tbits);
    if (nmaps < 0) return 0;
    const unsigned char *d_runs = p->d_arena + p->chain_off[reg] + fh_chain_run_bytes() * (size_t)p->chain_used_runs[reg];
    const unsigned char *d_maps = p->d_arena + p->chain_maps_off[reg];
    unsigned *d_sync = p->d_sync + (size_t)reg * p->chain_cap_runs + p->chain_used_runs[reg];
    if (reg == 1) {
        // immediate mode: descriptors and counters travel right before the launch
        cudaMemcpyAsync(const_cast<unsigned char *>(d_runs), h_runs, fh_chain_run_bytes() * (size_t)run, cudaMemcpyHostToDevice,
                        ctx->stream);
        const size_t mo = fh_chain_map_bytes() * (size_t)p->chain_used_maps[reg];
        cudaMemcpyAsync(const_cast<unsigned char *>(d_maps) + mo, h_maps + mo, fh_chain_map_bytes() * (size_t)nmaps,
                        cudaMemcpyHostToDevice, ctx->stream);
        cudaMemsetAsync(d_sync, 0, sizeof(unsigned) * (size_t)run, ctx->stream);
    }
    if (!launch_tile_chain(ctx->stream, d_runs, run, d_maps, dagger ? p->d_recs_dag : p->d_recs_fwd,
                           dagger ? p->d_tterms_dag : p->d_tterms_fwd, dagger ? p->d_ptab_dag : p->d_ptab_fwd, p->n, d_sync,
                           basis, grid, smem, rec_cap, term_cap, tab_cap, tbits))
        return 0;
    p->chain_used_runs[reg] += run;
    p->chain_used_maps[reg] += nmaps;
    return 1;
}

// Apply items [lo, hi) in execution order of the direction.  Consecutive tile runs go into ONE cooperative launch
// (k_tile_chain) when every tile of every run can be resident at once (18-20 qubits); otherwise one launch per item.
// Optional extras of the forward direction: `basis` (>= 0: the state is |basis> before the first item; a chain that starts
// the range synthesises it, else a set-basis launch) and `chk` (copy of the state after item chk_pos - 1, fused into
// the chain as a second tile store, else a device-to-device copy).
static int apply_range(fh_program *p, int lo, int hi, double2 *st, int dagger, double2 *chk = nullptr, int chk_pos = -1,
                       long long basis = -1) {
    fh_ctx *ctx = p->ctx;
    const size_t bytes = sizeof(double2) << p->n;
    const int count = hi - lo;
    bool basis_pending = basis >= 0;
    auto settle_basis = [&]() {
        if (basis_pending) launch_set_basis(ctx->stream, st, 1ull << p->n, (u64)basis);
        basis_pending = false;
    };
    if (chk && chk_pos == lo && !dagger) {
        settle_basis();
        FH_CUDA(cudaMemcpyAsync(chk, st, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    int k = 0;
    while (k < count) {
        const int idx = dagger ? hi - 1 - k : lo + k;
        const Item &it = p->items[idx];
        if (it.type == 3) {
            int run = 1;
            while (k + run < count && p->items[dagger ? hi - 1 - (k + run) : lo + k + run].type == 3) ++run;
            bool has_chk = false;
            for (int r = 0; r < run; ++r) {
                const int j = dagger ? hi - 1 - (k + r) : lo + k + r;
                has_chk = has_chk || (!dagger && chk && j + 1 == chk_pos);
            }
            const bool use_basis = basis_pending && k == 0;
            if ((run >= 2 || has_chk || use_basis) &&
                try_chain(p, lo, hi, k, run, st, dagger, chk, chk_pos, use_basis, basis >= 0 ? (u64)basis : 0ull)) {
                if (use_basis) basis_pending = false;
                k += run;
                continue;
            }
        }
        settle_basis();
        apply_item(p, it, st, dagger);
        if (!dagger && chk && idx + 1 == chk_pos && chk_pos != lo)
            FH_CUDA(cudaMemcpyAsync(chk, st, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        ++k;
    }
    settle_basis();
    return FH_OK;
}

extern "C" int fh_program_run(fh_program *p, fh_state *st, const double *thetas, int n_thetas, int first, int count,
                              int dagger) {
    FH_REQUIRE(p && st, "fh_program_run: NULL argument");
    FH_REQUIRE(p->finalized, "fh_program_run: program not finalized");
    FH_REQUIRE(st->n == p->n, "fh_program_run: state has %d qubits, program %d", st->n, p->n);
    FH_REQUIRE(first >= 0 && count >= 0 && first + count <= (int)p->items.size(), "fh_program_run: op range out of bounds");
    FH_TRY(upload_payload(p, thetas, n_thetas));
    p->chain_region = 1;
    p->chain_used_runs[1] = p->chain_used_maps[1] = 0;
    FH_TRY(apply_range(p, first, first + count, st->d, dagger));
    FH_CUDA(cudaGetLastError());
    FH_CUDA(cudaStreamSynchronize(p->ctx->stream));   // pinned payload may be rewritten by the next call
    return FH_OK;
}

// ----------------------------------------------------------------------------------------------
// adjoint processing of one logical op (gradient partial + undo on psi and lam)
// ----------------------------------------------------------------------------------------------
static void adjoint_logical(fh_program *p, int type, int index, double2 *psi, double2 *lam, bool want_grads) {
    fh_ctx *ctx = p->ctx;
    int used = 0;
    if (type == 1) {
        const PairOp &op = p->pairs[index];
        if (want_grads && op.kind == 1 && op.param >= 0) {
            launch_pair_adjoint(ctx->stream, ctx->sm_count, psi, lam, p->d_pairs + index, p->n, op.npos,
                                p->d_gpart + (size_t)p->n_segments * FH_GRAD_BLOCKS, FH_GRAD_BLOCKS, &used);
            p->seg_param.push_back(op.param);
            p->seg_scale.push_back(2.0 * op.gscale);
            p->n_segments++;
        } else {
            launch_pair(ctx->stream, ctx->sm_count, psi, p->d_pairs + index, p->n, op.npos, 1);
            launch_pair(ctx->stream, ctx->sm_count, lam, p->d_pairs + index, p->n, op.npos, 1);
        }
    } else {
        const DiagOp &d = p->diagops[index];
        if (want_grads && d.param >= 0) {
            launch_diag_adjoint(ctx->stream, ctx->sm_count, psi, lam, p->d_dterms + d.first, d.count, p->n,
                                p->d_gpart + (size_t)p->n_segments * FH_GRAD_BLOCKS, FH_GRAD_BLOCKS, &used);
            p->seg_param.push_back(d.param);
            p->seg_scale.push_back(2.0);
            p->n_segments++;
        } else {
            launch_diag(ctx->stream, ctx->sm_count, psi, p->d_dterms + d.first, d.count, p->n, 1, ctx->d_diag);
            launch_diag(ctx->stream, ctx->sm_count, lam, p->d_dterms + d.first, d.count, p->n, 1, ctx->d_diag);
        }
    }
}

static void adjoint_item(fh_program *p, const Item &it, double2 *psi, double2 *lam, bool want_grads) {
    if (it.type != 3) {
        adjoint_logical(p, it.type, it.index, psi, lam, want_grads);
        return;
    }
    if (!want_grads || !item_has_param(p, it)) {
        apply_item(p, it, psi, 1);
        apply_item(p, it, lam, 1);
        return;
    }
    const TileOp &t = p->tiles[it.index];
    static const bool unfused = getenv("FHSIM_UNFUSED_ADJOINT") != nullptr;
    if (!unfused && t.nbits <= FH_TILE_ADJOINT_MAX_BITS) {
        // one launch: gradient partials of every parametrised op of the run + the inverse run on psi and lam
        const TileLaunch tl = make_tile_launch(t, 1);
        launch_tile_adjoint(p->ctx->stream, psi, lam, tl, p->d_recs_dag, p->d_tterms_dag, p->n, p->d_gpart, p->n_segments);
        for (int s = t.first_sub + t.nsub - 1; s >= t.first_sub; --s) {
            const TileSub &sub = p->subs[s];
            if (sub.type == 1) {
                const PairOp &op = p->pairs[sub.index];
                if (op.kind == 1 && op.param >= 0) {
                    p->seg_param.push_back(op.param);
                    p->seg_scale.push_back(2.0 * op.gscale);
                    p->n_segments++;
                }
            } else if (p->diagops[sub.index].param >= 0) {
                p->seg_param.push_back(p->diagops[sub.index].param);
                p->seg_scale.push_back(2.0);
                p->n_segments++;
            }
        }
        return;
    }
    for (int s = t.first_sub + t.nsub - 1; s >= t.first_sub; --s)
        adjoint_logical(p, p->subs[s].type, p->subs[s].index, psi, lam, want_grads);
}

// Enqueue one whole evaluation on the stream (this is what gets captured into the CUDA graph).
// Result buffer h_res (doubles): [0, 2T) expvals re/im, [2T, 2T+2V) overlaps re/im.
static int enqueue_evaluation(fh_program *p, const EvalKey &k, fh_table *const *tables, const fh_pool *pool,
                              fh_state *const *targets, fh_state *state_out, bool capturing) {
    fh_ctx *ctx = p->ctx;
    const int n_items = (int)p->items.size();
    const size_t bytes = sizeof(double2) << p->n;
    const bool want_grads = k.want_grads != 0;
    const bool want_pool = pool != nullptr;
    const bool need_adjoint = want_grads || want_pool;
    if (k.sector) {
        // the whole evaluation on the sector-compressed state inside one cluster (sector_eval.cu): two launches
        double *d_pool_out_s = p->d_res + 64 + p->res_segs;
        FH_TRY(enqueue_payload_upload(p, false));
        FH_TRY(fh_sector_enqueue(p->sec, ctx, k.basis, p->d_pairs, p->d_dterms, p->d_res, pool, k.pool_first, k.pool_count,
                                 d_pool_out_s));
        const size_t n_res_s = want_pool ? 64 + (size_t)p->res_segs + (size_t)(k.pool_first + k.pool_count) : 4;
        FH_CUDA(cudaMemcpyAsync(p->h_res, p->d_res, sizeof(double) * n_res_s, cudaMemcpyDeviceToHost, ctx->stream));
        p->n_segments = 0;
        return FH_OK;
    }
    if (k.sector_dense) {
        // full-space tile kernels for the items before the fixed tail, then W, H, W^dagger and K3 on compressed vectors
        double *d_pool_out_s = p->d_res + 64 + p->res_segs;
        if (k.sector_prefix) {
            // the ops before the tail in ONE cluster kernel on the compressed state (DSMEM), which leaves psi_s in rank order
            FH_TRY(enqueue_payload_upload(p, false));
            FH_TRY(fh_sector_enqueue(p->sec_prefix, ctx, k.basis, p->d_pairs, p->d_dterms, p->d_res, nullptr, 0, 0, nullptr,
                                     fh_sector_dense_psi_buffer(p->sec_pool, pool)));
            FH_TRY(fh_sector_dense_enqueue(p->sec_dense, p->sec_pool, ctx, tables[0], nullptr, p->d_res, pool, k.pool_first, k.pool_count,
                                           d_pool_out_s));
        } else {
            FH_TRY(enqueue_payload_upload(p, capturing));
            p->chain_region = capturing ? 0 : 1;
            p->chain_used_runs[p->chain_region] = p->chain_used_maps[p->chain_region] = 0;
            FH_TRY(apply_range(p, 0, fh_sector_dense_tail_item(p->sec_dense), p->d_psi, 0, nullptr, -1, (long long)k.basis));
            FH_TRY(fh_sector_dense_enqueue(p->sec_dense, p->sec_pool, ctx, tables[0], p->d_psi, p->d_res, pool, k.pool_first, k.pool_count,
                                           d_pool_out_s));
        }
        const size_t n_res_s = want_pool ? 64 + (size_t)p->res_segs + (size_t)(k.pool_first + k.pool_count) : 4;
        FH_CUDA(cudaMemcpyAsync(p->h_res, p->d_res, sizeof(double) * n_res_s, cudaMemcpyDeviceToHost, ctx->stream));
        p->n_segments = 0;
        return FH_OK;
    }
    int first_param = n_items, last_param = -1;
    for (int i = 0; i < n_items; ++i)
        if (item_has_param(p, p->items[i])) {
            if (i < first_param) first_param = i;
            last_param = i;
        }
    // psi after items[0:chk_pos) is checkpointed so the fixed suffix is only unwound on lambda
    int chk_pos = 0;
    if (want_grads) chk_pos = last_param + 1;
    if (want_pool && k.pool_pos > chk_pos) chk_pos = k.pool_pos;

    FH_TRY(enqueue_payload_upload(p, capturing));
    p->chain_region = capturing ? 0 : 1;
    p->chain_used_runs[p->chain_region] = p->chain_used_maps[p->chain_region] = 0;
    double2 *psi = p->d_psi;
    // forward pass; the checkpoint of psi after items[0:chk_pos) (the fixed suffix is only unwound on lambda) and the
    // initial basis state ride along in the tile chains where they can
    if (need_adjoint && chk_pos < n_items)
        FH_TRY(apply_range(p, 0, n_items, psi, 0, p->d_chk, chk_pos, (long long)k.basis));
    else
        FH_TRY(apply_range(p, 0, n_items, psi, 0, nullptr, -1, (long long)k.basis));
    (void)bytes;
    for (int t = 0; t < k.n_tables; ++t) {
        const fh_table *tab = tables[t];
        double2 *h_out = (t == 0 && need_adjoint) ? p->d_lam : nullptr;
        // K2 in the sector: always when only <H> is wanted; with lambda (memset + scatter of the full vector) from 20 qubits on
        if (t == 0 && k.sector_k2 && (h_out == nullptr || p->n >= 20))
            FH_TRY(fh_sector_table_enqueue(p->sec_pool, ctx, tab, psi, h_out, p->d_res + 2 * t));
        else
            launch_apply_table(ctx->stream, ctx->sm_count, tab, psi, h_out, h_out ? 1 : 0, ctx->d_partials, p->d_res + 2 * t);
    }
    for (int v = 0; v < k.n_overlaps; ++v)
        launch_inner(ctx->stream, ctx->sm_count, targets[v]->d, psi, 1ull << p->n, ctx->d_partials,
                     p->d_res + 2 * k.n_tables + 2 * v);
    if (state_out) FH_CUDA(cudaMemcpyAsync(state_out->d, psi, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    if (!need_adjoint)
        FH_CUDA(cudaMemcpyAsync(p->h_res, p->d_res, sizeof(double) * 2 * (k.n_tables + k.n_overlaps + 1),
                                cudaMemcpyDeviceToHost, ctx->stream));
    double *d_pool_out = p->d_res + 64 + p->res_segs;          // pool outputs land next to the other results

    p->n_segments = 0;
    p->seg_param.clear();
    p->seg_scale.clear();
    if (need_adjoint) {
        double2 *lam = p->d_lam;
        if (want_grads && p->n_param_ops > 0)
            FH_CUDA(cudaMemsetAsync(p->d_gpart, 0, sizeof(double) * (size_t)p->n_param_ops * FH_GRAD_BLOCKS, ctx->stream));
        FH_TRY(apply_range(p, chk_pos, n_items, lam, 1));
        if (chk_pos < n_items) psi = p->d_chk;
        int stop = want_grads ? first_param : n_items;      // lowest item the sweep must undo
        if (want_pool && k.pool_pos < stop) stop = k.pool_pos;
        // K3: on sector-compressed copies of psi_s / lambda_s when the whole evaluation conserves (N_up, N_dn)
        auto enqueue_pool = [&](const double2 *ps, const double2 *lm) -> int {
            if (k.sector_pool) return fh_sector_pool_enqueue(p->sec_pool, ctx, ps, lm, pool, k.pool_first, k.pool_count, d_pool_out);
            return fh_enqueue_pool(pool, ps, lm, k.pool_first, k.pool_count, d_pool_out);
        };
        if (want_pool && k.pool_pos == chk_pos) FH_TRY(enqueue_pool(psi, lam));
        for (int i = chk_pos - 1; i >= stop; --i) {
            adjoint_item(p, p->items[i], psi, lam, want_grads);
            if (want_pool && k.pool_pos == i) FH_TRY(enqueue_pool(psi, lam));
        }
        if (p->n_segments > 0) launch_sum_strided(ctx->stream, p->d_gpart, FH_GRAD_BLOCKS, p->n_segments, p->d_gseg);
        // scalars, gradient segments and pool outputs in one copy
        const size_t n_res = 64 + (size_t)p->res_segs + (want_pool ? (size_t)(k.pool_first + k.pool_count) : 0);
        FH_CUDA(cudaMemcpyAsync(p->h_res, p->d_res, sizeof(double) * n_res, cudaMemcpyDeviceToHost, ctx->stream));
    }
    FH_CUDA(cudaGetLastError());
    return FH_OK;
}

static int ensure_workspaces(fh_program *p) {
    const size_t bytes = sizeof(double2) << p->n;
    if (!p->d_psi) FH_TRY(fh_ctx_scratch_get(p->ctx, bytes, reinterpret_cast<void **>(&p->d_psi)));
    if (!p->d_lam) FH_TRY(fh_ctx_scratch_get(p->ctx, bytes, reinterpret_cast<void **>(&p->d_lam)));
    if (!p->d_chk) FH_TRY(fh_ctx_scratch_get(p->ctx, bytes, reinterpret_cast<void **>(&p->d_chk)));
    return FH_OK;
}

extern "C" int fh_program_evaluate(fh_program *p, uint64_t basis_index, const double *thetas, int n_thetas, int n_tables,
                                   fh_table *const *tables, double *expvals, double *grads, const fh_pool *pool,
                                   int pool_pos, int pool_first, int pool_count, double *pool_out, int n_overlaps,
                                   fh_state *const *targets, double *overlaps, fh_state *state_out) {
    FH_REQUIRE(p, "fh_program_evaluate: program is NULL");
    FH_REQUIRE(p->finalized, "fh_program_evaluate: program not finalized");
    FH_REQUIRE(n_thetas == p->n_params && (n_thetas == 0 || thetas), "fh_program_evaluate: expected %d parameters, got %d",
               p->n_params, n_thetas);
    FH_REQUIRE(n_tables >= 1 && n_tables <= FH_MAX_RESULT_TABLES && tables && expvals,
               "fh_program_evaluate: need 1..%d observable tables", FH_MAX_RESULT_TABLES);
    FH_REQUIRE(n_overlaps >= 0 && n_overlaps <= FH_MAX_OVERLAPS, "fh_program_evaluate: at most %d overlap targets", FH_MAX_OVERLAPS);
    FH_REQUIRE(n_overlaps == 0 || (targets && overlaps), "fh_program_evaluate: NULL overlap arrays");
    FH_REQUIRE(basis_index < (1ull << p->n), "fh_program_evaluate: basis index out of range");
    for (int t = 0; t < n_tables; ++t)
        FH_REQUIRE(tables[t] && tables[t]->n == p->n, "fh_program_evaluate: table %d missing or wrong qubit count", t);
    for (int v = 0; v < n_overlaps; ++v)
        FH_REQUIRE(targets[v] && targets[v]->n == p->n, "fh_program_evaluate: target %d missing or wrong qubit count", v);
    FH_REQUIRE(!state_out || state_out->n == p->n, "fh_program_evaluate: state_out has wrong qubit count");
    if (pool) {
        FH_REQUIRE(pool->n == p->n && pool_out, "fh_program_evaluate: pool mismatch or pool_out NULL");
        FH_REQUIRE(pool_pos >= 0 && pool_pos <= (int)p->items.size(), "fh_program_evaluate: pool_pos out of range");
        FH_REQUIRE(pool_first >= 0 && pool_count >= 0 && pool_first + pool_count <= pool->n_out,
                   "fh_program_evaluate: pool range out of bounds");
    }
    fh_ctx *ctx = p->ctx;
    FH_CUDA(cudaSetDevice(ctx->device));
    FH_TRY(ensure_workspaces(p));
    if (pool && pool->n_out > p->res_pool_cap) {
        // grow the result buffers so the pool outputs fit behind the scalars and gradient segments
        FH_CUDA(cudaStreamSynchronize(ctx->stream));
        drop_graph(p);
        cudaFree(p->d_res);
        cudaFreeHost(p->h_res);
        p->d_res = p->h_res = nullptr;
        const size_t n_res = 64 + (size_t)p->res_segs + (size_t)pool->n_out;
        FH_CUDA(cudaMalloc(&p->d_res, sizeof(double) * n_res));
        FH_CUDA(cudaMallocHost(&p->h_res, sizeof(double) * n_res));
        p->d_gseg = p->d_res + 64;
        p->h_gseg = p->h_res + 64;
        p->res_pool_cap = pool->n_out;
    }

    // theta -> payload in the pinned staging buffers (the graph's first nodes copy them to the device)
    refresh_payload(p, thetas);

    EvalKey key;
    memset(&key, 0, sizeof(key));
    key.basis = basis_index;
    key.n_tables = n_tables;
    key.n_overlaps = n_overlaps;
    key.want_grads = grads != nullptr;
    for (int t = 0; t < n_tables; ++t) {
        key.tables[t] = tables[t];
        key.table_uid[t] = tables[t]->uid;
    }
    for (int v = 0; v < n_overlaps; ++v) {
        key.targets[v] = targets[v];
        key.target_uid[v] = targets[v]->uid;
    }
    key.pool = pool;
    key.pool_uid = pool ? pool->uid : 0;
    key.state_out_uid = state_out ? state_out->uid : 0;
    key.pool_pos = pool ? pool_pos : 0;
    key.pool_first = pool ? pool_first : 0;
    key.pool_count = pool ? pool_count : 0;
    key.state_out = state_out;
    // sector-resident path: energy / screening evaluations of number-conserving circuits whose sector fits one cluster
    p->last_sector = false;
    // (measured faster than the full-space path up to ~2 000 amplitudes, see sector_eval.cu; FHSIM_SECTOR=1 lifts the limit)
    if (!grads && n_overlaps == 0 && !state_out && n_tables == 1 && !(p->n & 1) && !getenv("FHSIM_NO_SECTOR")) {
        const int pool_flat = pool ? p->item_flat_first[pool_pos] : -1;
        const u64 max_dim = getenv("FHSIM_SECTOR") ? 0ull : 2048ull;
        FH_TRY(fh_sector_prepare(&p->sec, ctx, p->n, basis_index, p->pairs, p->diagops, p->dterms, p->flat, pool_flat, tables[0],
                                 pool, max_dim));
        key.sector = fh_sector_plan_eligible(p->sec) ? 1 : 0;
        p->last_sector = key.sector != 0;
    }
    p->last_sector_pool = false;
    p->last_sector_k2 = false;
    p->last_sector_dense = false;
    p->last_sector_prefix = false;
    if (!key.sector && n_tables >= 1 && !(p->n & 1) && p->n <= 31 && !getenv("FHSIM_NO_SECTOR_POOL")) {
        u64 upm = 0, dnm = 0;
        for (int b = 0; b < p->n; ++b) ((b & 1) ? upm : dnm) |= 1ull << b;          // even wires = up = odd index bits
        FH_TRY(fh_sector_pool_prepare(&p->sec_pool, ctx, p->n, upm, dnm, __builtin_popcountll(basis_index & upm),
                                      __builtin_popcountll(basis_index & dnm), p->pairs, p->flat, tables[0], pool));
        key.sector_pool = (pool && fh_sector_pool_plan_eligible(p->sec_pool)) ? 1 : 0;
        key.sector_k2 = fh_sector_pool_plan_table_ok(p->sec_pool) ? 1 : 0;
        // screening / energy-only calls whose trailing ops are a fixed single-species network (W): the whole tail in the sector
        if (key.sector_k2 && !grads && n_overlaps == 0 && !state_out && n_tables == 1 && (!pool || key.sector_pool)) {
            FH_TRY(fh_sector_dense_prepare(&p->sec_dense, p->sec_pool, p->n, p->pairs, p->diagops, p->dterms, p->flat, p->item_flat_first,
                                           pool ? pool_pos : 0, pool != nullptr));
            key.sector_dense = fh_sector_dense_ok(p->sec_dense) ? 1 : 0;
            // ... and the ops before it (the ansatz) in the cluster kernel: opt-in (FHSIM_SECTOR_PREFIX=1), see sector_eval.cu
            const int pf = key.sector_dense ? fh_sector_dense_first_flat(p->sec_dense) : 0;
            if (pf > 0 && getenv("FHSIM_SECTOR_PREFIX")) {
                FH_TRY(fh_sector_prepare(&p->sec_prefix, ctx, p->n, basis_index, p->pairs, p->diagops, p->dterms, p->flat, -1, tables[0],
                                         nullptr, 0, pf));
                key.sector_prefix = fh_sector_plan_eligible(p->sec_prefix) ? 1 : 0;
            }
        }
        p->last_sector_pool = key.sector_pool != 0;
        p->last_sector_k2 = key.sector_k2 != 0;
        p->last_sector_dense = key.sector_dense != 0;
        p->last_sector_prefix = key.sector_prefix != 0;
    }

    static const bool no_graph = getenv("FHSIM_NO_GRAPH") != nullptr;
    if (!p->ev0) {
        FH_CUDA(cudaEventCreate(&p->ev0));
        FH_CUDA(cudaEventCreate(&p->ev1));
    }
    if (no_graph) {
        const long long before = g_fh_launch_count;
        FH_CUDA(cudaEventRecord(p->ev0, ctx->stream));
        ++g_fh_tile_pdl_scope;
        const int rc = enqueue_evaluation(p, key, tables, pool, targets, state_out, false);
        --g_fh_tile_pdl_scope;
        FH_TRY(rc);
        FH_CUDA(cudaEventRecord(p->ev1, ctx->stream));
        p->last_launches = (int)(g_fh_launch_count - before);
    } else {
        if (!p->have_graph || !(p->key == key)) {
            drop_graph(p);
            const long long before = g_fh_launch_count;
            FH_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
            ++g_fh_tile_pdl_scope;
            const int rc = enqueue_evaluation(p, key, tables, pool, targets, state_out, true);
            --g_fh_tile_pdl_scope;
            cudaGraph_t g = nullptr;
            const cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
            if (rc != FH_OK) {
                if (g) cudaGraphDestroy(g);
                return rc;
            }
            if (e != cudaSuccess) {
                fh_set_error("fh_program_evaluate: stream capture failed: %s", cudaGetErrorString(e));
                return FH_ECUDA;
            }
            p->graph = g;
            FH_CUDA(cudaGraphInstantiate(&p->exec, p->graph, 0));
            p->key = key;
            p->have_graph = true;
            p->last_launches = (int)(g_fh_launch_count - before);
        }
        FH_CUDA(cudaEventRecord(p->ev0, ctx->stream));
        FH_CUDA(cudaGraphLaunch(p->exec, ctx->stream));
        FH_CUDA(cudaEventRecord(p->ev1, ctx->stream));
    }
    FH_CUDA(cudaStreamSynchronize(ctx->stream));
    {
        float ms = 0.f;
        FH_CUDA(cudaEventElapsedTime(&ms, p->ev0, p->ev1));
        p->last_ms = ms;
    }

    for (int t = 0; t < n_tables; ++t) expvals[t] = p->h_res[2 * t];
    for (int v = 0; v < 2 * n_overlaps; ++v) overlaps[v] = p->h_res[2 * n_tables + v];
    if (grads) {
        for (int q = 0; q < p->n_params; ++q) grads[q] = 0.0;
        for (int s = 0; s < p->n_segments; ++s) grads[p->seg_param[s]] += p->seg_scale[s] * p->h_gseg[s];
    }
    if (pool && pool_count > 0) memcpy(pool_out, p->h_res + 64 + p->res_segs + pool_first, sizeof(double) * pool_count);
    return FH_OK;
}

extern "C" int fh_program_payload_bytes(const fh_program *p, size_t *h2d_bytes, size_t *d2h_bytes) {
    FH_REQUIRE(p && p->finalized, "fh_program_payload_bytes: program missing or not finalized");
    if (h2d_bytes) *h2d_bytes = p->arena_bytes;
    if (d2h_bytes) *d2h_bytes = sizeof(double) * (64 + (size_t)p->res_segs + (size_t)p->res_pool_cap);
    return FH_OK;
}

extern "C" int fh_program_sector_info(const fh_program *p, int *active, int *cluster_size, uint64_t *sector_dim, int *n_ops,
                                      int *n_transposes, int *n_remote_ops) {
    FH_REQUIRE(p, "fh_program_sector_info: program is NULL");
    if (active) *active = p->last_sector ? 1 : ((p->last_sector_pool ? 2 : 0) | (p->last_sector_k2 ? 4 : 0) | (p->last_sector_dense ? 8 : 0) | (p->last_sector_prefix ? 16 : 0));
    u64 dim = 0;
    fh_sector_plan_describe(p->last_sector ? p->sec : nullptr, cluster_size, &dim, n_ops, n_transposes, n_remote_ops, nullptr);
    if (sector_dim) *sector_dim = dim;
    return FH_OK;
}

extern "C" int fh_program_last_stats(const fh_program *p, double *elapsed_ms, int *kernel_launches) {
    FH_REQUIRE(p, "fh_program_last_stats: program is NULL");
    if (elapsed_ms) *elapsed_ms = p->last_ms;
    if (kernel_launches) *kernel_launches = p->last_launches;
    return FH_OK;
}

// measurement: average device time (ms) of launching items [first, first+count) `reps` times back to back
// (payload must already be on the device, e.g. after fh_program_run / fh_program_evaluate)
extern "C" int fh_program_time_items(fh_program *p, fh_state *st, int first, int count, int dagger, int reps,
                                     double *ms_per_rep) {
    FH_REQUIRE(p && st && ms_per_rep, "fh_program_time_items: NULL argument");
    FH_REQUIRE(p->finalized && st->n == p->n, "fh_program_time_items: program not finalized or qubit mismatch");
    FH_REQUIRE(first >= 0 && count >= 0 && first + count <= (int)p->items.size() && reps >= 1,
               "fh_program_time_items: bad range");
    fh_ctx *ctx = p->ctx;
    if (!p->ev0) {
        FH_CUDA(cudaEventCreate(&p->ev0));
        FH_CUDA(cudaEventCreate(&p->ev1));
    }
    FH_CUDA(cudaEventRecord(p->ev0, ctx->stream));
    for (int r = 0; r < reps; ++r)
        for (int k = 0; k < count; ++k) apply_item(p, p->items[dagger ? first + count - 1 - k : first + k], st->d, dagger);
    FH_CUDA(cudaEventRecord(p->ev1, ctx->stream));
    FH_CUDA(cudaGetLastError());
    FH_CUDA(cudaEventSynchronize(p->ev1));
    float ms = 0.f;
    FH_CUDA(cudaEventElapsedTime(&ms, p->ev0, p->ev1));
    *ms_per_rep = ms / reps;
    return FH_OK;
}
