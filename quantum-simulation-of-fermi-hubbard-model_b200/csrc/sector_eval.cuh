// Sector-resident evaluation (sector_eval.cu): host interface used by program.cu / api.cu.
#pragma once
#include <vector>

#include "common.cuh"

struct fh_sector_plan;

struct SecFlatOp {
    int type;      // 1 pair (index into the program's PairOp array), 2 diag (index into its DiagOp array)
    int index;
};

// Build or reuse the plan of one program for (sector of `basis`, observable, pool, pool position in flat op units).
// Not stream-ordered (allocates and uploads): call outside stream capture.
int fh_sector_prepare(fh_sector_plan **slot, fh_ctx *ctx, int n, u64 basis, const std::vector<PairOp> &pairs,
                      const std::vector<DiagOp> &diagops, const std::vector<DiagTerm> &dterms, const std::vector<SecFlatOp> &flat,
                      int pool_flat, const fh_table *tab, const fh_pool *pool, u64 max_dim, int prefix_flat = -1);
// max_dim: sectors with more amplitudes are not planned (0: no limit)
bool fh_sector_plan_eligible(const fh_sector_plan *plan);
void fh_sector_plan_describe(const fh_sector_plan *plan, int *cluster, u64 *dim, int *nvops, int *transposes, int *remote_ops,
                             size_t *smem);
// Enqueue (capturable): E -> d_res[0..1]; pool outputs o in [first, first+count) -> d_pool_out[o]
int fh_sector_enqueue(fh_sector_plan *plan, fh_ctx *ctx, u64 basis, const PairOp *d_pairs, const DiagTerm *d_dterms, double *d_res,
                      const fh_pool *pool, int pool_first, int pool_count, double *d_pool_out, double2 *chk_override = nullptr);
void fh_sector_plan_free(fh_sector_plan *plan);
void fh_sector_forget_table(u64 uid);
void fh_sector_forget_pool(u64 uid);

// K3 in the sector for the full-space path: compress psi_s / lambda_s (full 2^n states) into rank order, then the pool kernel
struct fh_sector_pool_plan;
int fh_sector_pool_prepare(fh_sector_pool_plan **slot, fh_ctx *ctx, int n, u64 upmask, u64 dnmask, int n_up, int n_dn,
                           const std::vector<PairOp> &pairs, const std::vector<SecFlatOp> &flat, const fh_table *tab,
                           const fh_pool *pool);
bool fh_sector_pool_plan_eligible(const fh_sector_pool_plan *plan);
int fh_sector_pool_enqueue(fh_sector_pool_plan *plan, fh_ctx *ctx, const double2 *psi, const double2 *lam, const fh_pool *pool,
                           int pool_first, int pool_count, double *d_pool_out);
void fh_sector_pool_plan_free(fh_sector_pool_plan *plan);
void fh_sector_forget_pool_plan(u64 uid);
bool fh_sector_pool_plan_table_ok(const fh_sector_pool_plan *plan);
// K2 of a sector-confined full-space state on its compressed copy: out (may be NULL) <- H in, E -> d_result[0..1]
int fh_sector_table_enqueue(fh_sector_pool_plan *plan, fh_ctx *ctx, const fh_table *tab, const double2 *in, double2 *out,
                            double *d_result);
void fh_sector_forget_table_plan(u64 uid);

// Dense tail (sector_eval.cu): a trailing run of fixed single-species ops as two dense sector transforms, so W, H, W^dagger
// and K3 of a screening all run on compressed vectors
struct fh_sector_dense;
int fh_sector_dense_prepare(fh_sector_dense **slot, const fh_sector_pool_plan *plan, int n, const std::vector<PairOp> &pairs,
                            const std::vector<DiagOp> &diagops, const std::vector<DiagTerm> &dterms, const std::vector<SecFlatOp> &flat,
                            const std::vector<int> &item_flat_first, int min_item, bool exact);
bool fh_sector_dense_ok(const fh_sector_dense *d);
int fh_sector_dense_tail_item(const fh_sector_dense *d);
int fh_sector_dense_enqueue(fh_sector_dense *d, fh_sector_pool_plan *plan, fh_ctx *ctx, const fh_table *tab, const double2 *psi_full,
                            double *d_result, const fh_pool *pool, int pool_first, int pool_count, double *d_pool_out);
void fh_sector_dense_free(fh_sector_dense *d);
double2 *fh_sector_dense_psi_buffer(fh_sector_pool_plan *plan, const fh_pool *pool);
int fh_sector_dense_first_flat(const fh_sector_dense *d);
