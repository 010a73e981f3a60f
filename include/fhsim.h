/* fhsim.h -- C-ABI of libfhsim.so: the B200 (sm_100a) complex128 statevector engine.
 *
 * This is the drop-in boundary for the hot path of
 * chuntse0514/Quantum-Simulation-of-Fermi-Hubbard-model.  The reference has no FFI of its own: the
 * seam is PennyLane's device/QNode API as its drivers use it.  Each entry point cites the reference
 * call it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - state = complex128[2^n] (interleaved re,im), wire/qubit q <-> bit (n-1-q) of the flat index
 *     (wire 0 = MSB; reference linalg/exact_diagonalization.py:23, PennyLane qml.state()).
 *   - Pauli string = (x, z): i^k X^x Z^z with k = popcount(x&z) (every Y = iXZ),
 *     (P psi)[i] = i^k (-1)^popcount((i^x)&z) psi[i^x].
 *   - all functions return 0 on success, a negative FH_E* code otherwise; fh_last_error() gives the
 *     thread-local message.  No C++ exception crosses the boundary.
 *   - host arrays are borrowed for the duration of the call; outputs are caller-allocated.
 *   - every call is ordered on the stream given to fh_ctx_create; calls that return host scalars
 *     synchronise that stream.  Handles are not thread-safe.
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with FH_ECUDA.
 */
#ifndef FHSIM_H
#define FHSIM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FH_OK 0
#define FH_EINVAL (-1)   /* bad argument                      -> ValueError   */
#define FH_ECUDA (-2)    /* CUDA runtime error / no device     -> RuntimeError */
#define FH_ENOMEM (-3)   /* host or device allocation failed   -> MemoryError  */
#define FH_ESTATE (-4)   /* object used in the wrong state     -> RuntimeError */

typedef struct fh_ctx fh_ctx;         /* device + stream + scratch                                 */
typedef struct fh_state fh_state;     /* complex128[2^n] on the device                             */
typedef struct fh_table fh_table;     /* observable: packed Pauli table grouped by x-mask           */
typedef struct fh_pool fh_pool;       /* screening pool: generators in pair form                    */
typedef struct fh_program fh_program; /* compiled circuit (ordered ops, parameters, fused tiles)   */

int fh_version(void);
const char *fh_last_error(void);

/* ---- context ------------------------------------------------------------------------------
 * replaces qml.device('default.qubit.torch' | 'lightning.gpu', wires=n)  (models/adapt_vqe.py:299-304).
 * stream: a cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream) or NULL for a private one. */
int fh_ctx_create(int device, void *stream, fh_ctx **out);
int fh_ctx_destroy(fh_ctx *ctx);
int fh_ctx_sync(fh_ctx *ctx);
int fh_ctx_info(fh_ctx *ctx, int *sm_count, size_t *free_bytes, size_t *total_bytes);
/* write `bytes` of scratch (>= L2 size) so the next kernel starts with a cold L2 (benchmark hygiene) */
int fh_ctx_flush_l2(fh_ctx *ctx, size_t bytes);
/* CUDA-event stopwatch on the context's stream (measurement only): start records an event, stop records a
 * second one, synchronises and returns the elapsed device time in milliseconds */
int fh_ctx_timer_start(fh_ctx *ctx);
int fh_ctx_timer_stop(fh_ctx *ctx, double *elapsed_ms);

/* ---- state --------------------------------------------------------------------------------
 * replaces the PennyLane state tensor and qml.state() (models/adapt_vqe.py:358-359, 404-407). */
int fh_state_create(fh_ctx *ctx, int n_qubits, fh_state **out);
/* borrow caller-owned device memory (e.g. a torch complex128 tensor's data_ptr()) */
int fh_state_wrap(fh_ctx *ctx, int n_qubits, void *device_ptr, fh_state **out);
int fh_state_destroy(fh_state *st);
int fh_state_set_basis(fh_state *st, uint64_t index);                  /* |index>, qml.PauliX prep (adapt_vqe.py:328-329) */
int fh_state_copy(fh_state *dst, const fh_state *src);
int fh_state_to_host(const fh_state *st, double *out_re_im);           /* 2*2^n doubles */
int fh_state_from_host(fh_state *st, const double *in_re_im);
int fh_state_device_ptr(const fh_state *st, void **ptr);
int fh_state_inner(const fh_state *a, const fh_state *b, double *re, double *im);   /* <a|b>, fidelity (adapt_vqe.py:408) */
int fh_state_norm2(const fh_state *st, double *out);
/* dst[i] = src[i with bits a[k] <-> b[k] exchanged] (disjoint pairs, out of place, stream-ordered, no sync): the
 * local half of a global<->local qubit swap of a state sharded over ranks by its top index bits (the other half is
 * an all-to-all of the 2^g top-local chunks, done by the host with NCCL).  No reference counterpart: the reference
 * is single-device (models/adapt_vqe.py:156). */
int fh_state_swap_bits(fh_state *dst, const fh_state *src, int n_pairs, const int32_t *a, const int32_t *b);

/* ---- K1: rotations, immediate mode ----------------------------------------------------------
 * pair op: for every index i with (i & fixmask) == fixval, j = i ^ x, s = (-1)^popcount(i & zeta):
 *     psi[i] <- m00 psi[i] + s m01 psi[j];   psi[j] <- s m10 psi[i] + m11 psi[j]
 * m = {re00,im00, re01,im01, re10,im10, re11,im11}.  fixmask must contain the top bit of x and fixval
 * must have it clear.  Covers exp(-i theta P/2) (models/utils.py:58-83), the fused 8-string
 * fermionic double-excitation rotation (models/adapt_vqe.py:87-98 on a pool generator),
 * qml.SingleExcitation / PauliX / CNOT / RX / RY (adapt_vqe.py:329,353; vqe_hea.py:47-52). */
int fh_apply_pair(fh_state *st, uint64_t x, uint64_t fixmask, uint64_t fixval, uint64_t zeta, const double m[8]);
/* diagonal op: psi[i] <- exp(-i sum_m angle[m] (-1)^popcount(i & z[m])) psi[i]   (qml.RZ, Z-string rotations) */
int fh_apply_diag(fh_state *st, int n_terms, const uint64_t *z, const double *angle);
/* sequence of single-string rotations  prod_m exp(-i half_angle[m] P_m), applied in order m = 0..M-1
 * (the literal Trotterize_generator loop, models/adapt_vqe.py:91-98, one launch per string) */
int fh_apply_pauli_rot_batch(fh_state *st, int m, const uint64_t *x, const uint64_t *z, const double *half_angles);

/* ---- K2: observables ------------------------------------------------------------------------
 * replaces QubitOperator_to_qmlHamiltonian + qml.expval(qml.Hamiltonian) (models/utils.py:30-56,
 * adapt_vqe.py:357,361).  Terms are given in table order; coeff is the coefficient of the Hermitian
 * string.  The table may be re-uploaded at any time (iQCC dressing, iqcc_hubbard.py:184-189). */
int fh_table_upload(fh_ctx *ctx, int n_qubits, int n_terms, const uint64_t *x, const uint64_t *z,
                    const double *coeff_re, const double *coeff_im, fh_table **out);
int fh_table_free(fh_table *tab);
int fh_table_info(const fh_table *tab, int *n_terms, int *n_groups);
/* Number of shared-memory tile passes fh_apply_table uses for this table (0: the gather kernel).  From 22 qubits on, a table
 * whose x-mask groups are covered by at most 3 sets of 12 index bits (the Hubbard Hamiltonian: the up-orbital bits and the
 * down-orbital bits) is applied pass by pass from tiles staged in shared memory: 80 B of HBM traffic per amplitude for two
 * passes instead of one gathered partner per (amplitude, term group).  FHSIM_K2_GATHER=1 forces the gather kernel. */
int fh_table_tile_passes(const fh_table *tab, int *n_passes);
/* out <- H in (out may be NULL: expectation only); *e = <in|H|in>.  in and out must differ. */
int fh_apply_table(const fh_table *tab, const fh_state *in, fh_state *out, double *e_re, double *e_im);
/* The same for a state confined to the (n_up, n_dn) sector (even wires = up) and a table that conserves both numbers: `in` is
 * gathered into rank order, H is applied to the compressed vector (partners through two rank tables; 3x4: 13.7 MB, L2-resident)
 * and H in is scattered back into `out` (zeroed first; may be NULL: expectation only).  FH_EINVAL when the table does not
 * qualify.  e_re == e_im == NULL: enqueue only.  fh_program_evaluate takes this route by itself (fh_program_sector_info). */
int fh_apply_table_sector(const fh_table *tab, const fh_state *in, fh_state *out, int n_up, int n_dn, double *e_re, double *e_im);
/* out <- out + H in  (sum of partial Hamiltonians, e.g. one table per qubit layout of a sharded state); *e = <in|H|in> */
int fh_apply_table_accumulate(const fh_table *tab, const fh_state *in, fh_state *out, double *e_re, double *e_im);

/* ---- K3: pool screening ---------------------------------------------------------------------
 * replaces ADAPT.select_operator's append-and-backprop (models/adapt_vqe.py:297-310):
 *   g[o] = sum over entries e with out_index[e]==o of
 *          2 Im sum_{i in pattern e} ( conj(lam_i) s B psi_j + conj(lam_j) s conj(B) psi_i ),  j = i^x
 * which equals 2 Im <lambda| G_o |psi> for a generator whose x-mask groups are the entries. */
int fh_pool_upload(fh_ctx *ctx, int n_qubits, int n_entries, const uint64_t *x, const uint64_t *fixmask,
                   const uint64_t *fixval, const uint64_t *zeta, const double *b_re, const double *b_im,
                   const int32_t *out_index, int n_out, fh_pool **out);
int fh_pool_free(fh_pool *pool);
/* gradients of outputs [first, first+count) -> out[count]   (pool sharding across GPUs uses first/count) */
int fh_pool_gradients(const fh_pool *pool, const fh_state *psi, const fh_state *lambda, int first, int count, double *out);

/* The same gradients for states that live in the (n_up, n_dn) sector (every amplitude outside it is zero: any ADAPT / HVA
 * state, models/adapt_vqe.py:325-361), evaluated on sector-compressed copies: psi and lambda are gathered into rank order
 * (rank = rank_up * D_dn + rank_dn, even wires = up) and the connected pairs of every entry are enumerated as the product
 * of two short lists (up patterns x down patterns matching the entry) -- 3x3: 1 225 pairs per operator instead of 2^15; 3x4:
 * both compressed vectors (13.7 MB) stay in L2.  FH_EINVAL if an entry does not conserve both particle numbers.
 * out == NULL: enqueue only.  fh_program_evaluate screens this way by itself when the whole evaluation conserves them. */
int fh_pool_gradients_sector(const fh_pool *pool, const fh_state *psi, const fh_state *lambda, int n_up, int n_dn, int first,
                             int count, double *out);
/* The same with an arbitrary assignment of index bits to species: up_mask / dn_mask = the index bits of the up / down orbitals
 * (disjoint, together all n bits, at most 16 per species; sectors of up to 16 384 patterns per spin).  This is the form a SLAB of
 * a sharded state needs: after global<->local qubit swaps the local bits are a permutation of the orbitals, and the slab of
 * rank r holds the sector (N_up - ups among the rank bits, N_dn - downs among them) of its local orbitals.  4x4 over 8 GPUs:
 * 22 M compressed amplitudes per slab instead of 2^29. */
int fh_pool_gradients_sector_masks(const fh_pool *pool, const fh_state *psi, const fh_state *lambda, uint64_t up_mask,
                                   uint64_t dn_mask, int n_up, int n_dn, int first, int count, double *out);

/* ---- compiled circuits ----------------------------------------------------------------------
 * replaces the QNode tape built by ADAPT.circuit / HVA.circuit / IQCC.get_circuit / VQE.circuit
 * (models/adapt_vqe.py:325-361, hva.py:273-303, iqcc_hubbard.py:59-80, vqe_hea.py:43-57).
 * Ops are appended in circuit order.  kind 0: fixed matrix m[8]; kind 1: rotation
 * exp(-i (scale*theta[param]) Ghat) with (Ghat psi)[i] = s bhat psi[j], (Ghat psi)[j] = s conj(bhat) psi[i]. */
int fh_program_create(fh_ctx *ctx, int n_qubits, int n_params, fh_program **out);
int fh_program_destroy(fh_program *prog);
int fh_program_add_pair(fh_program *prog, uint64_t x, uint64_t fixmask, uint64_t fixval, uint64_t zeta,
                        int kind, int param, double scale, double bhat_re, double bhat_im, const double m[8]);
/* diagonal op; param < 0: fixed angles coef[m]; else angle[m] = theta[param] * coef[m].  z[m] == 0 is a plain phase
 * exp(-i angle[m]) (it appears when a sharded state folds the rank bits of a Z string into a sign) */
int fh_program_add_diag(fh_program *prog, int n_terms, const uint64_t *z, const double *coef, int param);
/* ops added between begin/end are fused into ONE shared-memory tile kernel over the given bit positions
 * (ascending; every pair op inside must have its x-mask within those bits) */
int fh_program_begin_tile(fh_program *prog, int n_bits, const int32_t *bits);
int fh_program_end_tile(fh_program *prog);
int fh_program_finalize(fh_program *prog);
int fh_program_info(const fh_program *prog, int *n_ops, int *n_launches);
/* apply ops [first, first+count) to st (dagger != 0: the inverse, in reverse order) */
int fh_program_run(fh_program *prog, fh_state *st, const double *thetas, int n_thetas, int first, int count, int dagger);

/* One full evaluation (the body of one optimiser step / one screening), captured as a CUDA graph:
 *   psi = ops |basis_index>;  expvals[t] = <psi|tables[t]|psi>  (tables[0] is the cost Hamiltonian);
 *   if grads != NULL: grads[p] = d expvals[0] / d theta[p] by the adjoint sweep
 *     (replaces loss.backward() through the QNode, models/adapt_vqe.py:416-418, hva.py:324-326);
 *   if pool != NULL: pool_out[k] = 2 Im <lambda_s| G_k |psi_s> with psi_s the state after ops[0:pool_pos)
 *     and lambda_s the back-propagated H psi  (replaces select_operator, adapt_vqe.py:297-310);
 *   if n_overlaps > 0: overlaps[2v], [2v+1] = <targets[v]|psi>  (fidelity, adapt_vqe.py:404-408);
 *   if state_out != NULL it receives psi.
 * The fused tile kernels of the graph are launched with programmatic stream serialization (their set-up overlaps the
 * previous kernel's tail; FHSIM_NO_PDL=1 disables it); results are bit-identical run to run. */
int fh_program_evaluate(fh_program *prog, uint64_t basis_index, const double *thetas, int n_thetas,
                        int n_tables, fh_table *const *tables, double *expvals,
                        double *grads,
                        const fh_pool *pool, int pool_pos, int pool_first, int pool_count, double *pool_out,
                        int n_overlaps, fh_state *const *targets, double *overlaps,
                        fh_state *state_out);

/* Which path the most recent fh_program_evaluate call took.  Energy / screening evaluations (grads == NULL, one table, no
 * overlaps, no state_out) of circuits that conserve N_up and N_dn (every ADAPT / HVA circuit: models/adapt_vqe.py:325-361,
 * hva.py:273-303) run on the SECTOR-COMPRESSED state (3x3: 15 876 amplitudes instead of 2^18) resident in the distributed
 * shared memory of one thread-block cluster: one cluster kernel (ansatz, W, H, W^dagger) + one pool kernel instead of ~19
 * launches.  Chosen automatically when every op, the observable and the pool map the sector of |basis_index> to itself and
 * psi + lambda fit one cluster and the sector has <= 2 048 amplitudes (FHSIM_SECTOR=1: any size that fits; FHSIM_NO_SECTOR=1:
 * never).  Otherwise, when the circuit, the observable and the pool conserve (N_up, N_dn), the full-space path still screens
 * the pool (K3) on sector-compressed copies of psi_s / lambda_s (3x3: 1 225 instead of 2^15 pairs per operator; 3x4: the
 * compressed vectors are L2-resident); FHSIM_NO_SECTOR_POOL=1 keeps K3 in the full space.
 * K2 of tables[0] takes the same shortcut (compress, gather over the term groups on the compressed vector, scatter of H psi back
 * when the adjoint part needs it: from 20 qubits on; FHSIM_NO_SECTOR_K2=1 disables it).
 * Screening / energy-only calls whose trailing ops are a FIXED network of single-species ops (the basis change W of the drivers,
 * models/adapt_vqe.py:344-356) run that whole tail on compressed vectors: W = S (U_up (x) U_dn) S with two dense sector blocks
 * built once per program, then H, W^dagger and K3 (sectors of <= 512 patterns per spin; FHSIM_NO_SECTOR_DENSE=1 disables it).
 * active: 1 if the last call ran in the cluster; otherwise a bit set: 2 = K3 in the sector, 4 = K2 in the sector, 8 = dense
 * tail; 0 = full space;
 * cluster_size: CTAs of the cluster; sector_dim: amplitudes; n_ops: steps of the cluster kernel (ops, transposes,
 * checkpoint, H, store); n_transposes / n_remote_ops: layout changes / ops that exchange amplitudes between CTAs. */
int fh_program_sector_info(const fh_program *prog, int *active, int *cluster_size, uint64_t *sector_dim, int *n_ops,
                           int *n_transposes, int *n_remote_ops);

/* measurement: bytes one fh_program_evaluate call copies host->device (the theta-dependent op payload, one pinned
 * arena) and at most device->host (scalars + gradient segments + pool outputs) */
int fh_program_payload_bytes(const fh_program *prog, size_t *h2d_bytes, size_t *d2h_bytes);
/* measurement: device time (CUDA events bracketing the graph launch on the context's stream) and number of
 * kernel launches of the most recent fh_program_evaluate call */
int fh_program_last_stats(const fh_program *prog, double *elapsed_ms, int *kernel_launches);
/* measurement: average device milliseconds of launching items [first, first+count) `reps` times back to back */
int fh_program_time_items(fh_program *prog, fh_state *st, int first, int count, int dagger, int reps, double *ms_per_rep);

/* ---- K4: Lanczos ground states --------------------------------------------------------------
 * replaces linalg/exact_diagonalization.py:34-51, 181-229 (scipy eigsh, which='SA') and
 * openfermion.get_ground_state (models/iqcc_hubbard.py:57).  Matrix-free on fh_apply_table.
 * n_up/n_dn >= 0 restrict to the (N_up, N_dn) sector (even wires = up); -1 = full space.
 * Finds the k lowest eigenpairs (degenerate levels resolved by deflation); evals[k];
 * evecs (may be NULL) receives k device states. */
int fh_lanczos(const fh_table *tab, int n_up, int n_dn, int k, double tol, int max_iter, uint64_t seed,
               double *evals, fh_state *const *evecs, int *iterations);
/* The same eigenproblem on SECTOR-COMPRESSED vectors (one amplitude per (N_up, N_dn) state, rank = rank_up * D_dn +
 * rank_dn), iteration scalars kept on the device: what linalg/exact_diagonalization.py:26-51 does when it restricts the
 * sparse matrix to the sector before calling eigsh.  3x3: 15 876 amplitudes per vector instead of 2^18; 4x4: 165 636 900
 * (2.65 GB) instead of 64 GiB, so the 32-qubit ground state of models/adapt_vqe.py:221-247 fits one GPU.  The table must
 * conserve both particle numbers (FH_EINVAL otherwise); fh_lanczos takes this path by itself whenever it applies.
 * evecs (may be NULL): k full 2^n device states (amplitudes outside the sector are zero).
 * compressed_out (may be NULL): host buffer of k * dim_sector complex128 values in rank order.
 * stats (may be NULL): double[4] = sector dimension, seconds in the iteration loops, matvecs, host synchronisations. */
int fh_lanczos_sector(const fh_table *tab, int n_up, int n_dn, int k, double tol, int max_iter, uint64_t seed,
                      double *evals, fh_state *const *evecs, double *compressed_out, int *iterations, double *stats);

/* ---- iQCC Hamiltonian dressing on packed Pauli tables (device) --------------------------------------------
 * replaces the symbolic update of models/iqcc_hubbard.py:184-189,
 *     H <- H + sin(tau)(-i/2)[H, P] + (1/2)(1 - cos tau)(P H P - H)  =  exp(i tau P/2) H exp(-i tau P/2),
 * on (x-mask, z-mask, coefficient) arrays resident on the device: XOR of masks, popcount phases, hash merge of the one
 * possible duplicate per string, ordered compaction.  Bit-identical to the host restatement PauliTable.dressed. */
typedef struct fh_ptable fh_ptable;
int fh_ptable_upload(fh_ctx *ctx, int n_qubits, int n_terms, const uint64_t *x, const uint64_t *z, const double *coeff_re,
                     const double *coeff_im, fh_ptable **out);          /* distinct strings (a canonical table) */
int fh_ptable_free(fh_ptable *table);
int fh_ptable_size(const fh_ptable *table, int *n_terms);
int fh_ptable_download(const fh_ptable *table, uint64_t *x, uint64_t *z, double *coeff_re, double *coeff_im);
int fh_ptable_dress(fh_ptable *table, uint64_t xp, uint64_t zp, double tau, double tol);   /* P = i^k X^xp Z^zp */

/* ---- multi-GPU: communicator for the sharded-state path (BASELINE cfg 5) -------------------
 * The reference is single-process / single-device (models/adapt_vqe.py:156 hard-codes cuda:0), so these entry points
 * have no reference counterpart; they carry the global<->local qubit swap of a state sharded by its top index bits.
 * One process per GPU.  Rank 0 creates the id, the launcher distributes the 128 bytes (any channel: file, socket,
 * torch.distributed broadcast), every rank calls fh_comm_init.  NCCL (NVLink 5 / NVSwitch) is loaded at run time. */
typedef struct fh_comm fh_comm;
int fh_comm_unique_id(unsigned char *out128);
int fh_comm_init(fh_ctx *ctx, const unsigned char *id128, int rank, int world, fh_comm **out);
int fh_comm_destroy(fh_comm *comm);
int fh_comm_info(const fh_comm *comm, int *rank, int *world, double *last_exchange_ms);
/* Global<->local qubit swap of a slab: optional local bit permutation (n_pairs disjoint transpositions a[k]<->b[k] of
 * index bits, staged in `scratch`) followed by the all-to-all of the world equal chunks: dst chunk p <- chunk `rank` of
 * rank p.  Chunk k+1 is permuted on the compute stream while chunk k is on the wire (communication stream).  Enqueues
 * only; later work on the context's stream is ordered after the exchange.  src, scratch, dst: distinct slabs. */
int fh_comm_swap_exchange(fh_comm *comm, const fh_state *src, fh_state *scratch, fh_state *dst, int n_pairs,
                          const int32_t *a, const int32_t *b);
int fh_comm_last_exchange_ms(fh_comm *comm, double *ms);
int fh_comm_all_reduce_sum(fh_comm *comm, double *values, int count);       /* host in/out, count <= 4096 */
int fh_comm_all_gather(fh_comm *comm, const double *in, int count, double *out);   /* host; out[world * count] */
int fh_comm_barrier(fh_comm *comm);

#ifdef __cplusplus
}
#endif
#endif /* FHSIM_H */
