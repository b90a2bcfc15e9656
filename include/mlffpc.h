/* libmlffpc -- C ABI of the B200-native fp64 PCG solve path for sGDML Hessian-kernel systems.
 *
 * The reference (bluecher31/mlff-preconditioner) is pure Python and has no FFI; this header is the
 * boundary a maintainer binds with ctypes/cffi.  Every entry point names the
 * reference routine it replaces (paths relative to /root/reference/src/sGDML/sgdml/).
 *
 * Conventions
 *  - plain C types only; all array arguments are CALLER-OWNED DEVICE pointers (e.g. torch
 *    tensor.data_ptr()), row-major, fp64 / int32 / int64 as stated; `stream` is a cudaStream_t
 *    passed as void* (torch.cuda.current_stream().cuda_stream); calls are stream-ordered.
 *  - every call returns 0 (MLFFPC_OK) or a negative status; the message is in mlffpc_last_error().
 *    No exceptions cross the ABI.  There is no CPU fallback anywhere in this library.
 *  - workspace sizes are queried first (*_workspace_bytes) and the caller allocates.
 *  - notation: N atoms, D = N(N-1)/2, S permutations, M training points, dim_i = 3N, n = 3NM.
 *    A context is sharded by training points: local rows are points [pt0, pt1), n_local = 3N(pt1-pt0).
 */
#ifndef MLFFPC_H
#define MLFFPC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MLFFPC_OK 0
#define MLFFPC_ERR_INVALID (-1)   /* bad argument                         -> ValueError      */
#define MLFFPC_ERR_CUDA (-2)      /* CUDA runtime error                   -> RuntimeError    */
#define MLFFPC_ERR_NOT_PSD (-3)   /* pivot <= 0 (incomplete_cholesky.py:62) -> AssertionError */
#define MLFFPC_ERR_LINALG (-4)    /* Cholesky breakdown (scipy LinAlgError) -> LinAlgError   */
#define MLFFPC_ERR_COMM (-5)      /* NCCL error                           -> RuntimeError    */
#define MLFFPC_ERR_UNSUPPORTED (-6) /* outside the hot-path contract      -> NotImplementedError */

typedef struct mlffpc_ctx mlffpc_ctx;

/* ---------------------------------------------------------------- lifecycle ---- */
int mlffpc_version(void);
int64_t mlffpc_launch_count(void); /* kernels launched by this library so far (process-wide) */
const char* mlffpc_last_error(void);
int mlffpc_create(mlffpc_ctx** out, int device);
int mlffpc_destroy(mlffpc_ctx* ctx);

/* Multi-GPU (one process per GPU).  NCCL is bound with dlopen(libnccl_path) -- pass the libnccl.so.2
 * that torch already loaded.  Rank 0 makes the 128-byte id, the host broadcasts it (torch.distributed),
 * every rank calls comm_init.  The reference has no collective at all. */
int mlffpc_comm_unique_id(const char* libnccl_path, void* id128_out);
int mlffpc_comm_init(mlffpc_ctx* ctx, const char* libnccl_path, const void* id128, int rank, int world);
int mlffpc_allreduce_sum(mlffpc_ctx* ctx, double* buf, int64_t count, void* stream);
int mlffpc_allgather(mlffpc_ctx* ctx, const void* send, void* recv, int64_t bytes_per_rank, void* stream);

/* NVLink peer-memory collectives (optional, after comm_init and set_geometry; csrc/peer.cuh).  Every rank allocates one
 * communication buffer and exports a 64-byte CUDA IPC handle; the host gathers the handles (torch.distributed) and
 * every rank maps all of them.  From then on the small collectives of the inner loops are fused into the solver's own
 * kernels: the dot-product kernels push their scalars into the peers' slots and the update kernels combine them, the
 * p-update kernel stores the search direction into every rank's replicated vector, the symmetric operator's finish
 * kernel pulls the peers' partial products (fused reduce-scatter), the pivot step's prepare kernel pushes
 * {candidate, factor row} -- no library collective per CG iteration / pivot step.  k_max bounds the k-vector of the
 * preconditioner apply that can travel this way (larger k falls back to ncclAllReduce).  When export or import fails on
 * ANY rank, call mlffpc_peer_disable on ALL ranks: the NCCL path is the fallback. */
int mlffpc_peer_export(mlffpc_ctx* ctx, int64_t k_max, void* handle64_out);
int mlffpc_peer_import(mlffpc_ctx* ctx, const void* handles /* world * 64 bytes, rank order */, int count);
int mlffpc_peer_disable(mlffpc_ctx* ctx);

/* ---------------------------------------------------------------- geometry ---- */
/* Replaces the per-call re-upload of descriptors in GDMLTorchPredict.__init__ (torchtools.py:80-97)
 * and the (R_desc, R_d_desc, tril_perms_lin) arguments every reference routine takes.
 * R_desc[M,D], R_d_desc[M,D,3] fp64; desc_perms[S,D] int32 = pi_p(d) (tril_perms_lin de-linearised,
 * train.py:783-790); atom_perms[S,N] int32 = task['perms'].  The input buffers must stay alive.
 * R_d_desc may be NULL for a prediction-only context (mlffpc_predict with beta).
 * [pt0, pt1) is this context's row block of training points (whole range for one GPU). */
int mlffpc_geometry_workspace_bytes(int64_t M, int N, int S, int64_t* bytes);
int mlffpc_set_geometry(mlffpc_ctx* ctx, int64_t M, int N, int S, const double* R_desc,
                        const double* R_d_desc, const int32_t* desc_perms, const int32_t* atom_perms,
                        double sig, int64_t pt0, int64_t pt1, void* workspace, int64_t workspace_bytes,
                        void* stream);

/* ---------------------------------------------------------------- kernel entries ---- */
/* -diag(K) for the local rows -> out[n_local].  Replaces IterativeCholesky._assemble_kernel_mat_diag
 * (solvers/iterative_cholesky.py:241-373). */
int mlffpc_kernel_diag(mlffpc_ctx* ctx, double* out, void* stream);

/* Explicit kernel rows: K_out[r, c] = K[row0 + r, c] for all n columns, row-major with leading
 * dimension ld >= n.  Replaces GDMLTrain._assemble_kernel_mat(col_idxs=np.s_[:]) (train.py:1121-1308,
 * worker :81-236). */
int mlffpc_kernel_assemble(mlffpc_ctx* ctx, double* K_out, int64_t ld, void* stream);

/* Column panel, transposed: out[c, r] = scale * K[row0 + r, cols[c]], c < b, leading dimension
 * ld >= n_local; cols is a device int64[b] of global column indices (any order).  Replaces
 * _assemble_kernel_mat(col_idxs=list) (train.py:1237-1263) and the one-column trick
 * IterativeCholesky._get_col_K = K_op.matvec(e_i) (solvers/iterative_cholesky.py:152-156).
 * Workspace: mlffpc_kernel_columns_workspace_bytes(ctx, b). */
int mlffpc_kernel_columns_workspace_bytes(mlffpc_ctx* ctx, int64_t b, int64_t* bytes);
int mlffpc_kernel_columns(mlffpc_ctx* ctx, const int64_t* cols, int64_t b, double* out, int64_t ld,
                          double scale, void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- operators ---- */
/* Assembled matvec  y[r] = alpha * sum_c K[r, c] x[c] + shift * x[x_off + r]   (r < n_rows).
 * With alpha = -1, shift = lam this is the reference's (-K_op).matvec (iterative_solver.py:416-443,
 * :996) on an explicit K.  HBM-bound: 8*n_rows*n_cols bytes per call. */
int mlffpc_gemv(mlffpc_ctx* ctx, const double* K, int64_t n_rows, int64_t n_cols, int64_t ld,
                const double* x, double* y, double alpha, double shift, int64_t x_off, void* stream);

/* Symmetric assembled matvec on one square: y = alpha * K x + shift * x reading only the lower triangle of
 * the symmetric K (by 32-row strips; each entry is used for y[r] and y[c]) -- about half the HBM traffic of
 * mlffpc_gemv.  Deterministic (no atomics). */
int mlffpc_symv_workspace_bytes(int64_t n, int64_t* bytes);
int mlffpc_symv(mlffpc_ctx* ctx, const double* K, int64_t n, int64_t ld, const double* x, double* y,
                double alpha, double shift, void* workspace, int64_t workspace_bytes, void* stream);

/* Symmetric tile operator (any number of ranks): the G x G grid of point blocks is dealt out so that each
 * unordered block pair is stored on exactly one rank and every rank holds ~ n_local * n / 2 entries
 * (rank g: the diagonal tile, tiles (g, g+d) for d < G/2 and half of tile (g, g+G/2) when G is even).
 * One matvec reads every stored entry once (rows and columns of the product), then one reduce-scatter of n
 * doubles combines the ranks.  Same result as mlffpc_gemv on the assembled row block up to summation order;
 * half the HBM bytes and half the memory (n = 270 000 fits on 4 B200s).
 *   storage_elems: doubles the caller allocates for Ksym (16-byte aligned)
 *   tiles: 8 int64 per tile {i_pt0, i_pt1, j_pt0, j_pt1, ld, offset, is_diagonal, 0}  (out may be NULL to count)
 *   assemble: fills Ksym from the geometry (replaces train.py:1121-1308 incl. its exploit_sym mirror, :207-210)
 *   apply: y_local = alpha * (K x)_local + shift * x_local; with partial_out != NULL the rank's full-length
 *          partial product (world * n_pad doubles) is written there instead and no collective is issued.
 * mlffpc_set_option(ctx, "symmetric_gemv", 1) makes mlffpc_pcg treat K_local as this storage.
 * Options "layout_world"/"layout_rank" override the partition taken from the communicator (rank emulation). */
int mlffpc_symop_storage_elems(mlffpc_ctx* ctx, int64_t* elems);
int mlffpc_symop_tiles(mlffpc_ctx* ctx, int64_t* out, int64_t max_tiles, int64_t* n_tiles);
int mlffpc_symop_assemble(mlffpc_ctx* ctx, double* Ksym, void* stream);
int mlffpc_symop_workspace_bytes(mlffpc_ctx* ctx, int64_t* bytes);
int mlffpc_symop_apply(mlffpc_ctx* ctx, const double* Ksym, const double* x_full, double* y_local, double alpha,
                       double shift, void* workspace, int64_t workspace_bytes, double* partial_out, void* stream);
/* Library switches (all optional; defaults in brackets):
 *   "symmetric_gemv"  [0]  mlffpc_pcg treats K_local as the symmetric tile storage
 *   "layout_world", "layout_rank"   tile partition override (rank emulation on one GPU)
 *   "pchol_graph"     [1]  look-ahead build: replay each chunk of 8 pivot steps as one CUDA-graph launch (the device owns the
 *                          step counter); 0 = launch the two kernels of every step individually
 *   "pchol_lookahead" [1]  blocked pivoted Cholesky with a candidate panel (0 = plain left-looking build)
 *   "assemble_legacy" [0]  first-generation assembly kernel (one CTA per 3N x 3N block)
 *   "gram_mode"       [1]  Gram matrices (mlffpc_syrk_rows) with (hi, lo) accumulation of the k-tile products;
 *                          0 = one running fp64 sum per entry (the round-1 kernel: 188 instead of the reference's 119
 *                          CG iterations on BASELINE.json configs[0])
 *   "gram_fold"       [1]  k-tiles (16 columns each) of DMMA accumulation per (hi, lo) fold: 1, 2 or 4 (cfg2: 920 / 945 /
 *                          1124 CG iterations -- the accuracy of E decides, so the default folds every k-tile)
 *   "pairs_kernel"    [0]  pair stage of the matrix-free operator / prediction: 1 = 64 x 64 tiles, synchronous staging;
 *                          2 = 128 x 64 tiles with a cp.async ring (8 x 4 pairs per thread, 8 warps per SM); 3 = 128 x 32
 *                          tiles (8 x 2 pairs per thread, 16 warps per SM); 0 = 1 for D < 64, else 3 (measured)
 *   "peer_kvec", "peer_pivots" [1]  use the mapped peer buffers for the apply's k-vector sum / the pivot-step message
 *   "tma_rows"        [1]  "T r" of the preconditioner apply on the TMA-fed row-strip kernel (csrc/symtma.cu); 0 = the
 *                          register-staged 4-row GEMV of round 1
 *   "defect_mode"     [1]  E = Q Q^T - I of the projected form from the DMMA kernel (1) or with exact products and sums
 *                          on the FP64 vector pipe (2, ~7x slower; the reference for the tests)
 *   "precon_reorth"   [0]  project the complement twice in the orthonormal-form preconditioner apply (Mk given, no E):
 *                          two more passes over the factor; the four-pass cross-check of the projected form
 *   diagnostics for the numerics study in DESIGN.md: "precon_accuracy" (1 Kahan, 2 two-halves summation of T r),
 *   "tgemv_msplit" (1/4/8), "syrk_chunk" (> 0: Gram by Kahan-summed column chunks) */
int mlffpc_set_option(mlffpc_ctx* ctx, const char* name, int64_t value);

/* Matrix-free matvec  y_local = alpha * (K v)_local + shift * v_local, v is the full n-vector.
 * Replaces GDMLPredict.set_alphas + predict (predict.py:400-449, 997-1052) ->
 * GDMLTorchPredict.set_alphas/_forward (torchtools.py:128-151, 172-272), numpy twin
 * _predict_wkr (predict.py:72-234), and desc.d_desc_dot_vec / vec_dot_d_desc (utils/desc.py:394-428). */
int mlffpc_matvec_free_workspace_bytes(mlffpc_ctx* ctx, int64_t* bytes);
int mlffpc_matvec_free(mlffpc_ctx* ctx, const double* v, double* y_local, double alpha, double shift,
                       void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- prediction / caller-side prep ---- */
/* Descriptors and compressed Jacobians of B geometries R[B, N, 3] (device): R_desc[B, D] = 1 / |r_a - r_b|,
 * R_d_desc[B, D, 3] = (r_a - r_b) / |r_a - r_b|^3 over the pairs d <-> (a_d > b_d) of np.tril_indices(N, -1).
 * Replaces Desc.from_R (utils/desc.py:292-358, workers :112-200) -- one-off host prep of GDMLTrain.train
 * (train.py:813-819) and of every prediction (predict.py:1083).  No geometry needs to be set. */
int mlffpc_desc_from_r(mlffpc_ctx* ctx, const double* R, int64_t B, int N, double* R_desc, double* R_d_desc, void* stream);
/* out[M, D] = J_m v_m for the full n-vector v: beta of the matvec and the model's R_d_desc_alpha.  Replaces
 * desc.d_desc_dot_vec (utils/desc.py:394-405) in GDMLTrain.create_model (train.py:640-645). */
int mlffpc_d_desc_dot_vec(mlffpc_ctx* ctx, const double* v, double* out, void* stream);
/* Energies and forces of B query geometries from the model whose training geometry is set in ctx and whose
 * coefficients are given either as v = alphas_F[n] (beta = NULL) or as beta[M, D] = J_j alpha_j = model['R_d_desc_alpha']
 * (v = NULL; train.py:640-645 -- then the context may have been set without R_d_desc):   F_out[B, 3N] = J_q^T f_q  and  E_out[B] = sum_jp e^(1 + rho^) (Delta . beta_jp)
 * (unscaled: the caller applies model['std'] and model['c']).  Replaces GDMLPredict.predict for arbitrary queries
 * (predict.py:997-1110 -> _predict_wkr :72-234 / GDMLTorchPredict._forward torchtools.py:172-272, incl. the energy
 * einsum :268) and, with the training geometries as queries, the energy pass of GDMLTrain._recov_int_const
 * (train.py:972-1119).  E_out may be NULL.  Rq_desc / Rq_d_desc as produced by mlffpc_desc_from_r. */
int mlffpc_predict_workspace_bytes(mlffpc_ctx* ctx, int64_t B, int64_t* bytes);
int mlffpc_predict(mlffpc_ctx* ctx, const double* Rq_desc, const double* Rq_d_desc, int64_t B, const double* v,
                   const double* beta, double* F_out, double* E_out, void* workspace, int64_t workspace_bytes,
                   void* stream);

/* ---------------------------------------------------------------- dense building blocks ---- */
/* Row-major fp64 GEMM on the FP64 tensor pipe (DMMA): C = alpha * A * op(B) + beta * C,
 * A[m,k], op(B) = B[k,n] (trans_b = 0) or B[n,k]^T (trans_b = 1).  Stands in for the host BLAS calls
 * L.T @ L, K_nm.T.dot(K_nm), B.dot(B.T) (iterative_cholesky.py:141, iterative_solver.py:230, :535). */
int mlffpc_dgemm(mlffpc_ctx* ctx, int trans_b, int64_t m, int64_t n, int64_t k, double alpha,
                 const double* A, int64_t lda, const double* B, int64_t ldb, double beta, double* C,
                 int64_t ldc, void* stream);
/* W[m,m] = X X^T + shift * I  for X[m, n_cols] row-major (both triangles written; summed over ranks
 * when a communicator is attached). */
int mlffpc_syrk_rows(mlffpc_ctx* ctx, const double* X, int64_t m, int64_t n_cols, int64_t ldx,
                     double shift, double* W, int64_t ldw, void* stream);
/* In-place lower Cholesky of W[m,m] (upper triangle zeroed).  info_host: 0 ok, j > 0 = breakdown at
 * column j.  Stands in for scipy.linalg.cholesky / cho_factor (iterative_cholesky.py:142,
 * iterative_solver.py:580). Synchronises the stream to return info. */
int mlffpc_potrf_lower(mlffpc_ctx* ctx, double* W, int64_t m, int64_t ldw, int* info_host, void* stream);
/* X <- Lf^{-1} X for lower-triangular Lf[m,m] and X[m, n_cols] row-major.  Stands in for
 * scipy.linalg.solve_triangular (iterative_cholesky.py:143, iterative_solver.py:219-226, :263-270). */
int mlffpc_trsm_rows(mlffpc_ctx* ctx, const double* Lf, int64_t m, int64_t ldl, double* X,
                     int64_t n_cols, int64_t ldx, void* stream);

/* ---------------------------------------------------------------- pivoted partial Cholesky ---- */
/* Greedy diagonal-pivoted partial Cholesky of A = -K (columns generated on the fly from the geometry).
 * Replaces incomplete_cholesky.pivoted_cholesky (solvers/incomplete_cholesky.py:24-93) together with
 * its column oracle (iterative_cholesky.py:152-156).
 *   Lt[k, n_local] (ld >= n_local): the factor TRANSPOSED, Lt[m, r] = L[row0 + r, m]
 *   diag[n_local]: in = -diag(K) local rows, out = residual diagonal
 *   index_columns[n] int64 (replicated, out): the reference's permutation, first k = pivot sequence
 *   forced_pivots: NULL, or device int64[k] to replay a pivot sequence (tie diagnostics)
 *   step_ms_host: NULL, or host float[k] receiving per-step times (info['time_cholesky'])
 * Returns MLFFPC_ERR_NOT_PSD when a pivot is <= 0. */
int mlffpc_pchol_workspace_bytes(mlffpc_ctx* ctx, int64_t k, int64_t* bytes);
int mlffpc_pchol_build(mlffpc_ctx* ctx, int64_t k, double* Lt, int64_t ld, double* diag,
                       int64_t* index_columns, const int64_t* forced_pivots, float* step_ms_host,
                       void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- preconditioner ---- */
/* Woodbury factor, in place: T = chol_lower(lam I_k + Lt Lt^T)^{-1} Lt   [k, n_local].
 * Replaces iterative_cholesky.py:141-143.  W is a k*k device scratch. */
int mlffpc_woodbury_factor(mlffpc_ctx* ctx, double* Lt, int64_t k, int64_t ld, double lam, double* W,
                           void* stream);
/* The same inverse (L L^T + lam I)^{-1} in an orthonormal basis, in place: Lt -> Qt [k, n_local] with
 * Qt Qt^T = I (CholeskyQR2 of L, two SYRK + POTRF + TRSM passes) and Mk[k,k] = (Qt L L^T Qt^T + lam I)^{-1}.
 * The Woodbury form above subtracts T^T T r from r with lam = 1e-10 in the denominator: rounding in the k x k
 * Gram of 1e5-long columns (~1e-13 |W|) is then a visible perturbation of the operator and CG stalls on a
 * plateau (DESIGN.md "Woodbury accuracy"); this form evaluates the range part without cancellation.  Same
 * operator, same traffic per apply.  W1, W2: k*k device scratch each. */
int mlffpc_orthonormal_factor(mlffpc_ctx* ctx, double* Lt, int64_t k, int64_t ld, double lam, double* Mk,
                              double* W1, double* W2, void* stream);
/* E[k,k] = Q Q^T - I for Q[k, n_cols] (this rank's columns; summed over ranks), accumulated in extended precision on
 * the DMMA pipe: every 16-column k-tile product is folded into an unevaluated (hi, lo) sum with an error-free
 * TwoSum (csrc/gramdd.cu), so entries of size 1e-15 come out to ~1e-18 -- a plain fp64 Gram's own rounding is as
 * large as E.  mlffpc_syrk_rows uses the same accumulation by default (option "gram_mode" = 0 selects the single
 * running sum of round 1). */
int mlffpc_gram_defect(mlffpc_ctx* ctx, const double* Q, int64_t k, int64_t n_cols, int64_t ld, double* E, void* stream);
/* mlffpc_orthonormal_factor followed by E = Qt Qt^T - I (mlffpc_gram_defect): the "projected" form.  With E the
 * apply uses the exact projector onto range(L) to first order,
 *     z = sign * ((r - Qt^T (w - E w)) / lam + Qt^T Mk w),   w = Qt r,
 * in the same two passes over the factor as the reference formula.  Without E the defect of Qt Qt^T - I (~1e-15) is
 * amplified by 1 / lam = 1e10 and splits the unit eigenvalue cluster of the preconditioned operator: the residual
 * parks on a plateau (cfg2: 1783 iterations with the reference formula, 900 with this form). */
int mlffpc_projected_factor(mlffpc_ctx* ctx, double* Lt, int64_t k, int64_t ld, double lam, double* Mk, double* E,
                            double* W1, double* W2, void* stream);
/* z = sign * (r - T^T (T r)) / lam  on the local rows (Mk == NULL), or with Mk from
 * mlffpc_orthonormal_factor  z = sign * ((r - T^T (T r)) / lam + T^T Mk (T r)), or with Mk and E from
 * mlffpc_projected_factor the projected form above.
 * u is a device scratch of 4 k + 8 doubles.
 * sign = +1: iterative_cholesky.py:145-148; sign = -1: the Nystroem operator
 * (iterative_solver.py:315-318) and _init_precon_operator_sb (:376-379). */
int mlffpc_precon_apply(mlffpc_ctx* ctx, const double* T, int64_t k, int64_t ld, double lam, double sign,
                        const double* r, double* z, double* u, const double* Mk, const double* E, void* stream);

/* ---------------------------------------------------------------- PCG ---- */
/* Preconditioned CG on A x = b, A = -K + lam I, with scipy-1.7.3 legacy stopping semantics
 * (||r|| <= tol ||b||, residual recomputed once on first hit; call site iterative_solver.py:995-1005).
 *   K_local: explicit local rows [n_local, n] (ld_k) or NULL for the matrix-free operator
 *   T: preconditioner factor [k, n_local] or NULL (identity); precon_sign, Mk, E (may be NULL) as in
 *      mlffpc_precon_apply
 *   b, x: local rows (x in: initial guess, out: solution)
 *   out_host[8] (host doubles): iterations, final ||r||, info (0 converged), ||b||,
 *       summed operator ms, operator calls, summed preconditioner ms (CUDA events), reserved
 *   resid_hist_host: NULL or host double[maxiter+1] receiving ||r|| per iteration
 *   resume_iters: 0, or the iteration count a previous call with the SAME workspace and x stopped at (its maxiter):
 *       the run continues with the saved r, p, rho -- one uninterrupted recurrence cut into checkpoint segments
 *       (the reference writes an unconverged model every ~2 minutes from scipy's callback, iterative_solver.py:919-954)
 * The stopping test runs on the device; the host reads the loop state one batch of iterations late and never waits
 * for a scalar inside the loop (csrc/pcg.cu).  Iteration count, x and r are those of the legacy loop.
 * Workspace: mlffpc_pcg_workspace_bytes. */
int mlffpc_pcg_workspace_bytes(mlffpc_ctx* ctx, int64_t k, int matrix_free, int64_t* bytes);
int mlffpc_pcg(mlffpc_ctx* ctx, const double* K_local, int64_t ld_k, double lam, const double* T,
               int64_t k, int64_t ld_t, double precon_sign, const double* Mk, const double* E, const double* b,
               double* x, double tol, int64_t maxiter, int64_t resume_iters, double* out_host, double* resid_hist_host,
               void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- small vector helpers ---- */
/* out[i] = X[rows[i], :] gathered rows etc. are done with torch indexing on the host side; the only
 * vector helpers exported are the fused CG kernels, for tests:                                   */
int mlffpc_dot(mlffpc_ctx* ctx, const double* a, const double* b, int64_t n, double* out_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MLFFPC_H */
