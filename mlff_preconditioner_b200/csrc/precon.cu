// Low-rank (Woodbury / Nystroem) preconditioner: factorisation and apply.
//   factor (reference iterative_cholesky.py:141-143):  W = lam I + Lt Lt^T, L2 = chol(W), T = L2^{-1} Lt
//   apply  (iterative_cholesky.py:145-148, iterative_solver.py:315-318):  z = sign (r - T^T (T r)) / lam
// T is [k, n_local] row-major: "T r" streams k long rows, "T^T u" combines columns -- both coalesced,
// 16 k n_local bytes of HBM traffic per apply.
#include "common.cuh"

namespace mlffpc {

__global__ void add_inplace_kernel(double* __restrict__ a, const double* __restrict__ b, int64_t n) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) a[t] += b[t];
}

// scratch of the twice-projected apply: owned by the context, (re)allocated when the shapes grow
int ensure_reorth_scratch(mlffpc_ctx* ctx, int64_t k) {
    if (!ctx->precon_reorth) return MLFFPC_OK;
    const int64_t need = ((ctx->n_local() + 31) / 32 * 32) + k + 64;
    if (ctx->reorth_scratch_len >= need) return MLFFPC_OK;
    if (ctx->reorth_scratch) cudaFree(ctx->reorth_scratch);
    ctx->reorth_scratch = nullptr;
    ctx->reorth_scratch_len = 0;
    MLFFPC_CUDA(cudaMalloc((void**)&ctx->reorth_scratch, (size_t)need * sizeof(double)));
    ctx->reorth_scratch_len = need;
    return MLFFPC_OK;
}

// w = T r on the local columns: the TMA row-strip kernel when the shapes allow it, else the register-staged GEMV
static int factor_times_vec(mlffpc_ctx* ctx, const double* T, int64_t k, int64_t nl, int64_t ld, const double* r, double* w,
                            cudaStream_t s) {
    if (ctx->tma_rows && rows_tma_usable(T, ld, k, nl)) {
        const int64_t need = rows_tma_ws_doubles(k, nl, ctx->num_sms);
        if (ctx->rows_ws_len < need) {
            if (ctx->rows_ws) cudaFree(ctx->rows_ws);
            ctx->rows_ws = nullptr;
            ctx->rows_ws_len = 0;
            MLFFPC_CUDA(cudaMalloc((void**)&ctx->rows_ws, (size_t)need * sizeof(double)));
            ctx->rows_ws_len = need;
        }
        return rows_gemv_tma(ctx, T, k, nl, ld, r, w, 1.0, ctx->rows_ws, s);
    }
    return launch_gemv_rows(T, k, nl, ld, r, w, 1.0, 0.0, 0, s, false);
}

// sum of a k-vector over the ranks: one small peer-memory kernel (push, flag, wait, rank-ordered sum) when the peer
// buffers are mapped, else ncclAllReduce
static int kvec_allreduce(mlffpc_ctx* ctx, double* w, int64_t k, cudaStream_t s) {
    if (ctx->comm.world <= 1) return MLFFPC_OK;
    if (peer_on(ctx) && ctx->peer_kvec && k <= peer_kmax(ctx)) return peer_allreduce_kvec(ctx, w, k, s);
    return comm_allreduce_sum(ctx->comm, w, (size_t)k, s);
}

// u: device scratch of 4 k + 8 doubles.  Mk == NULL: Woodbury form z = sign (r - T^T T r) / lam.
// Mk != NULL: T holds an orthonormal basis Q^T of range(L) and Mk = (Q^T L L^T Q + lam I)^{-1}:
//   z = sign ( (r - Q (Q^T r)) / lam + Q Mk (Q^T r) ).
int precon_apply(mlffpc_ctx* ctx, const double* T, int64_t k, int64_t ld, double lam, double sign,
                 const double* r, double* z, double* u, cudaStream_t s, const double* Mk, const double* E) {
    const int64_t nl = ctx->n_local();
    if (Mk && E) {
        // Projected form, two passes over the factor.  T = Qt has orthonormal rows up to the defect
        // E = Qt Qt^T - I (~1e-15), which the complement term (r - Qt^T Qt r) / lam would amplify by 1 / lam = 1e10.
        // With w = Qt r the exact projector onto range(Qt^T) is Qt^T (I + E)^{-1} Qt = Qt^T (I - E) Qt + O(E^2):
        //   z = sign ( (r - Qt^T (w - E w)) / lam + Qt^T Mk w ).
        // E comes from the extended-precision Gram of gramdd.cu; both k x k products are replicated on every rank.
        const int64_t ko = (k + 3) & ~(int64_t)1;
        double* w = u;
        double* g1 = u + ko;       // w - E w
        double* g2 = u + 2 * ko;   // Mk w
        MLFFPC_TRY(factor_times_vec(ctx, T, k, nl, ld, r, w, s));
        MLFFPC_TRY(kvec_allreduce(ctx, w, k, s));
        MLFFPC_TRY(launch_gemv_rows(E, k, k, k, w, g1, -1.0, 1.0, 0, s, false));
        MLFFPC_TRY(launch_gemv_rows(Mk, k, k, k, w, g2, 1.0, 0.0, 0, s, false));
        return launch_tgemv_cols(T, k, nl, ld, g1, z, 1, r, sign / lam, ctx->num_sms, s, false, ctx->tgemv_msplit, g2, sign);
    }
    // u = T r  (local part), summed over ranks
    const bool comp = ctx->precon_accuracy == 1;
    if (ctx->precon_accuracy == 2 && nl >= 4) {
        // diagnostics: u as the sum of two half-length products (the summation order of a 2-rank run)
        const int64_t h = (nl / 2) & ~(int64_t)1;
        double* u2 = u + k + 2;
        MLFFPC_TRY(launch_gemv_rows(T, k, h, ld, r, u, 1.0, 0.0, 0, s, false));
        MLFFPC_TRY(launch_gemv_rows(T + h, k, nl - h, ld, r + h, u2, 1.0, 0.0, 0, s, false));
        add_inplace_kernel<<<(unsigned)((k + 255) / 256), 256, 0, s>>>(u, u2, k);
        MLFFPC_LAUNCH_CHECK();
    } else if (comp) {
        MLFFPC_TRY(launch_gemv_rows(T, k, nl, ld, r, u, 1.0, 0.0, 0, s, comp));
    } else {
        MLFFPC_TRY(factor_times_vec(ctx, T, k, nl, ld, r, u, s));
    }
    MLFFPC_TRY(kvec_allreduce(ctx, u, k, s));
    if (Mk && ctx->precon_reorth && ctx->reorth_scratch) {
        // Orthonormal form with the complement projected twice ("twice is enough"): Qt Qt^T = I + E with
        // |E| ~ 1e-16 sqrt(n), and (I - Qt^T Qt) a / lam leaks E / lam ~ 1e-4 of a range vector back into the
        // range.  rp = (I - Qt^T Qt)^2 r removes the leak; z = rp / lam + Qt^T Mk (w + w2).  Same kernels, two
        // more passes over the factor.  (CPU experiment: profiles/r01v_precon_forms_cpu.txt.)
        double* rp = ctx->reorth_scratch;       // n_local doubles
        double* w2 = u + k + 2;
        MLFFPC_TRY(launch_tgemv_cols(T, k, nl, ld, u, rp, 1, r, 1.0, ctx->num_sms, s, false, ctx->tgemv_msplit));  // rp = r - Qt^T w
        MLFFPC_TRY(launch_gemv_rows(T, k, nl, ld, rp, w2, 1.0, 0.0, 0, s, false));                                   // w2 = Qt rp
        MLFFPC_TRY(comm_allreduce_sum(ctx->comm, w2, (size_t)k, s));
        add_inplace_kernel<<<(unsigned)((k + 255) / 256), 256, 0, s>>>(u, w2, k);                                   // u = w + w2
        MLFFPC_LAUNCH_CHECK();
        double* mu = ctx->reorth_scratch + ((nl + 31) / 32 * 32);  // k doubles
        MLFFPC_TRY(launch_gemv_rows(Mk, k, k, k, u, mu, 1.0, 0.0, 0, s, false));                                     // mu = Mk (w + w2)
        // z = sign ((rp - Qt^T w2) / lam + Qt^T mu)
        return launch_tgemv_cols(T, k, nl, ld, w2, z, 1, rp, sign / lam, ctx->num_sms, s, false, ctx->tgemv_msplit, mu, sign);
    }
    if (Mk) {
        double* u2 = u + k + 2;
        MLFFPC_TRY(launch_gemv_rows(Mk, k, k, k, u, u2, 1.0, 0.0, 0, s, false));  // replicated k x k product
        return launch_tgemv_cols(T, k, nl, ld, u, z, 1, r, sign / lam, ctx->num_sms, s, false, ctx->tgemv_msplit, u2, sign);
    }
    // z = sign (r - T^T u) / lam
    MLFFPC_TRY(launch_tgemv_cols(T, k, nl, ld, u, z, 1, r, sign / lam, ctx->num_sms, s, comp, ctx->tgemv_msplit));
    return MLFFPC_OK;
}

__global__ void transpose_kernel(const double* __restrict__ A, double* __restrict__ At, int64_t m) {
    __shared__ double tile[32][33];
    const int64_t bx = (int64_t)blockIdx.x * 32, by = (int64_t)blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y)
        if (by + i < m && bx + threadIdx.x < m) tile[i][threadIdx.x] = A[(by + i) * m + bx + threadIdx.x];
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y)
        if (bx + i < m && by + threadIdx.x < m) At[(bx + i) * m + by + threadIdx.x] = tile[threadIdx.x][i];
}

__global__ void set_identity_kernel(double* __restrict__ A, int64_t m) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r = blockIdx.y;
    if (c < m) A[r * m + c] = (c == r) ? 1.0 : 0.0;
}

// lower -> full symmetric, diagonal += shift (packed k x k)
__global__ void mirror_shift_kernel(double* W, int64_t m, double shift) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r = blockIdx.y;
    if (c >= m) return;
    if (c > r) W[r * m + c] = W[c * m + r];
    else if (c == r) W[r * m + c] += shift;
}

__global__ void scale_copy_kernel(const double* __restrict__ r, double* __restrict__ z, int64_t n, double a) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) z[t] = a * r[t];
}

}  // namespace mlffpc

using namespace mlffpc;

extern "C" {

int mlffpc_woodbury_factor(mlffpc_ctx* ctx, double* Lt, int64_t k, int64_t ld, double lam, double* W,
                           void* stream) {
    MLFFPC_REQUIRE(ctx && ctx->M > 0, "woodbury_factor: geometry not set");
    MLFFPC_REQUIRE(Lt && W && k > 0 && ld >= ctx->n_local(), "woodbury_factor: bad argument");
    ProfWindow pw = prof_window("woodbury");
    pw.step(pw.first);
    PhaseTimer pt((cudaStream_t)stream);
    int st = mlffpc_syrk_rows(ctx, Lt, k, ctx->n_local(), ld, lam, W, k, stream);
    pt.lap("woodbury: gram");
    int info = 0;
    if (st == MLFFPC_OK) st = mlffpc_potrf_lower(ctx, W, k, k, &info, stream);
    pt.lap("woodbury: potrf");
    if (st == MLFFPC_OK && info != 0) {
        set_error("%d-th leading minor of the array is not positive definite", info);
        st = MLFFPC_ERR_LINALG;
    }
    if (st == MLFFPC_OK) st = mlffpc_trsm_rows(ctx, W, k, k, Lt, ctx->n_local(), ld, stream);
    pt.lap("woodbury: trsm");
    pw.end();
    return st;
}

int mlffpc_orthonormal_factor(mlffpc_ctx* ctx, double* Lt, int64_t k, int64_t ld, double lam, double* Mk,
                              double* W1, double* W2, void* stream) {
    MLFFPC_REQUIRE(ctx && ctx->M > 0, "orthonormal_factor: geometry not set");
    MLFFPC_REQUIRE(Lt && Mk && W1 && W2 && k > 0 && ld >= ctx->n_local() && lam > 0.0, "orthonormal_factor: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t nl = ctx->n_local();
    const dim3 g2((unsigned)((k + 255) / 256), (unsigned)k);
    const dim3 gt((unsigned)((k + 31) / 32), (unsigned)((k + 31) / 32)), bt(32, 8);
    auto chol = [&](double* W) -> int {
        int info = 0;
        MLFFPC_TRY(mlffpc_potrf_lower(ctx, W, k, k, &info, stream));
        if (info != 0) {
            set_error("%d-th leading minor of the array is not positive definite", info);
            return MLFFPC_ERR_LINALG;
        }
        return MLFFPC_OK;
    };
    ProfWindow pw = prof_window("woodbury");
    pw.step(pw.first);
    PhaseTimer pt(s);
    // one TRSM order for the whole factorisation, the k x k solve included (dense.cu, mlffpc_trsm_rows)
    struct OrderGuard {
        mlffpc_ctx* c;
        OrderGuard(mlffpc_ctx* cc, int64_t cols) : c(cc) { c->trsm_order = cols < 32768 ? 1 : 0; }
        ~OrderGuard() { c->trsm_order = -1; }
    } order_guard(ctx, nl);
    // CholeskyQR2 of L (rows of Lt):  Lt = C1 C2 Qt with Qt Qt^T = I to working precision (cond(L) << 1e8)
    MLFFPC_TRY(mlffpc_syrk_rows(ctx, Lt, k, nl, ld, 0.0, W1, k, stream)); pt.lap("orthonormal: gram 1");
    MLFFPC_TRY(chol(W1)); pt.lap("orthonormal: potrf 1");
    MLFFPC_TRY(mlffpc_trsm_rows(ctx, W1, k, k, Lt, nl, ld, stream)); pt.lap("orthonormal: trsm 1");
    MLFFPC_TRY(mlffpc_syrk_rows(ctx, Lt, k, nl, ld, 0.0, W2, k, stream)); pt.lap("orthonormal: gram 2");
    MLFFPC_TRY(chol(W2)); pt.lap("orthonormal: potrf 2");
    MLFFPC_TRY(mlffpc_trsm_rows(ctx, W2, k, k, Lt, nl, ld, stream)); pt.lap("orthonormal: trsm 2");
    // B = C1 C2 (lower triangular): L L^T = Q (B^T B) Q^T.   S = B^T B + lam I
    MLFFPC_TRY(dgemm(false, k, k, k, 1.0, W1, k, W2, k, 0.0, Mk, k, false, s));
    transpose_kernel<<<gt, bt, 0, s>>>(Mk, W2, k);                                   // W2 = B^T
    MLFFPC_LAUNCH_CHECK();
    MLFFPC_TRY(dgemm(true, k, k, k, 1.0, W2, k, W2, k, 0.0, W1, k, true, s));         // W1 = B^T B (lower tiles)
    mirror_shift_kernel<<<g2, 256, 0, s>>>(W1, k, lam);
    MLFFPC_LAUNCH_CHECK();
    // Mk = S^{-1} = Y^T Y with Y = chol(S)^{-1}
    MLFFPC_TRY(chol(W1));
    set_identity_kernel<<<g2, 256, 0, s>>>(W2, k);
    MLFFPC_LAUNCH_CHECK();
    MLFFPC_TRY(mlffpc_trsm_rows(ctx, W1, k, k, W2, k, k, stream));                    // W2 = Y
    transpose_kernel<<<gt, bt, 0, s>>>(W2, W1, k);                                   // W1 = Y^T
    MLFFPC_LAUNCH_CHECK();
    MLFFPC_TRY(dgemm(true, k, k, k, 1.0, W1, k, W1, k, 0.0, Mk, k, true, s));         // Mk = Y^T Y (lower tiles)
    mirror_shift_kernel<<<g2, 256, 0, s>>>(Mk, k, 0.0);
    MLFFPC_LAUNCH_CHECK();
    pt.lap("orthonormal: k x k inverse Mk");
    pw.end();
    return MLFFPC_OK;
}

int mlffpc_projected_factor(mlffpc_ctx* ctx, double* Lt, int64_t k, int64_t ld, double lam, double* Mk, double* E,
                            double* W1, double* W2, void* stream) {
    MLFFPC_REQUIRE(E != nullptr, "projected_factor: E is NULL");
    MLFFPC_TRY(mlffpc_orthonormal_factor(ctx, Lt, k, ld, lam, Mk, W1, W2, stream));
    ProfWindow pw = prof_window("woodbury");
    pw.step(pw.first);
    PhaseTimer pt((cudaStream_t)stream);
    const int st = gram_dd(ctx, Lt, k, ctx->n_local(), ld, E, k, 0.0, true, (cudaStream_t)stream, ctx->defect_mode == 2);
    pt.lap("projected: defect E");
    pw.end();
    return st;
}

int mlffpc_precon_apply(mlffpc_ctx* ctx, const double* T, int64_t k, int64_t ld, double lam, double sign,
                        const double* r, double* z, double* u, const double* Mk, const double* E, void* stream) {
    MLFFPC_REQUIRE(ctx && ctx->M > 0, "precon_apply: geometry not set");
    MLFFPC_REQUIRE(r && z && lam > 0.0, "precon_apply: bad argument");
    MLFFPC_REQUIRE(!E || Mk, "precon_apply: the defect matrix E needs the orthonormal-form Mk");
    cudaStream_t s = (cudaStream_t)stream;
    if (k == 0 || T == nullptr) {
        const int64_t nl = ctx->n_local();
        scale_copy_kernel<<<(unsigned)((nl + 255) / 256), 256, 0, s>>>(r, z, nl, sign / lam);
        MLFFPC_LAUNCH_CHECK();
        return MLFFPC_OK;
    }
    MLFFPC_REQUIRE(u && ld >= ctx->n_local(), "precon_apply: bad argument");
    if (Mk) MLFFPC_TRY(ensure_reorth_scratch(ctx, k));
    return precon_apply(ctx, T, k, ld, lam, sign, r, z, u, s, Mk, E);
}

}  // extern "C"
