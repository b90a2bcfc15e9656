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

// u: device scratch of 2 k + 4 doubles
int precon_apply(mlffpc_ctx* ctx, const double* T, int64_t k, int64_t ld, double lam, double sign,
                 const double* r, double* z, double* u, cudaStream_t s) {
    const int64_t nl = ctx->n_local();
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
    } else {
        MLFFPC_TRY(launch_gemv_rows(T, k, nl, ld, r, u, 1.0, 0.0, 0, s, comp));
    }
    MLFFPC_TRY(comm_allreduce_sum(ctx->comm, u, (size_t)k, s));
    // z = sign (r - T^T u) / lam
    MLFFPC_TRY(launch_tgemv_cols(T, k, nl, ld, u, z, 1, r, sign / lam, ctx->num_sms, s, comp));
    return MLFFPC_OK;
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
    int st = mlffpc_syrk_rows(ctx, Lt, k, ctx->n_local(), ld, lam, W, k, stream);
    int info = 0;
    if (st == MLFFPC_OK) st = mlffpc_potrf_lower(ctx, W, k, k, &info, stream);
    if (st == MLFFPC_OK && info != 0) {
        set_error("%d-th leading minor of the array is not positive definite", info);
        st = MLFFPC_ERR_LINALG;
    }
    if (st == MLFFPC_OK) st = mlffpc_trsm_rows(ctx, W, k, k, Lt, ctx->n_local(), ld, stream);
    pw.end();
    return st;
}

int mlffpc_precon_apply(mlffpc_ctx* ctx, const double* T, int64_t k, int64_t ld, double lam, double sign,
                        const double* r, double* z, double* u, void* stream) {
    MLFFPC_REQUIRE(ctx && ctx->M > 0, "precon_apply: geometry not set");
    MLFFPC_REQUIRE(r && z && lam > 0.0, "precon_apply: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    if (k == 0 || T == nullptr) {
        const int64_t nl = ctx->n_local();
        scale_copy_kernel<<<(unsigned)((nl + 255) / 256), 256, 0, s>>>(r, z, nl, sign / lam);
        MLFFPC_LAUNCH_CHECK();
        return MLFFPC_OK;
    }
    MLFFPC_REQUIRE(u && ld >= ctx->n_local(), "precon_apply: bad argument");
    return precon_apply(ctx, T, k, ld, lam, sign, r, z, u, s);
}

}  // extern "C"
