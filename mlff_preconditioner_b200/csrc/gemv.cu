// HBM-bound fp64 matrix-vector kernels over row-major matrices.
//  * gemv_rows:  y = alpha * K x + shift * x_local     (K streamed once, 8*rows*cols bytes)
//  * tgemv_cols: out[c] = sum_m T[m, c] w[m]            (column combination, coalesced across c)
#include "common.cuh"

namespace mlffpc {

// rows per CTA (template parameter): x is re-read from L2 once per GEMV_ROWS rows of K; 8 for the big
// operator, 4 when the matrix has too few rows to fill the GPU with 8-row CTAs (the k x n factor T)
constexpr int GEMV_THREADS = 256;

__device__ __forceinline__ double2 ld_stream2(const double* p) {
    // streaming (evict-first) 128-bit load: K is touched exactly once per matvec
    return __ldcs(reinterpret_cast<const double2*>(p));
}

// Kahan-compensated accumulate: s += a*b with the running error in c
__device__ __forceinline__ void kahan_fma(double a, double b, double& s, double& c) {
    const double y = fma(a, b, -c);
    const double t = s + y;
    c = (t - s) - y;
    s = t;
}

template <bool VEC2, int GEMV_ROWS, bool COMP>
__global__ void __launch_bounds__(GEMV_THREADS)
gemv_rows_kernel(const double* __restrict__ K, int64_t n_rows, int64_t n_cols, int64_t ld,
                 const double* __restrict__ x, double* __restrict__ y, double alpha, double shift,
                 int64_t x_off) {
    __shared__ double red[GEMV_ROWS][GEMV_THREADS / 32];
    const int64_t r0 = (int64_t)blockIdx.x * GEMV_ROWS;
    const int tid = threadIdx.x;
    double acc[GEMV_ROWS], cmp[GEMV_ROWS];
#pragma unroll
    for (int r = 0; r < GEMV_ROWS; ++r) acc[r] = cmp[r] = 0.0;

    const double* rowp[GEMV_ROWS];
#pragma unroll
    for (int r = 0; r < GEMV_ROWS; ++r) {
        const int64_t rr = (r0 + r < n_rows) ? (r0 + r) : (n_rows - 1);  // clamp: tail rows recompute the last row
        rowp[r] = K + rr * ld;
    }

    if (VEC2) {
        const int64_t nv = n_cols >> 1;
        constexpr int UNR = 8 / GEMV_ROWS;  // keep 8 independent 128-bit loads in flight per thread
        int64_t c = tid;
        for (; c + (UNR - 1) * GEMV_THREADS < nv; c += UNR * GEMV_THREADS) {
            double2 xv[UNR], kv[UNR][GEMV_ROWS];
#pragma unroll
            for (int u = 0; u < UNR; ++u) xv[u] = __ldg(reinterpret_cast<const double2*>(x) + c + u * GEMV_THREADS);
#pragma unroll
            for (int u = 0; u < UNR; ++u)
#pragma unroll
                for (int r = 0; r < GEMV_ROWS; ++r) kv[u][r] = ld_stream2(rowp[r] + 2 * (c + u * GEMV_THREADS));
#pragma unroll
            for (int u = 0; u < UNR; ++u)
#pragma unroll
                for (int r = 0; r < GEMV_ROWS; ++r) {
                    if (COMP) {
                        kahan_fma(kv[u][r].x, xv[u].x, acc[r], cmp[r]);
                        kahan_fma(kv[u][r].y, xv[u].y, acc[r], cmp[r]);
                    } else {
                        acc[r] = fma(kv[u][r].y, xv[u].y, fma(kv[u][r].x, xv[u].x, acc[r]));
                    }
                }
        }
        for (; c < nv; c += GEMV_THREADS) {
            const double2 xv = __ldg(reinterpret_cast<const double2*>(x) + c);
#pragma unroll
            for (int r = 0; r < GEMV_ROWS; ++r) {
                const double2 kv = ld_stream2(rowp[r] + 2 * c);
                if (COMP) {
                    kahan_fma(kv.x, xv.x, acc[r], cmp[r]);
                    kahan_fma(kv.y, xv.y, acc[r], cmp[r]);
                } else {
                    acc[r] = fma(kv.y, xv.y, fma(kv.x, xv.x, acc[r]));
                }
            }
        }
        if ((n_cols & 1) && tid == 0) {
            const double xs = x[n_cols - 1];
#pragma unroll
            for (int r = 0; r < GEMV_ROWS; ++r) acc[r] = fma(rowp[r][n_cols - 1], xs, acc[r]);
        }
    } else {
        for (int64_t c = tid; c < n_cols; c += GEMV_THREADS) {
            const double xs = __ldg(x + c);
#pragma unroll
            for (int r = 0; r < GEMV_ROWS; ++r) acc[r] = fma(__ldcs(rowp[r] + c), xs, acc[r]);
        }
    }

    const int lane = tid & 31, w = tid >> 5;
#pragma unroll
    for (int r = 0; r < GEMV_ROWS; ++r) {
        const double v = warp_sum(acc[r]);
        if (lane == 0) red[r][w] = v;
    }
    __syncthreads();
    if (tid < GEMV_ROWS) {
        const int64_t row = r0 + tid;
        if (row < n_rows) {
            double v = 0.0;
#pragma unroll
            for (int i = 0; i < GEMV_THREADS / 32; ++i) v += red[tid][i];
            v *= alpha;
            if (shift != 0.0) v = fma(shift, x[x_off + row], v);
            y[row] = v;
        }
    }
}

// out[c] = post( sum_{m < k} T[m*ld + c] * w[m] ) ; threads over columns, MSPLIT slices of m per CTA.
// POST: 0 = plain store; 1 = precon apply: out = sign*(r - acc)/lam.
constexpr int TGEMV_THREADS = 256;

// MODE 0: plain; 1: Kahan-compensated (diagnostics); 2: two weight vectors in one pass over T,
//   out = sign * ((r - sum T w) / lam + sum T w2)      (orthonormal-basis form of the low-rank inverse)
template <int MSPLIT, int MODE>
__global__ void __launch_bounds__(TGEMV_THREADS)
tgemv_cols_kernel(const double* __restrict__ T, int64_t k, int64_t n_cols, int64_t ld,
                  const double* __restrict__ w, const double* __restrict__ w2, double* __restrict__ out, int post,
                  const double* __restrict__ r, double sign_over_lam, double sign) {
    constexpr int COLS = TGEMV_THREADS / MSPLIT;
    __shared__ double red[MSPLIT][COLS];
    __shared__ double red2[MODE == 2 ? MSPLIT : 1][COLS];
    const int tc = threadIdx.x % COLS, ts = threadIdx.x / COLS;
    const int64_t c = (int64_t)blockIdx.x * COLS + tc;
    double acc = 0.0, cmp = 0.0, acc2 = 0.0;
    if (c < n_cols) {
        const double* Tp = T + c;
        int64_t m = ts;
        // 8 independent loads in flight per thread
        for (; m + 7 * MSPLIT < k; m += 8 * MSPLIT) {
            double t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) t[u] = __ldcs(Tp + (m + u * MSPLIT) * ld);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (MODE == 1) kahan_fma(t[u], __ldg(w + m + u * MSPLIT), acc, cmp);
                else acc = fma(t[u], __ldg(w + m + u * MSPLIT), acc);
                if (MODE == 2) acc2 = fma(t[u], __ldg(w2 + m + u * MSPLIT), acc2);
            }
        }
        for (; m < k; m += MSPLIT) {
            const double t = __ldcs(Tp + m * ld);
            if (MODE == 1) kahan_fma(t, __ldg(w + m), acc, cmp);
            else acc = fma(t, __ldg(w + m), acc);
            if (MODE == 2) acc2 = fma(t, __ldg(w2 + m), acc2);
        }
    }
    if (MSPLIT > 1) {
        red[ts][tc] = acc;
        if (MODE == 2) red2[ts][tc] = acc2;
        __syncthreads();
        if (ts != 0) return;
#pragma unroll
        for (int s = 1; s < MSPLIT; ++s) {
            acc += red[s][tc];
            if (MODE == 2) acc2 += red2[s][tc];
        }
    }
    if (c < n_cols) {
        if (MODE == 2) acc = fma(sign_over_lam, r[c] - acc, sign * acc2);
        else if (post == 1) acc = sign_over_lam * (r[c] - acc);
        out[c] = acc;
    }
}

int launch_gemv_rows(const double* K, int64_t n_rows, int64_t n_cols, int64_t ld, const double* x,
                     double* y, double alpha, double shift, int64_t x_off, cudaStream_t s, bool compensated) {
    if (n_rows <= 0) return MLFFPC_OK;
    const bool vec2 = (ld % 2 == 0) && (((uintptr_t)K | (uintptr_t)x) % 16 == 0);
    const bool few_rows = n_rows < 8 * 148 * 12;  // fewer than ~4 waves of 8-row CTAs
    const int rows = few_rows ? 4 : 8;
    const unsigned grid = (unsigned)((n_rows + rows - 1) / rows);
    if (compensated && vec2 && few_rows)
        gemv_rows_kernel<true, 4, true><<<grid, GEMV_THREADS, 0, s>>>(K, n_rows, n_cols, ld, x, y, alpha, shift, x_off);
    else if (compensated && vec2)
        gemv_rows_kernel<true, 8, true><<<grid, GEMV_THREADS, 0, s>>>(K, n_rows, n_cols, ld, x, y, alpha, shift, x_off);
    else if (vec2 && few_rows)
        gemv_rows_kernel<true, 4, false><<<grid, GEMV_THREADS, 0, s>>>(K, n_rows, n_cols, ld, x, y, alpha, shift, x_off);
    else if (vec2)
        gemv_rows_kernel<true, 8, false><<<grid, GEMV_THREADS, 0, s>>>(K, n_rows, n_cols, ld, x, y, alpha, shift, x_off);
    else if (few_rows)
        gemv_rows_kernel<false, 4, false><<<grid, GEMV_THREADS, 0, s>>>(K, n_rows, n_cols, ld, x, y, alpha, shift, x_off);
    else
        gemv_rows_kernel<false, 8, false><<<grid, GEMV_THREADS, 0, s>>>(K, n_rows, n_cols, ld, x, y, alpha, shift, x_off);
    MLFFPC_LAUNCH_CHECK();
    return MLFFPC_OK;
}

int launch_tgemv_cols(const double* T, int64_t k, int64_t n_cols, int64_t ld, const double* w,
                      double* out, int post, const double* r, double sign_over_lam, int num_sms,
                      cudaStream_t s, bool compensated, int force_msplit, const double* w2, double sign) {
    if (n_cols <= 0) return MLFFPC_OK;
    // enough threads to keep HBM busy: aim for >= 2 full waves of 256-thread CTAs
    const int64_t want = (int64_t)num_sms * 2048;
    const int ms = force_msplit ? force_msplit : ((n_cols >= want || k < 64) ? 1 : (n_cols * 4 >= want || k < 256) ? 4 : 8);
    const int mode = w2 ? 2 : (compensated ? 1 : 0);
#define MLFFPC_TGEMV(MS, MD, COLS)                                                                                  \
    tgemv_cols_kernel<MS, MD><<<(unsigned)((n_cols + (COLS)-1) / (COLS)), TGEMV_THREADS, 0, s>>>(                   \
        T, k, n_cols, ld, w, w2, out, post, r, sign_over_lam, sign)
#define MLFFPC_TGEMV_MODES(MS, COLS)                                 \
    do {                                                             \
        if (mode == 2) MLFFPC_TGEMV(MS, 2, COLS);                    \
        else if (mode == 1) MLFFPC_TGEMV(MS, 1, COLS);               \
        else MLFFPC_TGEMV(MS, 0, COLS);                              \
    } while (0)
    if (ms == 1) MLFFPC_TGEMV_MODES(1, 256);
    else if (ms == 4) MLFFPC_TGEMV_MODES(4, 64);
    else MLFFPC_TGEMV_MODES(8, 32);
#undef MLFFPC_TGEMV_MODES
#undef MLFFPC_TGEMV
    MLFFPC_LAUNCH_CHECK();
    return MLFFPC_OK;
}

}  // namespace mlffpc

using namespace mlffpc;

extern "C" {

int mlffpc_gemv(mlffpc_ctx* ctx, const double* K, int64_t n_rows, int64_t n_cols, int64_t ld,
                const double* x, double* y, double alpha, double shift, int64_t x_off, void* stream) {
    MLFFPC_REQUIRE(ctx && K && x && y, "gemv: NULL argument");
    MLFFPC_REQUIRE(n_rows >= 0 && n_cols > 0 && ld >= n_cols && x_off >= 0, "gemv: bad dimensions");
    MLFFPC_REQUIRE(shift == 0.0 || x_off + n_rows <= n_cols, "gemv: shift term indexes x out of range");
    return launch_gemv_rows(K, n_rows, n_cols, ld, x, y, alpha, shift, x_off, (cudaStream_t)stream);
}

}  // extern "C"
