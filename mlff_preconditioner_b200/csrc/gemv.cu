// HBM-bound fp64 matrix-vector kernels over row-major matrices.
//  * gemv_rows:  y = alpha * K x + shift * x_local     (K streamed once, 8*rows*cols bytes)
//  * tgemv_cols: out[c] = sum_m T[m, c] w[m]            (column combination, coalesced across c)
#include "common.cuh"

namespace mlffpc {

constexpr int GEMV_ROWS = 8;      // rows per CTA: x is re-read from L2 once per 8 rows of K
constexpr int GEMV_THREADS = 256;

__device__ __forceinline__ double2 ld_stream2(const double* p) {
    // streaming (evict-first) 128-bit load: K is touched exactly once per matvec
    return __ldcs(reinterpret_cast<const double2*>(p));
}

template <bool VEC2>
__global__ void __launch_bounds__(GEMV_THREADS)
gemv_rows_kernel(const double* __restrict__ K, int64_t n_rows, int64_t n_cols, int64_t ld,
                 const double* __restrict__ x, double* __restrict__ y, double alpha, double shift,
                 int64_t x_off) {
    __shared__ double red[GEMV_ROWS][GEMV_THREADS / 32];
    const int64_t r0 = (int64_t)blockIdx.x * GEMV_ROWS;
    const int tid = threadIdx.x;
    double acc[GEMV_ROWS];
#pragma unroll
    for (int r = 0; r < GEMV_ROWS; ++r) acc[r] = 0.0;

    const double* rowp[GEMV_ROWS];
#pragma unroll
    for (int r = 0; r < GEMV_ROWS; ++r) {
        const int64_t rr = (r0 + r < n_rows) ? (r0 + r) : (n_rows - 1);  // clamp: tail rows recompute the last row
        rowp[r] = K + rr * ld;
    }

    if (VEC2) {
        const int64_t nv = n_cols >> 1;
        for (int64_t c = tid; c < nv; c += GEMV_THREADS) {
            const double2 xv = __ldg(reinterpret_cast<const double2*>(x) + c);
            double2 kv[GEMV_ROWS];
#pragma unroll
            for (int r = 0; r < GEMV_ROWS; ++r) kv[r] = ld_stream2(rowp[r] + 2 * c);
#pragma unroll
            for (int r = 0; r < GEMV_ROWS; ++r) acc[r] = fma(kv[r].y, xv.y, fma(kv[r].x, xv.x, acc[r]));
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

template <int MSPLIT>
__global__ void __launch_bounds__(TGEMV_THREADS)
tgemv_cols_kernel(const double* __restrict__ T, int64_t k, int64_t n_cols, int64_t ld,
                  const double* __restrict__ w, double* __restrict__ out, int post,
                  const double* __restrict__ r, double sign_over_lam) {
    constexpr int COLS = TGEMV_THREADS / MSPLIT;
    __shared__ double red[MSPLIT][COLS];
    const int tc = threadIdx.x % COLS, ts = threadIdx.x / COLS;
    const int64_t c = (int64_t)blockIdx.x * COLS + tc;
    double acc = 0.0;
    if (c < n_cols) {
        const double* Tp = T + c;
        int64_t m = ts;
        // 8 independent loads in flight per thread
        for (; m + 7 * MSPLIT < k; m += 8 * MSPLIT) {
            double t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) t[u] = __ldcs(Tp + (m + u * MSPLIT) * ld);
#pragma unroll
            for (int u = 0; u < 8; ++u) acc = fma(t[u], __ldg(w + m + u * MSPLIT), acc);
        }
        for (; m < k; m += MSPLIT) acc = fma(__ldcs(Tp + m * ld), __ldg(w + m), acc);
    }
    if (MSPLIT > 1) {
        red[ts][tc] = acc;
        __syncthreads();
        if (ts != 0) return;
#pragma unroll
        for (int s = 1; s < MSPLIT; ++s) acc += red[s][tc];
    }
    if (c < n_cols) {
        if (post == 1) acc = sign_over_lam * (r[c] - acc);
        out[c] = acc;
    }
}

// ---- symmetric matvec: read only the lower triangle (by row strips), use every entry twice ----------
// Strip s = rows [s*TR, (s+1)*TR).  For the columns left of the strip's diagonal block each loaded K[r,c]
// contributes to y[r] (row sum, kept in registers) and to y[c] (column sum over the strip's TR rows, written
// to ws[s, c] -- plain coalesced stores, no atomics, so the result is deterministic).  The TR x TR diagonal
// block is applied one-sided.  A second pass adds the column partials.  HBM traffic ~ 4 n^2 + 8 n^2/TR bytes.
constexpr int SYMV_THREADS = 256;

template <int TR>
__global__ void __launch_bounds__(SYMV_THREADS, (TR <= 32 ? 2 : 1))
symv_strip_kernel(const double* __restrict__ K, int64_t n, int64_t ld, const double* __restrict__ x,
                  double* __restrict__ y1, double* __restrict__ ws, int64_t ld_ws) {
    __shared__ double xs[TR];
    __shared__ double red[TR][SYMV_THREADS / 32];
    __shared__ double dsum[TR];
    const int tid = threadIdx.x;
    const int64_t s = (int64_t)gridDim.x - 1 - blockIdx.x;  // longest strips first
    const int64_t r0 = s * TR;
    const int nr = (int)((n - r0 < TR) ? (n - r0) : TR);
    if (tid < TR) xs[tid] = (tid < nr) ? x[r0 + tid] : 0.0;
    __syncthreads();

    double acc[TR];
#pragma unroll
    for (int i = 0; i < TR; ++i) acc[i] = 0.0;
    const double* base = K + r0 * ld;
    const int64_t last = (int64_t)(nr - 1) * ld;  // rows past the end re-read the last valid row (xs = 0 there)

    const int64_t nv = r0 >> 1;  // r0 is a multiple of TR (even): column pairs never straddle the diagonal block
    for (int64_t c2 = tid; c2 < nv; c2 += SYMV_THREADS) {
        const double2 xv = __ldg(reinterpret_cast<const double2*>(x) + c2);
        double2 cacc = make_double2(0.0, 0.0);
#pragma unroll
        for (int b = 0; b < TR / 8; ++b) {
            double2 kv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int64_t off = (int64_t)(b * 8 + i) * ld;
                kv[i] = ld_stream2(base + (off <= last ? off : last) + 2 * c2);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const double xr = xs[b * 8 + i];
                acc[b * 8 + i] = fma(kv[i].y, xv.y, fma(kv[i].x, xv.x, acc[b * 8 + i]));
                cacc.x = fma(kv[i].x, xr, cacc.x);
                cacc.y = fma(kv[i].y, xr, cacc.y);
            }
        }
        *reinterpret_cast<double2*>(ws + s * ld_ws + 2 * c2) = cacc;
    }

    // diagonal block, one-sided: 8 lanes per row
    {
        const int cp = tid & 7;
        for (int r = tid >> 3; r < TR; r += SYMV_THREADS / 8) {
            double v = 0.0;
            if (r < nr)
                for (int c = cp; c < nr; c += 8) v = fma(base[(int64_t)r * ld + r0 + c], xs[c], v);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            if (cp == 0) dsum[r] = v;
        }
    }

    const int lane = tid & 31, w = tid >> 5;
#pragma unroll
    for (int i = 0; i < TR; ++i) {
        const double v = warp_sum(acc[i]);
        if (lane == 0) red[i][w] = v;
    }
    __syncthreads();
    if (tid < nr) {
        double v = dsum[tid];
#pragma unroll
        for (int i = 0; i < SYMV_THREADS / 32; ++i) v += red[tid][i];
        y1[r0 + tid] = v;
    }
}

// y[c] = alpha * (y1[c] + sum_{strips below c's strip} ws[s, c]) + shift * x[c]
template <int TR>
__global__ void symv_reduce_kernel(const double* __restrict__ y1, const double* __restrict__ ws, int64_t ld_ws,
                                   int64_t n, int64_t nstrips, const double* __restrict__ x,
                                   double* __restrict__ y, double alpha, double shift) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    double acc = y1[c];
    int64_t s = c / TR + 1;
    const double* p = ws + c;
    for (; s + 7 < nstrips; s += 8) {
        double t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) t[u] = __ldcs(p + (s + u) * ld_ws);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += t[u];
    }
    for (; s < nstrips; ++s) acc += __ldcs(p + s * ld_ws);
    acc *= alpha;
    if (shift != 0.0) acc = fma(shift, x[c], acc);
    y[c] = acc;
}

constexpr int SYMV_TR = 32;

int64_t symv_ws_bytes(int64_t n) {
    const int64_t nstrips = (n + SYMV_TR - 1) / SYMV_TR;
    const int64_t ld_ws = (n + 1) & ~(int64_t)1;
    return (nstrips * ld_ws + n + 64) * 8 + 512;
}

// y = alpha * K x + shift * x for symmetric K (only the lower triangle by SYMV_TR-row strips is read)
int launch_symv(const double* K, int64_t n, int64_t ld, const double* x, double* y, double alpha, double shift,
                void* workspace, cudaStream_t s) {
    MLFFPC_REQUIRE(ld % 2 == 0 && (((uintptr_t)K | (uintptr_t)x) % 16 == 0),
                   "symv: K and x must be 16-byte aligned with an even leading dimension");
    const int64_t nstrips = (n + SYMV_TR - 1) / SYMV_TR;
    const int64_t ld_ws = (n + 1) & ~(int64_t)1;
    double* wsd = (double*)(((uintptr_t)workspace + 255) / 256 * 256);
    double* y1 = wsd;
    double* ws = wsd + ((n + 31) / 32 * 32);
    symv_strip_kernel<SYMV_TR><<<(unsigned)nstrips, SYMV_THREADS, 0, s>>>(K, n, ld, x, y1, ws, ld_ws);
    MLFFPC_LAUNCH_CHECK();
    symv_reduce_kernel<SYMV_TR><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(y1, ws, ld_ws, n, nstrips, x, y, alpha, shift);
    MLFFPC_LAUNCH_CHECK();
    return MLFFPC_OK;
}

int launch_gemv_rows(const double* K, int64_t n_rows, int64_t n_cols, int64_t ld, const double* x,
                     double* y, double alpha, double shift, int64_t x_off, cudaStream_t s) {
    if (n_rows <= 0) return MLFFPC_OK;
    const unsigned grid = (unsigned)((n_rows + GEMV_ROWS - 1) / GEMV_ROWS);
    const bool vec2 = (ld % 2 == 0) && (((uintptr_t)K | (uintptr_t)x) % 16 == 0);
    if (vec2)
        gemv_rows_kernel<true><<<grid, GEMV_THREADS, 0, s>>>(K, n_rows, n_cols, ld, x, y, alpha, shift, x_off);
    else
        gemv_rows_kernel<false><<<grid, GEMV_THREADS, 0, s>>>(K, n_rows, n_cols, ld, x, y, alpha, shift, x_off);
    MLFFPC_LAUNCH_CHECK();
    return MLFFPC_OK;
}

int launch_tgemv_cols(const double* T, int64_t k, int64_t n_cols, int64_t ld, const double* w,
                      double* out, int post, const double* r, double sign_over_lam, int num_sms,
                      cudaStream_t s) {
    if (n_cols <= 0) return MLFFPC_OK;
    // enough threads to keep HBM busy: aim for >= 2 full waves of 256-thread CTAs
    const int64_t want = (int64_t)num_sms * 2048;
    if (n_cols >= want || k < 64) {
        tgemv_cols_kernel<1><<<(unsigned)((n_cols + 255) / 256), TGEMV_THREADS, 0, s>>>(T, k, n_cols, ld, w, out, post, r, sign_over_lam);
    } else if (n_cols * 4 >= want || k < 256) {
        tgemv_cols_kernel<4><<<(unsigned)((n_cols + 63) / 64), TGEMV_THREADS, 0, s>>>(T, k, n_cols, ld, w, out, post, r, sign_over_lam);
    } else {
        tgemv_cols_kernel<8><<<(unsigned)((n_cols + 31) / 32), TGEMV_THREADS, 0, s>>>(T, k, n_cols, ld, w, out, post, r, sign_over_lam);
    }
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

int mlffpc_symv_workspace_bytes(int64_t n, int64_t* bytes) {
    MLFFPC_REQUIRE(bytes && n > 0, "symv_workspace_bytes: bad argument");
    *bytes = symv_ws_bytes(n);
    return MLFFPC_OK;
}

int mlffpc_symv(mlffpc_ctx* ctx, const double* K, int64_t n, int64_t ld, const double* x, double* y,
                double alpha, double shift, void* workspace, int64_t workspace_bytes, void* stream) {
    MLFFPC_REQUIRE(ctx && K && x && y && workspace, "symv: NULL argument");
    MLFFPC_REQUIRE(n > 0 && ld >= n, "symv: bad dimensions");
    MLFFPC_REQUIRE(workspace_bytes >= symv_ws_bytes(n), "symv: workspace too small");
    return launch_symv(K, n, ld, x, y, alpha, shift, workspace, (cudaStream_t)stream);
}

int mlffpc_set_option(mlffpc_ctx* ctx, const char* name, int64_t value) {
    MLFFPC_REQUIRE(ctx && name, "set_option: NULL argument");
    if (std::string(name) == "symmetric_gemv") {
        ctx->use_symv = value != 0;
        return MLFFPC_OK;
    }
    set_error("set_option: unknown option '%s'", name);
    return MLFFPC_ERR_INVALID;
}

}  // extern "C"
