// fp64 dense building blocks on the DMMA pipe (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4; tcgen05 has
// no fp64 kind on sm_100): row-major GEMM, SYRK over long rows, blocked Cholesky, blocked row-TRSM.
#include <stdlib.h>

#include "common.cuh"

namespace mlffpc {

__device__ __forceinline__ void dmma884(double& c0, double& c1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc, bool pred) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = pred ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(gsrc), "r"(sz));
}
// 16-byte copy of which only the first `src_bytes` (0, 8 or 16) are read; the rest is zero-filled
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ int vec2_bytes(bool row_ok, int64_t idx, int64_t lim) {
    if (!row_ok || idx >= lim) return 0;
    return (idx + 1 < lim) ? 16 : 8;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

// C[m,n] = alpha * A[m,k] * op(B) + beta * C.   BK = 16, 3-stage cp.async pipeline, 256 or 512 threads.
// smem strides are == 4 (mod 16) doubles so the m8n8k4 fragment loads are bank-conflict free.
constexpr int GEMM_BK = 16;
constexpr int GEMM_STAGES = 3;

template <int BM, int BN, bool TRANSB>
struct GemmSmem {
    static constexpr int A_STRIDE = GEMM_BK + 4;                  // As[BM][A_STRIDE]
    static constexpr int B_STRIDE = TRANSB ? (GEMM_BK + 4) : (BN + 4);  // Bs[BN][..] or Bs[BK][..]
    static constexpr int A_ELEMS = BM * A_STRIDE;
    static constexpr int B_ELEMS = TRANSB ? BN * B_STRIDE : GEMM_BK * B_STRIDE;
    static constexpr int STAGE_ELEMS = A_ELEMS + B_ELEMS;
    static constexpr size_t BYTES = (size_t)GEMM_STAGES * STAGE_ELEMS * sizeof(double);
};

template <int BM, int BN, int WARPS_M, int WARPS_N, bool TRANSB, bool VEC2>
__global__ void __launch_bounds__(WARPS_M * WARPS_N * 32)
dgemm_kernel(int64_t m, int64_t n, int64_t k, double alpha, const double* __restrict__ A, int64_t lda,
             const double* __restrict__ B, int64_t ldb, double beta, double* __restrict__ C, int64_t ldc,
             int lower_only, int64_t k_chunk, int64_t c_zstride) {
    // split-K: slice blockIdx.z multiplies columns [z k_chunk, (z+1) k_chunk) of A with the matching part of B
    // and writes its own partial product C + z c_zstride (the caller adds the partials in a fixed order)
    if (gridDim.z > 1) {
        const int64_t kb = (int64_t)blockIdx.z * k_chunk;
        A += kb;
        B += TRANSB ? kb : kb * ldb;
        C += (int64_t)blockIdx.z * c_zstride;
        k = (k - kb < k_chunk) ? (k - kb) : k_chunk;
        if (k < 0) k = 0;
    }
    using SM = GemmSmem<BM, BN, TRANSB>;
    constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;  // warp tile
    constexpr int TM = WM / 8, TN = WN / 8;              // 8x8 mma tiles per warp
    constexpr int NT = WARPS_M * WARPS_N * 32;  // 8 warps (64x32 or 32x64 warp tiles) or 16 warps (32x32)
    extern __shared__ double smem[];

    const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
    if (lower_only && n0 > m0 + BM - 1) return;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp / WARPS_N) * WM, wn = (warp % WARPS_N) * WN;
    const int lr = lane >> 2, lc = lane & 3;

    double acc[TM][TN][2];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int64_t ktiles = (k + GEMM_BK - 1) / GEMM_BK;

    auto load_stage = [&](int stage, int64_t kt) {
        double* As = smem + (size_t)stage * SM::STAGE_ELEMS;
        double* Bs = As + SM::A_ELEMS;
        const int64_t k0 = kt * GEMM_BK;
        constexpr int V = VEC2 ? 2 : 1;
        // A tile: BM x BK, k contiguous
        constexpr int A_ITEMS = BM * GEMM_BK / V;
        for (int t = tid; t < A_ITEMS; t += NT) {
            const int r = t / (GEMM_BK / V), c = (t % (GEMM_BK / V)) * V;
            const bool ok = (m0 + r < m) && (k0 + c < k);
            const double* src = ok ? (A + (m0 + r) * lda + k0 + c) : A;
            if (VEC2) cp_async16(As + r * SM::A_STRIDE + c, src, vec2_bytes(m0 + r < m, k0 + c, k));
            else cp_async8(As + r * SM::A_STRIDE + c, src, ok);
        }
        if (TRANSB) {  // B[n, k], k contiguous -> Bs[BN][BK+4]
            constexpr int B_ITEMS = BN * GEMM_BK / V;
            for (int t = tid; t < B_ITEMS; t += NT) {
                const int r = t / (GEMM_BK / V), c = (t % (GEMM_BK / V)) * V;
                const bool ok = (n0 + r < n) && (k0 + c < k);
                const double* src = ok ? (B + (n0 + r) * ldb + k0 + c) : B;
                if (VEC2) cp_async16(Bs + r * SM::B_STRIDE + c, src, vec2_bytes(n0 + r < n, k0 + c, k));
                else cp_async8(Bs + r * SM::B_STRIDE + c, src, ok);
            }
        } else {  // B[k, n], n contiguous -> Bs[BK][BN+4]
            constexpr int B_ITEMS = GEMM_BK * BN / V;
            for (int t = tid; t < B_ITEMS; t += NT) {
                const int r = t / (BN / V), c = (t % (BN / V)) * V;
                const bool ok = (k0 + r < k) && (n0 + c < n);
                const double* src = ok ? (B + (k0 + r) * ldb + n0 + c) : B;
                if (VEC2) cp_async16(Bs + r * SM::B_STRIDE + c, src, vec2_bytes(k0 + r < k, n0 + c, n));
                else cp_async8(Bs + r * SM::B_STRIDE + c, src, ok);
            }
        }
    };

#pragma unroll
    for (int s = 0; s < GEMM_STAGES - 1; ++s) {
        if (s < ktiles) load_stage(s, s);
        cp_async_commit();
    }

    for (int64_t kt = 0; kt < ktiles; ++kt) {
        cp_async_wait<GEMM_STAGES - 2>();
        __syncthreads();
        {
            const int64_t nk = kt + GEMM_STAGES - 1;
            if (nk < ktiles) load_stage((int)(nk % GEMM_STAGES), nk);
            cp_async_commit();
        }
        const double* As = smem + (size_t)(kt % GEMM_STAGES) * SM::STAGE_ELEMS;
        const double* Bs = As + SM::A_ELEMS;
#pragma unroll
        for (int kk = 0; kk < GEMM_BK; kk += 4) {
            double af[TM], bf[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) af[i] = As[(wm + i * 8 + lr) * SM::A_STRIDE + kk + lc];
#pragma unroll
            for (int j = 0; j < TN; ++j)
                bf[j] = TRANSB ? Bs[(wn + j * 8 + lr) * SM::B_STRIDE + kk + lc]
                               : Bs[(kk + lc) * SM::B_STRIDE + wn + j * 8 + lr];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();

#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int64_t row = m0 + wm + i * 8 + lr;
        if (row >= m) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int64_t col = n0 + wn + j * 8 + 2 * lc;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                if (col + e < n) {
                    double* p = C + row * ldc + col + e;
                    const double v = alpha * acc[i][j][e];
                    *p = (beta == 0.0) ? v : fma(beta, *p, v);
                }
            }
        }
    }
}

template <int BM, int BN, int WARPS_M, int WARPS_N, bool TRANSB, bool VEC2>
static int launch_dgemm(int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda,
                        const double* B, int64_t ldb, double beta, double* C, int64_t ldc,
                        bool lower_only, cudaStream_t s, int nsplit = 1, int64_t c_zstride = 0) {
    using SM = GemmSmem<BM, BN, TRANSB>;
    auto kern = dgemm_kernel<BM, BN, WARPS_M, WARPS_N, TRANSB, VEC2>;
    MLFFPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::BYTES));
    dim3 grid((unsigned)((n + BN - 1) / BN), (unsigned)((m + BM - 1) / BM), (unsigned)nsplit);
    MLFFPC_REQUIRE(grid.y <= 65535, "dgemm: m = %lld too large for this launch shape", (long long)m);
    // chunks are multiples of the k tile and even, so 16-byte staging stays aligned in every slice
    int64_t k_chunk = (k + nsplit - 1) / nsplit;
    k_chunk = (k_chunk + GEMM_BK - 1) / GEMM_BK * GEMM_BK;
    kern<<<grid, WARPS_M * WARPS_N * 32, SM::BYTES, s>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, lower_only ? 1 : 0, k_chunk,
                                               c_zstride);
    MLFFPC_LAUNCH_CHECK();
    return MLFFPC_OK;
}

// Tile shapes.  The square ones serve the big products; the narrow ones keep a product with few output columns
// (the contraction of the matrix-free operator: n = D + 1 = 37 ... 211) from padding n up to 64 / 128.
enum GemmShapeId { GS_128x128_16W, GS_128x128_8W, GS_128x64, GS_64x128, GS_128x112, GS_128x80, GS_256x40 };

static GemmShapeId gemm_pick_shape(int64_t m, int64_t n) {
    static const bool warps16 = [] { const char* e = getenv("MLFFPC_DGEMM16"); return !(e && e[0] == '0'); }();
    static const bool narrow_tiles = [] { const char* e = getenv("MLFFPC_DGEMM_NARROW"); return !(e && e[0] == '0'); }();
    if (narrow_tiles && n <= 40 && m >= 512) return GS_256x40;
    if (n <= 64) return GS_128x64;
    if (m <= 64) return GS_64x128;  // few rows, many columns (the look-ahead panel update, TRSM tails)
    if (narrow_tiles && n <= 512) {  // least padding of n; ties go to the wider tile
        auto padded = [n](int64_t bn) { return (n + bn - 1) / bn * bn; };
        const int64_t p128 = padded(128), p112 = padded(112), p80 = padded(80);
        if (p112 < p128 && p112 <= p80) return GS_128x112;
        if (p80 < p128 && p80 < p112) return GS_128x80;
    }
    // 16 warps with 32x32 warp tiles on the 128x128 tile (twice the warps per SM to cover the
    // shared-load -> DMMA latency): 29.0 vs 27.0 TFLOP/s at 8192^3; MLFFPC_DGEMM16=0 selects the 8-warp kernel
    return warps16 ? GS_128x128_16W : GS_128x128_8W;
}

template <int BM, int BN, int WARPS_M, int WARPS_N>
static int launch_shape(bool transB, bool vec2, int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda,
                        const double* B, int64_t ldb, double beta, double* C, int64_t ldc, bool lower_only,
                        cudaStream_t s, int nsplit, int64_t c_zstride) {
    if (transB)
        return vec2 ? launch_dgemm<BM, BN, WARPS_M, WARPS_N, true, true>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, lower_only, s, nsplit, c_zstride)
                    : launch_dgemm<BM, BN, WARPS_M, WARPS_N, true, false>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, lower_only, s, nsplit, c_zstride);
    return vec2 ? launch_dgemm<BM, BN, WARPS_M, WARPS_N, false, true>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, lower_only, s, nsplit, c_zstride)
                : launch_dgemm<BM, BN, WARPS_M, WARPS_N, false, false>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, lower_only, s, nsplit, c_zstride);
}

template <int BM, int BN, int WARPS_M, int WARPS_N>
static int shape_ctas_per_sm() {  // resident CTAs per SM of the non-transposed, 16-byte staged variant
    using SM = GemmSmem<BM, BN, false>;
    auto kern = dgemm_kernel<BM, BN, WARPS_M, WARPS_N, false, true>;
    int nb = 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::BYTES) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, WARPS_M * WARPS_N * 32, SM::BYTES) != cudaSuccess || nb < 1) {
        cudaGetLastError();
        nb = 1;
    }
    return nb;
}

#define MLFFPC_GEMM_SHAPES(X)                                                                                     \
    X(GS_128x128_16W, 128, 128, 4, 4) X(GS_128x128_8W, 128, 128, 2, 4) X(GS_128x64, 128, 64, 4, 2)              \
    X(GS_64x128, 64, 128, 2, 4) X(GS_128x112, 128, 112, 8, 2) X(GS_128x80, 128, 80, 8, 2) X(GS_256x40, 256, 40, 16, 1)

int dgemm(bool transB, int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda,
          const double* B, int64_t ldb, double beta, double* C, int64_t ldc, bool lower_only,
          cudaStream_t s, int nsplit, int64_t c_zstride) {
    if (m <= 0 || n <= 0) return MLFFPC_OK;
    if (nsplit < 1) nsplit = 1;
    const bool vec2 = (lda % 2 == 0) && (ldb % 2 == 0) && (((uintptr_t)A | (uintptr_t)B) % 16 == 0);
    switch (gemm_pick_shape(m, n)) {
#define X(ID, BM, BN, WM, WN) \
    case ID: return launch_shape<BM, BN, WM, WN>(transB, vec2, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, lower_only, s, nsplit, c_zstride);
        MLFFPC_GEMM_SHAPES(X)
#undef X
    }
    return MLFFPC_ERR_INVALID;
}

// Split-K factor for a product with few output tiles and a long k (C = A[m,k] B[k,n], row-major B): the number of
// slices that fills whole waves of the tile shape dgemm() will pick, each slice keeping >= 512 columns of k.
int dgemm_split_k(int64_t m, int64_t n, int64_t k, int num_sms) {
    int bm = 128, bn = 128, occ = 1;
    switch (gemm_pick_shape(m, n)) {
#define X(ID, BM, BN, WM, WN) \
    case ID: { static const int o = shape_ctas_per_sm<BM, BN, WM, WN>(); bm = BM; bn = BN; occ = o; break; }
        MLFFPC_GEMM_SHAPES(X)
#undef X
    }
    const int64_t tiles = ((m + bm - 1) / bm) * ((n + bn - 1) / bn);
    const int64_t slots = (int64_t)num_sms * occ;
    int64_t max_ns = k / 512;
    if (max_ns > 32) max_ns = 32;
    if (max_ns < 1) max_ns = 1;
    int best = 1;
    double best_eff = -1.0;
    for (int64_t ns = 1; ns <= max_ns; ++ns) {
        const int64_t ctas = tiles * ns, waves = (ctas + slots - 1) / slots;
        // efficiency of the last wave, discounted when the grid does not fill the GPU once
        const double eff = (double)ctas / (double)(waves * slots);
        if (eff > best_eff + 0.01) { best_eff = eff; best = (int)ns; }
    }
    return best;
}

// ---- SYRK helpers ---------------------------------------------------------------------------
__global__ void symmetrize_shift_kernel(double* W, int64_t m, int64_t ldw, double shift) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r = blockIdx.y;
    if (c >= m) return;
    if (c > r) W[r * ldw + c] = W[c * ldw + r];
    else if (c == r) W[r * ldw + c] += shift;
}

// W (+ running compensation C) += P, entrywise Kahan summation over the lower triangle (first = 1: W = P, C = 0)
__global__ void kahan_accumulate_kernel(double* __restrict__ W, double* __restrict__ C, const double* __restrict__ P,
                                        int64_t m, int64_t ldw, int first) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r = blockIdx.y;
    if (c >= m || c > r) return;
    const int64_t iw = r * ldw + c, ip = r * m + c;
    if (first) { W[iw] = P[ip]; C[ip] = 0.0; return; }
    const double y = P[ip] - C[ip];
    const double t = W[iw] + y;
    C[ip] = (t - W[iw]) - y;
    W[iw] = t;
}

// ---- blocked Cholesky (lower, in place) --------------------------------------------------------
constexpr int POTRF_NB = 64;

// factor the nb x nb diagonal block at (j0, j0) in shared memory; writes info (1-based column) on breakdown
__global__ void potrf_diag_kernel(double* W, int64_t ldw, int64_t j0, int nb, int* info) {
    __shared__ double a[POTRF_NB][POTRF_NB + 1];
    const int tid = threadIdx.x;
    for (int t = tid; t < nb * nb; t += blockDim.x) a[t / nb][t % nb] = W[(j0 + t / nb) * ldw + j0 + t % nb];
    __syncthreads();
    for (int j = 0; j < nb; ++j) {
        __shared__ double piv;
        if (tid == 0) {
            const double d = a[j][j];
            if (!(d > 0.0)) {
                if (*info == 0) *info = (int)(j0 + j + 1);
                piv = 1.0;  // keep going with garbage; caller reads info
            } else {
                piv = sqrt(d);
            }
            a[j][j] = piv;
        }
        __syncthreads();
        const double pj = a[j][j];
        for (int r = j + 1 + tid; r < nb; r += blockDim.x) a[r][j] /= pj;
        __syncthreads();
        // trailing update of the block: a[r][c] -= a[r][j] a[c][j], c in (j, r]
        for (int t = tid; t < (nb - j - 1) * (nb - j - 1); t += blockDim.x) {
            const int r = j + 1 + t / (nb - j - 1), c = j + 1 + t % (nb - j - 1);
            if (c <= r) a[r][c] -= a[r][j] * a[c][j];
        }
        __syncthreads();
    }
    for (int t = tid; t < nb * nb; t += blockDim.x) {
        const int r = t / nb, c = t % nb;
        W[(j0 + r) * ldw + j0 + c] = (c <= r) ? a[r][c] : 0.0;
    }
}

// panel below the diagonal block: rows r in [j0+nb, m): W[r, j0:j0+nb] <- W[r, j0:j0+nb] L11^{-T}; one thread per row
__global__ void potrf_panel_kernel(double* W, int64_t ldw, int64_t j0, int nb, int64_t m) {
    __shared__ double l11[POTRF_NB][POTRF_NB + 1];
    for (int t = threadIdx.x; t < nb * nb; t += blockDim.x) l11[t / nb][t % nb] = W[(j0 + t / nb) * ldw + j0 + t % nb];
    __syncthreads();
    const int64_t r = j0 + nb + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    double x[POTRF_NB];
    double* row = W + r * ldw + j0;
#pragma unroll 4
    for (int c = 0; c < nb; ++c) x[c] = row[c];
    for (int c = 0; c < nb; ++c) {
        double v = x[c];
        for (int t = 0; t < c; ++t) v = fma(-x[t], l11[c][t], v);
        x[c] = v / l11[c][c];
    }
    for (int c = 0; c < nb; ++c) row[c] = x[c];
}

__global__ void zero_upper_kernel(double* W, int64_t m, int64_t ldw) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r = blockIdx.y;
    if (c < m && c > r) W[r * ldw + c] = 0.0;
}

// ---- blocked row-TRSM: X <- Lf^{-1} X, X[m, n_cols] row-major ------------------------------------
constexpr int TRSM_NB = 32;
constexpr int TRSM_OB = 256;

// diagonal solve for block rows [j0, j0+nb): one thread per column of X
__global__ void trsm_diag_kernel(const double* __restrict__ Lf, int64_t ldl, int64_t j0, int nb, double* X,
                                 int64_t n_cols, int64_t ldx) {
    __shared__ double l[TRSM_NB][TRSM_NB + 1];
    for (int t = threadIdx.x; t < nb * nb; t += blockDim.x) l[t / nb][t % nb] = Lf[(j0 + t / nb) * ldl + j0 + t % nb];
    __syncthreads();
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cols) return;
    double x[TRSM_NB];
#pragma unroll
    for (int r = 0; r < TRSM_NB; ++r)
        if (r < nb) x[r] = X[(j0 + r) * ldx + c];
#pragma unroll
    for (int r = 0; r < TRSM_NB; ++r) {
        if (r < nb) {
            double v = x[r];
#pragma unroll
            for (int t = 0; t < TRSM_NB; ++t)
                if (t < r) v = fma(-l[r][t], x[t], v);
            x[r] = v / l[r][r];
        }
    }
#pragma unroll
    for (int r = 0; r < TRSM_NB; ++r)
        if (r < nb) X[(j0 + r) * ldx + c] = x[r];
}

}  // namespace mlffpc

using namespace mlffpc;

extern "C" {

int mlffpc_dgemm(mlffpc_ctx* ctx, int trans_b, int64_t m, int64_t n, int64_t k, double alpha,
                 const double* A, int64_t lda, const double* B, int64_t ldb, double beta, double* C,
                 int64_t ldc, void* stream) {
    MLFFPC_REQUIRE(ctx && A && B && C, "dgemm: NULL argument");
    MLFFPC_REQUIRE(m >= 0 && n >= 0 && k >= 0 && lda >= k && ldc >= n && ldb >= (trans_b ? k : n),
                   "dgemm: bad dimensions");
    return dgemm(trans_b != 0, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, false, (cudaStream_t)stream);
}

int mlffpc_syrk_rows(mlffpc_ctx* ctx, const double* X, int64_t m, int64_t n_cols, int64_t ldx,
                     double shift, double* W, int64_t ldw, void* stream) {
    MLFFPC_REQUIRE(ctx && X && W && m > 0 && n_cols >= 0 && ldx >= n_cols && ldw >= m, "syrk_rows: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    ProfWindow pw = prof_window("syrk");
    pw.step(pw.first);
    int st_gemm = MLFFPC_OK;
    const int64_t chunk = ctx->syrk_chunk;
    if (ctx->gram_mode == 1 && chunk <= 0) {
        // default: k-tile products folded into an unevaluated (hi, lo) sum on the DMMA pipe (gramdd.cu); includes the
        // cross-rank reduction, the mirror to the upper triangle and the diagonal shift
        st_gemm = gram_dd(ctx, X, m, n_cols, ldx, W, ldw, shift, false, s);
        pw.end();
        return st_gemm;
    }
    if (chunk > 0 && n_cols > chunk) {
        // The Gram of 1e5-long rows loses ~sqrt(n) eps relative accuracy in a single running sum; against
        // lam = 1e-10 that is a visible perturbation of the Woodbury inverse ("Woodbury accuracy").
        // Column chunks are multiplied separately (DMMA) and their partial Grams added with Kahan compensation.
        double* tmp = nullptr;
        MLFFPC_CUDA(cudaMallocAsync((void**)&tmp, (size_t)(2 * m * m) * sizeof(double), s));
        double* comp = tmp + m * m;
        const dim3 grid((unsigned)((m + 255) / 256), (unsigned)m);
        for (int64_t c0 = 0; c0 < n_cols && st_gemm == MLFFPC_OK; c0 += chunk) {
            const int64_t w = (n_cols - c0 < chunk) ? (n_cols - c0) : chunk;
            st_gemm = dgemm(true, m, m, w, 1.0, X + c0, ldx, X + c0, ldx, 0.0, tmp, m, true, s);
            if (st_gemm != MLFFPC_OK) break;
            kahan_accumulate_kernel<<<grid, 256, 0, s>>>(W, comp, tmp, m, ldw, c0 == 0 ? 1 : 0);
            ++g_launches;
        }
        cudaFreeAsync(tmp, s);
    } else {
        st_gemm = dgemm(true, m, m, n_cols, 1.0, X, ldx, X, ldx, 0.0, W, ldw, true, s);
    }
    pw.end();
    MLFFPC_TRY(st_gemm);
    if (ctx->comm.world > 1) {
        // the lower tiles of every rank are summed; entries above the diagonal tiles are rebuilt below
        MLFFPC_REQUIRE(ldw == m, "syrk_rows: multi-GPU reduction needs a packed W (ldw == m)");
        symmetrize_shift_kernel<<<dim3((unsigned)((m + 255) / 256), (unsigned)m), 256, 0, s>>>(W, m, ldw, 0.0);
        MLFFPC_LAUNCH_CHECK();
        MLFFPC_TRY(comm_allreduce_sum(ctx->comm, W, (size_t)(m * m), s));
    }
    symmetrize_shift_kernel<<<dim3((unsigned)((m + 255) / 256), (unsigned)m), 256, 0, s>>>(W, m, ldw, shift);
    MLFFPC_LAUNCH_CHECK();
    return MLFFPC_OK;
}

int mlffpc_potrf_lower(mlffpc_ctx* ctx, double* W, int64_t m, int64_t ldw, int* info_host, void* stream) {
    MLFFPC_REQUIRE(ctx && W && info_host && m > 0 && ldw >= m, "potrf_lower: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    int* d_info = (int*)(ctx->scal + MLFFPC_NUM_SCAL - 2);
    MLFFPC_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int), s));
    ProfWindow pw = prof_window("potrf");
    for (int64_t j0 = 0; j0 < m; j0 += POTRF_NB) {
        pw.step(j0 / POTRF_NB);
        const int nb = (int)((m - j0 < POTRF_NB) ? (m - j0) : POTRF_NB);
        potrf_diag_kernel<<<1, 256, 0, s>>>(W, ldw, j0, nb, d_info);
        MLFFPC_LAUNCH_CHECK();
        const int64_t rest = m - j0 - nb;
        if (rest > 0) {
            potrf_panel_kernel<<<(unsigned)((rest + 127) / 128), 128, 0, s>>>(W, ldw, j0, nb, m);
            MLFFPC_LAUNCH_CHECK();
            // A22 -= L21 L21^T (lower tiles only)
            double* A22 = W + (j0 + nb) * ldw + (j0 + nb);
            const double* L21 = W + (j0 + nb) * ldw + j0;
            MLFFPC_TRY(dgemm(true, rest, rest, nb, -1.0, L21, ldw, L21, ldw, 1.0, A22, ldw, true, s));
        }
    }
    pw.end();
    zero_upper_kernel<<<dim3((unsigned)((m + 255) / 256), (unsigned)m), 256, 0, s>>>(W, m, ldw);
    MLFFPC_LAUNCH_CHECK();
    int* h_info = (int*)(ctx->h_scal + MLFFPC_NUM_SCAL - 2);
    MLFFPC_CUDA(cudaMemcpyAsync(h_info, d_info, sizeof(int), cudaMemcpyDeviceToHost, s));
    MLFFPC_CUDA(cudaStreamSynchronize(s));
    *info_host = *h_info;
    return MLFFPC_OK;
}

int mlffpc_trsm_rows(mlffpc_ctx* ctx, const double* Lf, int64_t m, int64_t ldl, double* X,
                     int64_t n_cols, int64_t ldx, void* stream) {
    MLFFPC_REQUIRE(ctx && Lf && X && m > 0 && ldl >= m && n_cols >= 0 && ldx >= n_cols, "trsm_rows: bad argument");
    if (n_cols == 0) return MLFFPC_OK;
    cudaStream_t s = (cudaStream_t)stream;
    // Two-level blocking.  Outer level, left-looking: the TRSM_OB rows of a panel first receive the contribution of ALL
    // solved rows above them in one GEMM with a long k (X[J0:J1] -= Lf[J0:J1, 0:J0] X[0:J0]: every row of X is written
    // once); right-looking: each solved panel updates everything below it with a k = 256 GEMM (re-reads and re-writes
    // the rest of X per slab).  On 108 000 columns the first runs at 29 instead of 21 TFLOP/s (121 vs 136 ms); on the short
    // per-rank slices of an 8-GPU run the two cost the same, and there the right-looking order is kept: for cfg2 the
    // measured defect matrix E of the projected form is only just accurate enough, and the last bits of the factor move
    // the CG iteration count on 8 GPUs between 936 (this order) and 1123-1140 (the other; profiles/r02p_*, DESIGN.md
    // sections 5 and 10).  MLFFPC_TRSM_RIGHT=0/1 forces one.
    // Inside a panel the 32-row diagonal solves update only the rest of the panel.
    ProfWindow pw = prof_window("trsm");
    static const int forced = [] { const char* e = getenv("MLFFPC_TRSM_RIGHT"); return !e ? -1 : (e[0] == '1' ? 1 : 0); }();
    const bool right_looking = forced >= 0 ? forced == 1 : (ctx->trsm_order >= 0 ? ctx->trsm_order == 1 : n_cols < 32768);
    for (int64_t J0 = 0; J0 < m; J0 += TRSM_OB) {
        pw.step(J0 / TRSM_OB);
        const int64_t J1 = (J0 + TRSM_OB < m) ? (J0 + TRSM_OB) : m;
        if (J0 > 0 && !right_looking) {
            MLFFPC_TRY(dgemm(false, J1 - J0, n_cols, J0, -1.0, Lf + J0 * ldl, ldl, X, ldx, 1.0, X + J0 * ldx, ldx, false, s));
        }
        for (int64_t j0 = J0; j0 < J1; j0 += TRSM_NB) {
            const int nb = (int)((J1 - j0 < TRSM_NB) ? (J1 - j0) : TRSM_NB);
            trsm_diag_kernel<<<(unsigned)((n_cols + 127) / 128), 128, 0, s>>>(Lf, ldl, j0, nb, X, n_cols, ldx);
            MLFFPC_LAUNCH_CHECK();
            const int64_t rest = J1 - j0 - nb;
            if (rest > 0) {
                // X[j0+nb:J1, :] -= Lf[j0+nb:J1, j0:j0+nb] X[j0:j0+nb, :]
                MLFFPC_TRY(dgemm(false, rest, n_cols, nb, -1.0, Lf + (j0 + nb) * ldl + j0, ldl, X + j0 * ldx, ldx,
                                 1.0, X + (j0 + nb) * ldx, ldx, false, s));
            }
        }
        if (right_looking && m - J1 > 0) {  // X[J1:] -= Lf[J1:, J0:J1] X[J0:J1]
            MLFFPC_TRY(dgemm(false, m - J1, n_cols, J1 - J0, -1.0, Lf + J1 * ldl + J0, ldl, X + J0 * ldx, ldx, 1.0,
                             X + J1 * ldx, ldx, false, s));
        }
    }
    pw.end();
    return MLFFPC_OK;
}

}  // extern "C"
