// Gram matrices of long rows with extended-precision accumulation on the FP64 tensor pipe.
//
//   W = X X^T,  X[m, n_cols] row-major with n_cols ~ 1e5   (L^T L of the Woodbury formula, the CholeskyQR
//   passes, the Nystroem inner matrices: reference iterative_cholesky.py:141, iterative_solver.py:230, :535)
//
// A single running fp64 sum over 1e5 terms carries ~ eps sqrt(n) |W| of rounding, which against the
// regularisation lam = 1e-10 is a visible perturbation of the preconditioner (measured on the reference's own
// CPU-runnable case, n = 9990: 188 CG iterations with the plain DMMA Gram, 119 with LAPACK's blocked Gram and
// 119 with this kernel).  Here every 16-column k-tile is multiplied on the DMMA pipe into a *zeroed*
// accumulator (16 terms: rounding ~ eps |partial|) and folded into an unevaluated sum (hi, lo) with an
// error-free TwoSum, so the result is the exact sum of the k-tile products up to O(eps^2).  The same kernel
// yields E = Q Q^T - I of a numerically orthonormal factor to ~1e-18 per entry -- what the two-pass projected
// preconditioner apply needs (precon.cu) and what a plain fp64 Gram cannot deliver because its own rounding is
// as large as E.
//
// Across ranks (rows of X sharded by columns) the partial (hi, lo) pairs are gathered and folded in rank order
// with the same TwoSum, so every rank holds bit-identical results.
#include "common.cuh"

namespace mlffpc {

namespace {

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
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

// (hi, lo) += v, error-free: hi + lo + v is preserved exactly up to the rounding of lo (~eps^2)
__device__ __forceinline__ void dd_add(double& hi, double& lo, const double v) {
    const double s = hi + v;
    const double bb = s - hi;
    const double e = (hi - (s - bb)) + (v - bb);
    hi = s;
    lo += e;
}

constexpr int GD_BM = 64, GD_BK = 16, GD_STAGES = 3, GD_THREADS = 128;
constexpr int GD_STRIDE = GD_BK + 4;            // == 4 (mod 16): conflict-free m8n8k4 fragment loads
constexpr int GD_TILE = GD_BM * GD_STRIDE;       // doubles per operand tile
constexpr size_t GD_SMEM = (size_t)GD_STAGES * 2 * GD_TILE * sizeof(double);

// lower 64 x 64 tiles of X X^T; slice blockIdx.z takes the k-tiles [z * kt_per, (z + 1) * kt_per) and writes its
// own (hi, lo) pair at offset z * zstride
template <bool VEC2, int FOLD>
__global__ void __launch_bounds__(GD_THREADS, 2)
gram_dd_kernel(const double* __restrict__ X, int64_t m, int64_t k, int64_t ldx, double* __restrict__ Whi,
               double* __restrict__ Wlo, int64_t ldw, int64_t kt_per, int64_t zstride) {
    extern __shared__ double gd_smem[];
    // linear index over the lower block triangle -> (bi, bj), bj <= bi
    const int64_t t = blockIdx.x;
    int64_t bi = (int64_t)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while (bi * (bi + 1) / 2 > t) --bi;
    while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
    const int64_t bj = t - bi * (bi + 1) / 2;
    const int64_t m0 = bi * GD_BM, n0 = bj * GD_BM;

    const int64_t ktiles_all = (k + GD_BK - 1) / GD_BK;
    const int64_t kt0 = (int64_t)blockIdx.z * kt_per;
    int64_t kt1 = kt0 + kt_per;
    if (kt1 > ktiles_all) kt1 = ktiles_all;
    const int64_t ktiles = kt1 > kt0 ? kt1 - kt0 : 0;
    Whi += (int64_t)blockIdx.z * zstride;
    Wlo += (int64_t)blockIdx.z * zstride;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
    const int lr = lane >> 2, lc = lane & 3;

    double acc[4][4][2], hi[4][4][2], lo[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) acc[i][j][e] = hi[i][j][e] = lo[i][j][e] = 0.0;

    auto load_stage = [&](int stage, int64_t kt) {
        double* As = gd_smem + (size_t)stage * 2 * GD_TILE;
        double* Bs = As + GD_TILE;
        const int64_t k0 = (kt0 + kt) * GD_BK;
        constexpr int V = VEC2 ? 2 : 1;
        constexpr int ITEMS = GD_BM * GD_BK / V;
        for (int q = tid; q < ITEMS; q += GD_THREADS) {
            const int r = q / (GD_BK / V), c = (q % (GD_BK / V)) * V;
            const int64_t kc = k0 + c;
            {
                const bool rok = m0 + r < m;
                const bool ok = rok && kc < k;
                const double* src = ok ? (X + (m0 + r) * ldx + kc) : X;
                if (VEC2) cp_async16(As + r * GD_STRIDE + c, src, !ok ? 0 : (kc + 1 < k ? 16 : 8));
                else cp_async8(As + r * GD_STRIDE + c, src, ok);
            }
            {
                const bool rok = n0 + r < m;
                const bool ok = rok && kc < k;
                const double* src = ok ? (X + (n0 + r) * ldx + kc) : X;
                if (VEC2) cp_async16(Bs + r * GD_STRIDE + c, src, !ok ? 0 : (kc + 1 < k ? 16 : 8));
                else cp_async8(Bs + r * GD_STRIDE + c, src, ok);
            }
        }
    };

#pragma unroll
    for (int s = 0; s < GD_STAGES - 1; ++s) {
        if (s < ktiles) load_stage(s, s);
        cp_async_commit();
    }
    for (int64_t kt = 0; kt < ktiles; ++kt) {
        cp_async_wait<GD_STAGES - 2>();
        __syncthreads();
        {
            const int64_t nk = kt + GD_STAGES - 1;
            if (nk < ktiles) load_stage((int)(nk % GD_STAGES), nk);
            cp_async_commit();
        }
        const double* As = gd_smem + (size_t)(kt % GD_STAGES) * 2 * GD_TILE;
        const double* Bs = As + GD_TILE;
#pragma unroll
        for (int kk = 0; kk < GD_BK; kk += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = As[(wm + i * 8 + lr) * GD_STRIDE + kk + lc];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = Bs[(wn + j * 8 + lr) * GD_STRIDE + kk + lc];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        // fold the 16 FOLD-term partial products into the unevaluated sums and restart the accumulators
        if (FOLD > 1 && ((kt + 1) % FOLD) != 0 && kt + 1 < ktiles) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    dd_add(hi[i][j][e], lo[i][j][e], acc[i][j][e]);
                    acc[i][j][e] = 0.0;
                }
    }
    cp_async_wait<0>();

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t row = m0 + wm + i * 8 + lr;
        if (row >= m) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t col = n0 + wn + j * 8 + 2 * lc;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                if (col + e < m) {
                    Whi[row * ldw + col + e] = hi[i][j][e];
                    Wlo[row * ldw + col + e] = lo[i][j][e];
                }
            }
        }
    }
}

// The same tiles with every product formed exactly (TwoProduct by FMA) and added with TwoSum on the FP64 vector
// pipe: ~10 instructions per term instead of 1/256 of a DMMA, so ~7x the time of the kernel above -- the reference
// answer for tests and for option "defect_mode" = 2.  256 threads, 4 x 4 outputs per thread.
__global__ void __launch_bounds__(256)
gram_dd_exact_kernel(const double* __restrict__ X, int64_t m, int64_t k, int64_t ldx, double* __restrict__ Whi,
                     double* __restrict__ Wlo, int64_t ldw, int64_t kt_per, int64_t zstride) {
    __shared__ double As[GD_BK][GD_BM + 1];
    __shared__ double Bs[GD_BK][GD_BM + 1];
    const int64_t t = blockIdx.x;
    int64_t bi = (int64_t)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while (bi * (bi + 1) / 2 > t) --bi;
    while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
    const int64_t bj = t - bi * (bi + 1) / 2;
    const int64_t m0 = bi * GD_BM, n0 = bj * GD_BM;
    const int64_t ktiles_all = (k + GD_BK - 1) / GD_BK;
    const int64_t kt0 = (int64_t)blockIdx.z * kt_per;
    int64_t kt1 = kt0 + kt_per;
    if (kt1 > ktiles_all) kt1 = ktiles_all;
    Whi += (int64_t)blockIdx.z * zstride;
    Wlo += (int64_t)blockIdx.z * zstride;
    const int tid = threadIdx.x, tr = tid >> 4, tc = tid & 15;  // outputs (tr + 16 i, tc + 16 j)
    double hi[4][4], lo[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) hi[i][j] = lo[i][j] = 0.0;
    for (int64_t kt = kt0; kt < kt1; ++kt) {
        const int64_t k0 = kt * GD_BK;
        for (int q = tid; q < GD_BM * GD_BK; q += 256) {
            const int r = q / GD_BK, c = q % GD_BK;
            const bool kok = k0 + c < k;
            As[c][r] = (kok && m0 + r < m) ? X[(m0 + r) * ldx + k0 + c] : 0.0;
            Bs[c][r] = (kok && n0 + r < m) ? X[(n0 + r) * ldx + k0 + c] : 0.0;
        }
        __syncthreads();
#pragma unroll 4
        for (int kk = 0; kk < GD_BK; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][tr + 16 * i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tc + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double pr = a[i] * b[j];
                    const double pe = fma(a[i], b[j], -pr);
                    dd_add(hi[i][j], lo[i][j], pr);
                    lo[i][j] += pe;
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t row = m0 + tr + 16 * i;
        if (row >= m) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t col = n0 + tc + 16 * j;
            if (col < m) {
                Whi[row * ldw + col] = hi[i][j];
                Wlo[row * ldw + col] = lo[i][j];
            }
        }
    }
}

// out (hi, lo)[i] = sum over slices z of (hi_z[i], lo_z[i]) in slice order (error-free in hi, plain in lo);
// `count` contiguous elements per slice, slices `stride` apart, lo_z = hi_z + lo_off
__global__ void dd_fold_slices_kernel(const double* __restrict__ hi_slices, int64_t lo_off, int64_t stride, int nslices,
                                      int64_t count, double* __restrict__ out_hi, double* __restrict__ out_lo) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    double H = hi_slices[i], L = hi_slices[lo_off + i];
    for (int z = 1; z < nslices; ++z) {
        dd_add(H, L, hi_slices[(int64_t)z * stride + i]);
        L += hi_slices[(int64_t)z * stride + lo_off + i];
    }
    out_hi[i] = H;
    out_lo[i] = L;
}

// lower triangle of (hi, lo) -> full symmetric fp64 matrix: out = hi + lo (+ shift on the diagonal), or with
// minus_identity the defect (hi - 1) + lo on the diagonal (hi - 1 is exact for hi in [1/2, 2])
__global__ void dd_finish_kernel(const double* __restrict__ hi, const double* __restrict__ lo, int64_t m, int64_t ldw,
                                 double* __restrict__ out, int64_t ld_out, double shift, int minus_identity) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r = blockIdx.y;
    if (c > r || c >= m) return;
    double h = hi[r * ldw + c];
    const double l = lo[r * ldw + c];
    if (r == c && minus_identity) h -= 1.0;
    double v = h + l;
    if (r == c && !minus_identity) v += shift;
    out[r * ld_out + c] = v;
    out[c * ld_out + r] = v;
}

}  // namespace

// W[m, m] (ld_out) = X X^T + shift I, or X X^T - I when minus_identity; X[m, n_cols] are this rank's columns.
// Scratch (2 m^2 doubles + split-K slices) comes from the stream-ordered allocator.
int gram_dd(mlffpc_ctx* ctx, const double* X, int64_t m, int64_t n_cols, int64_t ldx, double* out, int64_t ld_out,
            double shift, bool minus_identity, cudaStream_t s, bool exact) {
    if (m <= 0) return MLFFPC_OK;
    const int64_t nb = (m + GD_BM - 1) / GD_BM;
    const int64_t tiles = nb * (nb + 1) / 2;
    const int64_t ktiles = (n_cols + GD_BK - 1) / GD_BK;
    // split-K when the lower block triangle alone cannot fill the GPU (two CTAs per SM); >= 64 k-tiles per slice
    int64_t nsplit = 1;
    if (tiles < 2 * (int64_t)ctx->num_sms) {
        nsplit = (2 * (int64_t)ctx->num_sms + tiles - 1) / tiles;
        const int64_t cap = ktiles / 64 > 1 ? ktiles / 64 : 1;
        if (nsplit > cap) nsplit = cap;
        if (nsplit > 64) nsplit = 64;
    }
    const int64_t kt_per = (ktiles + nsplit - 1) / nsplit;
    const int64_t mm = m * m;
    double* buf = nullptr;
    MLFFPC_CUDA(cudaMallocAsync((void**)&buf, (size_t)(2 * mm * (nsplit > 1 ? nsplit + 1 : 1)) * sizeof(double), s));
    double* hi = buf;            // final pair first
    double* lo = buf + mm;
    double* slices = buf + 2 * mm;  // [nsplit][2][m*m] when nsplit > 1
    const bool vec2 = (ldx % 2 == 0) && ((uintptr_t)X % 16 == 0);
    int st = MLFFPC_OK;
    do {
        // k-tiles per fold (option "gram_fold").  Measured on cfg2 (n = 108 000, profiles/r02k_*): folding every k-tile
        // 920 CG iterations, every 2nd 945, every 4th 1124 -- the 17 ms per Gram that a longer interval saves cost far
        // more in iterations, so the default stays 1.
        int fold = ctx->gram_fold;
        if (fold == 0) fold = 1;
        auto kern = fold >= 4 ? (vec2 ? gram_dd_kernel<true, 4> : gram_dd_kernel<false, 4>)
                  : fold == 2 ? (vec2 ? gram_dd_kernel<true, 2> : gram_dd_kernel<false, 2>)
                              : (vec2 ? gram_dd_kernel<true, 1> : gram_dd_kernel<false, 1>);
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GD_SMEM);
        if (e != cudaSuccess) { st = cuda_fail(e, "gram_dd attribute", __FILE__, __LINE__); break; }
        if (tiles > 0x7fffffffLL) { set_error("gram_dd: m = %lld too large", (long long)m); st = MLFFPC_ERR_INVALID; break; }
        const dim3 grid((unsigned)tiles, 1, (unsigned)nsplit);
        double* dst_hi = nsplit == 1 ? hi : slices;
        double* dst_lo = nsplit == 1 ? lo : slices + mm;
        const int64_t zs = nsplit == 1 ? 0 : 2 * mm;
        if (exact) gram_dd_exact_kernel<<<grid, 256, 0, s>>>(X, m, n_cols, ldx, dst_hi, dst_lo, m, kt_per, zs);
        else kern<<<grid, GD_THREADS, GD_SMEM, s>>>(X, m, n_cols, ldx, dst_hi, dst_lo, m, kt_per, zs);
        ++g_launches;
        e = cudaGetLastError();
        if (e != cudaSuccess) { st = cuda_fail(e, "gram_dd launch", __FILE__, __LINE__); break; }
        if (nsplit > 1) {
            // entries above the block diagonal of a slice are never written; they are folded too but never read
            dd_fold_slices_kernel<<<(unsigned)((mm + 255) / 256), 256, 0, s>>>(slices, mm, 2 * mm, (int)nsplit, mm, hi, lo);
            ++g_launches;
        }
        const int world = ctx->comm.world;
        if (world > 1) {
            // gather the ranks' pairs by row blocks and fold them in rank order: bit-identical on every rank
            int64_t blk_rows = ((int64_t)16 << 20) / (2 * (int64_t)world * m);  // <= 128 MB of gather buffer
            if (blk_rows < 1) blk_rows = 1;
            if (blk_rows > m) blk_rows = m;
            const int64_t blk = blk_rows * m;
            double* send = nullptr;
            e = cudaMallocAsync((void**)&send, (size_t)(2 * blk * (world + 1)) * sizeof(double), s);
            if (e != cudaSuccess) { st = cuda_fail(e, "gram_dd gather buffer", __FILE__, __LINE__); break; }
            double* recv = send + 2 * blk;
            for (int64_t r0 = 0; r0 < m && st == MLFFPC_OK; r0 += blk_rows) {
                const int64_t rows = (m - r0 < blk_rows) ? (m - r0) : blk_rows;
                const int64_t cnt = rows * m;
                e = cudaMemcpyAsync(send, hi + r0 * m, (size_t)cnt * 8, cudaMemcpyDeviceToDevice, s);
                if (e == cudaSuccess) e = cudaMemcpyAsync(send + blk, lo + r0 * m, (size_t)cnt * 8, cudaMemcpyDeviceToDevice, s);
                if (e != cudaSuccess) { st = cuda_fail(e, "gram_dd pack", __FILE__, __LINE__); break; }
                st = comm_allgather(ctx->comm, send, recv, (size_t)(2 * blk) * 8, s);
                if (st != MLFFPC_OK) break;
                dd_fold_slices_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, s>>>(recv, blk, 2 * blk, world, cnt,
                                                                                  hi + r0 * m, lo + r0 * m);
                ++g_launches;
            }
            cudaFreeAsync(send, s);
            if (st != MLFFPC_OK) break;
        }
        dd_finish_kernel<<<dim3((unsigned)((m + 255) / 256), (unsigned)m), 256, 0, s>>>(hi, lo, m, m, out, ld_out, shift,
                                                                                     minus_identity ? 1 : 0);
        ++g_launches;
        e = cudaGetLastError();
        if (e != cudaSuccess) { st = cuda_fail(e, "gram_dd finish", __FILE__, __LINE__); break; }
    } while (0);
    cudaFreeAsync(buf, s);
    return st;
}

}  // namespace mlffpc

using namespace mlffpc;

extern "C" {

int mlffpc_gram_defect(mlffpc_ctx* ctx, const double* Q, int64_t k, int64_t n_cols, int64_t ld, double* E, void* stream) {
    MLFFPC_REQUIRE(ctx && Q && E && k > 0 && n_cols >= 0 && ld >= n_cols, "gram_defect: bad argument");
    return gram_dd(ctx, Q, k, n_cols, ld, E, k, 0.0, true, (cudaStream_t)stream, ctx->defect_mode == 2);
}

}  // extern "C"
