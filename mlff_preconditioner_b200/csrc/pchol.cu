// Greedy diagonal-pivoted partial Cholesky of A = -K with on-the-fly columns.
// Mirrors incomplete_cholesky.pivoted_cholesky (reference solvers/incomplete_cholesky.py:41-80):
//   step m: i* = first argmax of diag over the *current permuted order* index_columns[m:], swap,
//           L[pi,m] = sqrt(diag[pi]);  L[rest,m] = (A[rest,pi] - L[rest,:m] L[pi,:m]) / L[pi,m];
//           diag[rest] -= L[rest,m]^2.
// The factor is kept transposed, Lt[m, r] = L[row0 + r, m], so that the Schur dot streams m
// coalesced rows of length n_local -- the 4 n k^2-byte HBM stream that bounds the build.
#include <vector>

#include "common.cuh"

namespace mlffpc {

int launch_columns_device_col(mlffpc_ctx* ctx, const int64_t* col_dev, double* out, double scale, cudaStream_t s);

struct Cand {
    double val;
    double pos;  // position in the permuted order (exact in a double: n < 2^40)
    double idx;  // global row index
    double pad;
};

__device__ __forceinline__ bool cand_better(double v, double p, double bv, double bp) {
    return (v > bv) || (v == bv && p < bp);
}

// block-wide argmax with first-position tie break; result in sm[0] (Cand)
__device__ __forceinline__ void block_argmax(double v, double p, double i, Cand* sm) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, o);
        const double op = __shfl_xor_sync(0xffffffffu, p, o);
        const double oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (cand_better(ov, op, v, p)) { v = ov; p = op; i = oi; }
    }
    __syncthreads();
    if (lane == 0) { sm[1 + w].val = v; sm[1 + w].pos = p; sm[1 + w].idx = i; }
    __syncthreads();
    if (w == 0) {
        if (lane < nw) { v = sm[1 + lane].val; p = sm[1 + lane].pos; i = sm[1 + lane].idx; }
        else { v = -1e300; p = 1e300; i = -1.0; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, v, o);
            const double op = __shfl_xor_sync(0xffffffffu, p, o);
            const double oi = __shfl_xor_sync(0xffffffffu, i, o);
            if (cand_better(ov, op, v, p)) { v = ov; p = op; i = oi; }
        }
        if (lane == 0) { sm[0].val = v; sm[0].pos = p; sm[0].idx = i; }
    }
    __syncthreads();
}

__global__ void pchol_init_kernel(int64_t n, int64_t* index_columns, int32_t* pos) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) { index_columns[t] = t; pos[t] = (int32_t)t; }
}

// per-CTA candidates over the local residual diagonal (used once, before step 0)
__global__ void pchol_scan_kernel(const double* __restrict__ diag, int64_t n_local, int64_t row0,
                                  const int32_t* __restrict__ pos, int64_t m, Cand* partials) {
    __shared__ Cand sm[40];
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double v = -1e300, p = 1e300, i = -1.0;
    if (r < n_local) {
        const int32_t ps = pos[row0 + r];
        if (ps >= m) { v = diag[r]; p = (double)ps; i = (double)(row0 + r); }
    }
    block_argmax(v, p, i, sm);
    if (threadIdx.x == 0) partials[blockIdx.x] = sm[0];
}

// one CTA: reduce `count` candidates -> best (local candidate of this rank) into out[0]
__global__ void pchol_reduce_kernel(const Cand* __restrict__ partials, int count, Cand* out) {
    __shared__ Cand sm[40];
    double v = -1e300, p = 1e300, i = -1.0;
    for (int t = threadIdx.x; t < count; t += blockDim.x) {
        const Cand c = partials[t];
        if (cand_better(c.val, c.pos, v, p)) { v = c.val; p = c.pos; i = c.idx; }
    }
    block_argmax(v, p, i, sm);
    if (threadIdx.x == 0) out[0] = sm[0];
}

// one thread: choose the global pivot among `world` rank candidates (or the forced one), apply the
// reference's swap to index_columns / pos, publish pivot index + sqrt(pivot) and flag non-PSD.
//   st[0] = pivot index (as int64 bits in piv_idx), st_d[0] = sqrt(pivot), flag != 0 -> not PSD
__global__ void pchol_select_kernel(const Cand* __restrict__ cands, int world, int64_t m,
                                    const int64_t* __restrict__ forced, const double* __restrict__ diag,
                                    int64_t row0, int64_t n_local, int64_t* index_columns, int32_t* pos,
                                    int64_t* piv_idx, double* piv_val, int* flag) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double v = -1e300, p = 1e300, i = -1.0;
    for (int r = 0; r < world; ++r) {
        const Cand c = cands[r];
        if (cand_better(c.val, c.pos, v, p)) { v = c.val; p = c.pos; i = c.idx; }
    }
    int64_t pi = (int64_t)i;
    if (forced) {
        pi = forced[m];
        // value of a forced pivot: only its owner knows it; the owner publishes, others get it by broadcast
        if (pi >= row0 && pi < row0 + n_local) v = diag[pi - row0];
        else v = 1.0;  // placeholder on non-owners (overwritten by the owner's broadcast)
    }
    const int32_t i_argmax = pos[pi];
    const int64_t e = index_columns[m];
    index_columns[m] = pi;
    index_columns[i_argmax] = e;
    pos[pi] = (int32_t)m;
    pos[e] = i_argmax;
    piv_idx[0] = pi;
    if (!(v > 0.0)) { if (*flag == 0) *flag = (int)(m + 1); v = 1.0; }
    piv_val[0] = sqrt(v);
}

// lrow[m'] = Lt[m', pi - row0] for m' < m on the owner, 0 elsewhere (replicated by an allreduce-sum)
__global__ void pchol_gather_row_kernel(const double* __restrict__ Lt, int64_t ld, int64_t m,
                                        const int64_t* __restrict__ piv_idx, int64_t row0, int64_t n_local,
                                        double* __restrict__ lrow) {
    const int64_t pi = piv_idx[0];
    const bool owner = (pi >= row0 && pi < row0 + n_local);
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < m) lrow[t] = owner ? Lt[t * ld + (pi - row0)] : 0.0;
}

// the HBM-bound step: Schur dot over m previous rows of Lt, new row m, diagonal update, candidates
constexpr int PCHOL_THREADS = 256;
template <int MSPLIT>
__global__ void __launch_bounds__(PCHOL_THREADS)
pchol_update_kernel(double* __restrict__ Lt, int64_t ld, int64_t m, int64_t n_local, int64_t row0,
                    const double* __restrict__ col, const double* __restrict__ lrow,
                    const int64_t* __restrict__ piv_idx, const double* __restrict__ piv_val,
                    double* __restrict__ diag, const int32_t* __restrict__ pos, Cand* partials) {
    constexpr int COLS = PCHOL_THREADS / MSPLIT;
    __shared__ double red[MSPLIT][COLS];
    __shared__ Cand sm[40];
    const int tc = threadIdx.x % COLS, ts = threadIdx.x / COLS;
    const int64_t r = (int64_t)blockIdx.x * COLS + tc;
    double acc = 0.0;
    if (r < n_local) {
        const double* Lp = Lt + r;
        int64_t q = ts;
        for (; q + 7 * MSPLIT < m; q += 8 * MSPLIT) {
            double t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) t[u] = __ldcs(Lp + (q + u * MSPLIT) * ld);
#pragma unroll
            for (int u = 0; u < 8; ++u) acc = fma(t[u], __ldg(lrow + q + u * MSPLIT), acc);
        }
        for (; q < m; q += MSPLIT) acc = fma(__ldcs(Lp + q * ld), __ldg(lrow + q), acc);
    }
    if (MSPLIT > 1) {
        red[ts][tc] = acc;
        __syncthreads();
        if (ts == 0) {
#pragma unroll
            for (int s = 1; s < MSPLIT; ++s) acc += red[s][tc];
        }
    }
    double v = -1e300, p = 1e300, i = -1.0;
    if (ts == 0 && r < n_local) {
        const int64_t g = row0 + r;
        const int32_t ps = pos[g];
        double l;
        if (ps > m) {  // still a candidate row: i_pi = index_columns[m+1:]
            l = (col[r] - acc) / piv_val[0];
            const double dn = diag[r] - l * l;
            diag[r] = dn;
            v = dn; p = (double)ps; i = (double)g;
        } else if (g == piv_idx[0]) {
            l = piv_val[0];
        } else {
            l = 0.0;
        }
        Lt[m * ld + r] = l;
    }
    block_argmax(v, p, i, sm);
    if (threadIdx.x == 0) partials[blockIdx.x] = sm[0];
}

static int pchol_msplit(int64_t n_local, int64_t m, int num_sms) {
    const int64_t want = (int64_t)num_sms * 1024;
    if (n_local >= want || m < 64) return 1;
    if (n_local * 4 >= want || m < 256) return 4;
    return 8;
}

struct PcholWs {
    int64_t off_col, off_lrow, off_pos, off_cands, off_gathered, off_piv, total;
};
static PcholWs pchol_layout(const mlffpc_ctx* c, int64_t k) {
    auto up = [](int64_t x) { return (x + 255) / 256 * 256; };
    PcholWs w;
    int64_t o = 0;
    w.off_col = o; o = up(o + c->n_local() * 8);
    w.off_lrow = o; o = up(o + (k + 1) * 8);
    w.off_pos = o; o = up(o + c->n * 4);
    w.off_cands = o; o = up(o + 64);                 // this rank's candidate
    w.off_gathered = o; o = up(o + 64 * 1024);        // all ranks' candidates (<= 1024 ranks)
    w.off_piv = o; o = up(o + 64);                    // piv_idx (int64), piv_val (double), flag (int)
    w.total = o + 256;
    return w;
}

}  // namespace mlffpc

using namespace mlffpc;

extern "C" {

int mlffpc_pchol_workspace_bytes(mlffpc_ctx* ctx, int64_t k, int64_t* bytes) {
    MLFFPC_REQUIRE(ctx && bytes && k >= 0 && ctx->M > 0, "pchol_workspace_bytes: bad argument / geometry not set");
    *bytes = pchol_layout(ctx, k).total;
    return MLFFPC_OK;
}

int mlffpc_pchol_build(mlffpc_ctx* ctx, int64_t k, double* Lt, int64_t ld, double* diag,
                       int64_t* index_columns, const int64_t* forced_pivots, float* step_ms_host,
                       void* workspace, int64_t workspace_bytes, void* stream) {
    MLFFPC_REQUIRE(ctx && ctx->M > 0, "pchol_build: geometry not set");
    MLFFPC_REQUIRE(diag && index_columns && workspace && (Lt || k == 0), "pchol_build: NULL argument");
    MLFFPC_REQUIRE(k >= 0 && k <= ctx->n, "max_rank = %lld is too large", (long long)k);
    const int64_t nl = ctx->n_local(), row0 = ctx->row0(), n = ctx->n;
    MLFFPC_REQUIRE(ld >= nl, "pchol_build: ld %lld < n_local %lld", (long long)ld, (long long)nl);
    const PcholWs w = pchol_layout(ctx, k);
    MLFFPC_REQUIRE(workspace_bytes >= w.total, "pchol_build: workspace too small (%lld < %lld)",
                   (long long)workspace_bytes, (long long)w.total);
    MLFFPC_REQUIRE(n < ((int64_t)1 << 31), "pchol_build: n too large for int32 positions");
    cudaStream_t s = (cudaStream_t)stream;
    char* base = (char*)(((uintptr_t)workspace + 255) / 256 * 256);
    double* col = (double*)(base + w.off_col);
    double* lrow = (double*)(base + w.off_lrow);
    int32_t* pos = (int32_t*)(base + w.off_pos);
    Cand* my_cand = (Cand*)(base + w.off_cands);
    Cand* all_cand = (Cand*)(base + w.off_gathered);
    int64_t* piv_idx = (int64_t*)(base + w.off_piv);
    double* piv_val = (double*)(base + w.off_piv + 8);
    int* flag = (int*)(base + w.off_piv + 16);
    Cand* partials = (Cand*)ctx->partials;  // MLFFPC_MAX_PARTIALS Cand slots (4 doubles each)
    const int world = ctx->comm.world;
    MLFFPC_REQUIRE(world <= 1024, "pchol_build: too many ranks");

    MLFFPC_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), s));
    pchol_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, index_columns, pos);
    MLFFPC_LAUNCH_CHECK();

    // candidate partials: the update kernel writes one per CTA; size the scan the same way
    int64_t n_part = (nl + 255) / 256;
    std::vector<cudaEvent_t> ev;
    if (step_ms_host) {
        ev.resize((size_t)k + 1);
        for (auto& e : ev) MLFFPC_CUDA(cudaEventCreate(&e));
        MLFFPC_CUDA(cudaEventRecord(ev[0], s));
    }

    int status = MLFFPC_OK;
    int prev_parts = 0;
    ProfWindow pw = prof_window("pchol");
    for (int64_t m = 0; m < k && status == MLFFPC_OK; ++m) {
        pw.step(m);
        // (1) candidates -> this rank's best
        if (m == 0) {
            MLFFPC_REQUIRE(n_part <= MLFFPC_MAX_PARTIALS, "pchol_build: n_local too large (%lld rows)", (long long)nl);
            pchol_scan_kernel<<<(unsigned)n_part, 256, 0, s>>>(diag, nl, row0, pos, 0, partials);
            prev_parts = (int)n_part;
        }
        pchol_reduce_kernel<<<1, 256, 0, s>>>(partials, prev_parts, my_cand);
        status = comm_allgather(ctx->comm, my_cand, all_cand, sizeof(Cand), s);
        if (status != MLFFPC_OK) break;
        // (2) pivot, swap, sqrt
        pchol_select_kernel<<<1, 32, 0, s>>>(all_cand, world, m, forced_pivots, diag, row0, nl, index_columns,
                                             pos, piv_idx, piv_val, flag);
        if (forced_pivots && world > 1) {
            set_error("pchol_build: forced_pivots is a single-GPU diagnostic");
            status = MLFFPC_ERR_UNSUPPORTED;
            break;
        }
        // (3) pivot row of the factor, replicated
        if (m > 0) {
            pchol_gather_row_kernel<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(Lt, ld, m, piv_idx, row0, nl, lrow);
            // the owner is only known on the device: replicate with an allreduce-sum of the zero-padded row
            if (world > 1) status = comm_allreduce_sum(ctx->comm, lrow, (size_t)m, s);
        }
        if (status != MLFFPC_OK) break;
        // (4) column pi of A = -K on the local rows
        status = launch_columns_device_col(ctx, piv_idx, col, -1.0, s);
        if (status != MLFFPC_OK) break;
        // (5) Schur update, new factor row, residual diagonal, next candidates
        const int ms = pchol_msplit(nl, m, ctx->num_sms);
        if ((nl + (256 / ms) - 1) / (256 / ms) > MLFFPC_MAX_PARTIALS) {
            set_error("pchol_build: n_local too large for the candidate buffer");
            status = MLFFPC_ERR_INVALID;
            break;
        }
        if (ms == 1) {
            prev_parts = (int)((nl + 255) / 256);
            pchol_update_kernel<1><<<prev_parts, PCHOL_THREADS, 0, s>>>(Lt, ld, m, nl, row0, col, lrow, piv_idx, piv_val, diag, pos, partials);
        } else if (ms == 4) {
            prev_parts = (int)((nl + 63) / 64);
            pchol_update_kernel<4><<<prev_parts, PCHOL_THREADS, 0, s>>>(Lt, ld, m, nl, row0, col, lrow, piv_idx, piv_val, diag, pos, partials);
        } else {
            prev_parts = (int)((nl + 31) / 32);
            pchol_update_kernel<8><<<prev_parts, PCHOL_THREADS, 0, s>>>(Lt, ld, m, nl, row0, col, lrow, piv_idx, piv_val, diag, pos, partials);
        }
        g_launches += (m > 0) ? 4 : 4;  // scan|gather + reduce + select + update (the column kernel counts itself)
        {
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) { status = cuda_fail(e, "pchol step", __FILE__, __LINE__); break; }
        }
        if (step_ms_host) cudaEventRecord(ev[(size_t)m + 1], s);
    }

    pw.end();
    int h_flag = 0;
    if (status == MLFFPC_OK) {
        cudaError_t e = cudaMemcpyAsync(ctx->h_scal, flag, sizeof(int), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) status = cuda_fail(e, "pchol flag readback", __FILE__, __LINE__);
        else h_flag = *(int*)ctx->h_scal;
    }
    if (step_ms_host) {
        if (status == MLFFPC_OK)
            for (int64_t m = 0; m < k; ++m) cudaEventElapsedTime(&step_ms_host[m], ev[(size_t)m], ev[(size_t)m + 1]);
        for (auto& e : ev) cudaEventDestroy(e);
    }
    if (status == MLFFPC_OK && h_flag != 0) {
        set_error("given matrix is not PSD (pivot <= 0 at step %d)", h_flag - 1);
        return MLFFPC_ERR_NOT_PSD;
    }
    return status;
}

}  // extern "C"
