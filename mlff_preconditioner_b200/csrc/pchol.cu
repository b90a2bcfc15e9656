// Greedy diagonal-pivoted partial Cholesky of A = -K with on-the-fly columns.
// Mirrors incomplete_cholesky.pivoted_cholesky (reference solvers/incomplete_cholesky.py:41-80):
//   step m: i* = first argmax of diag over the *current permuted order* index_columns[m:], swap,
//           L[pi,m] = sqrt(diag[pi]);  L[rest,m] = (A[rest,pi] - L[rest,:m] L[pi,:m]) / L[pi,m];
//           diag[rest] -= L[rest,m]^2.
// The factor is kept transposed, Lt[m, r] = L[row0 + r, m], so that the Schur dot streams m
// coalesced rows of length n_local -- the 4 n k^2-byte HBM stream that bounds the plain build.
//
// Blocked variant ("look-ahead", default for k >= 128): the greedy pivot order cannot be known in advance,
// but the next pivots almost always come from the rows with the largest residual diagonal.  A panel of
// LA_C candidate columns is kept with the Schur correction of ALL columns chosen so far already applied
// by one DMMA GEMM   R0[cands, :] = A[cands, :] - L[cands, :m0] L[:, :m0]^T   (a rank-m0 update of a
// LA_C-column block, one pass over the factor for LA_C columns instead of one pass per column).  While
// the arg-max pivot is one of the candidates, a step only applies the columns chosen since the panel was
// built (m - m0 <= a few dozen rows of Lt); when it is not, the panel is rebuilt from the current top
// LA_C residual diagonals (which contain the arg-max by construction).  Same pivots, same factor up to
// summation order; the HBM stream drops from 4 n k^2 bytes to about 4 n k^2 / (average run length).
#include <string.h>

#include <vector>

#include <time.h>

#include "common.cuh"
#include "peer.cuh"

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
                                    int64_t* piv_idx, double* piv_val, int* flag, const int32_t* cslot,
                                    int* slot_out) {
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
    if (cslot) *slot_out = cslot[pi];
    if (!(v > 0.0)) { if (*flag == 0) *flag = (int)(m + 1); v = 1.0; }
    piv_val[0] = sqrt(v);
}

// lrow[m'] = Lt[m', pi - row0] for m' < m on the owner, 0 elsewhere (replicated by an allreduce-sum)
__global__ void pchol_gather_row_kernel(const double* __restrict__ Lt, int64_t ld, int64_t m0, int64_t m,
                                        const int64_t* __restrict__ piv_idx, int64_t row0, int64_t n_local,
                                        double* __restrict__ lrow) {
    const int64_t pi = piv_idx[0];
    const bool owner = (pi >= row0 && pi < row0 + n_local);
    const int64_t t = m0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < m) lrow[t] = owner ? Lt[t * ld + (pi - row0)] : 0.0;
}

// the HBM-bound step: Schur dot over m previous rows of Lt, new row m, diagonal update, candidates
constexpr int PCHOL_THREADS = 256;
template <int MSPLIT>
__global__ void __launch_bounds__(PCHOL_THREADS)
pchol_update_kernel(double* __restrict__ Lt, int64_t ld, int64_t m0, int64_t m, int64_t n_local, int64_t row0,
                    const double* __restrict__ col, const int* __restrict__ slot_ptr, int64_t ld_panel,
                    const double* __restrict__ lrow, const int64_t* __restrict__ piv_idx,
                    const double* __restrict__ piv_val, double* __restrict__ diag,
                    const int32_t* __restrict__ pos, Cand* partials) {
    // col: the pivot column of A on the local rows (m0 = 0), or the candidate panel whose row *slot_ptr holds
    // that column with the first m0 factor columns already applied
    constexpr int COLS = PCHOL_THREADS / MSPLIT;
    __shared__ double red[MSPLIT][COLS];
    __shared__ Cand sm[40];
    const int tc = threadIdx.x % COLS, ts = threadIdx.x / COLS;
    const int64_t r = (int64_t)blockIdx.x * COLS + tc;
    double acc = 0.0;
    if (r < n_local) {
        const double* Lp = Lt + r;
        int64_t q = m0 + ts;
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
            const double* cp = slot_ptr ? (col + (int64_t)(*slot_ptr) * ld_panel) : col;
            l = (cp[r] - acc) / piv_val[0];
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


// ---- look-ahead candidate panel -----------------------------------------------------------------------
constexpr int LA_C = 64;        // candidate columns kept in the panel
constexpr int LA_LCAP = 128;    // per-rank candidate list capacity (top-LA_C plus ties in the threshold bin)
constexpr int LA_MAXE = 4096;   // merge capacity: world * LA_LCAP <= LA_MAXE

__device__ __forceinline__ unsigned long long la_key(double v) {
    return v > 0.0 ? (unsigned long long)__double_as_longlong(v) : 0ull;  // positive doubles order like integers
}

// One CTA: this rank's rows with the largest residual diagonal among the rows not chosen yet (pos >= m).
// Radix select on the upper 33 bits of the value (3 passes of 11 bits), then an index-ordered compaction of
// everything at or above the threshold -> deterministic list of >= min(want, available) entries.
__global__ void __launch_bounds__(1024)
pchol_topc_kernel(const double* __restrict__ diag, int64_t nl, int64_t row0, const int32_t* __restrict__ pos,
                  int64_t m, int want, Cand* __restrict__ list) {
    __shared__ int hist[2048];
    __shared__ unsigned long long s_prefix;
    __shared__ int s_need;
    __shared__ int s_scan[1024];
    const int tid = threadIdx.x;
    unsigned long long prefix = 0;
    int need = want;
    for (int pass = 0; pass < 3; ++pass) {
        const int shift = 52 - 11 * pass;
        for (int b = tid; b < 2048; b += 1024) hist[b] = 0;
        __syncthreads();
        for (int64_t i = tid; i < nl; i += 1024) {
            const unsigned long long key = (pos[row0 + i] >= m) ? la_key(diag[i]) : 0ull;
            if (key != 0ull && (pass == 0 || (key >> (shift + 11)) == prefix))
                atomicAdd(&hist[(int)((key >> shift) & 2047ull)], 1);
        }
        __syncthreads();
        if (tid == 0) {
            int cum = 0, digit = -1;
            for (int b = 2047; b >= 0; --b) {
                if (cum + hist[b] >= need) { digit = b; break; }
                cum += hist[b];
            }
            if (digit < 0) { digit = 0; cum = 0; }  // fewer than `need` rows are left: the threshold admits them all
            s_prefix = (prefix << 11) | (unsigned long long)digit;
            s_need = need - cum;
        }
        __syncthreads();
        prefix = s_prefix;
        need = s_need;
        __syncthreads();
    }
    // selected: key != 0 and (key >> 30) >= prefix ; compaction in row order (each thread owns a contiguous chunk)
    const int64_t L = (nl + 1023) / 1024;
    const int64_t i0 = (int64_t)tid * L, i1 = (i0 + L < nl) ? (i0 + L) : nl;
    int cnt = 0;
    for (int64_t i = i0; i < i1; ++i) {
        const unsigned long long key = (pos[row0 + i] >= m) ? la_key(diag[i]) : 0ull;
        if (key != 0ull && (key >> 30) >= prefix) ++cnt;
    }
    s_scan[tid] = cnt;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int v = (tid >= o) ? s_scan[tid - o] : 0;
        __syncthreads();
        s_scan[tid] += v;
        __syncthreads();
    }
    int off = s_scan[tid] - cnt;
    const int total = s_scan[1023];
    for (int64_t i = i0; i < i1; ++i) {
        const int32_t ps = pos[row0 + i];
        const double v = diag[i];
        const unsigned long long key = (ps >= m) ? la_key(v) : 0ull;
        if (key != 0ull && (key >> 30) >= prefix) {
            if (off < LA_LCAP) { list[off].val = v; list[off].pos = (double)ps; list[off].idx = (double)(row0 + i); list[off].pad = 0.0; }
            ++off;
        }
    }
    for (int t = tid; t < LA_LCAP; t += 1024)
        if (t >= total) { list[t].val = -1e300; list[t].pos = 1e300; list[t].idx = -1.0; list[t].pad = 0.0; }
}

// One CTA: merge the ranks' lists, keep the LA_C best (value descending, position ascending = the pivot
// rule), make sure the current pivot is among them, and rewrite the candidate tables.
//   cand_idx[LA_C] (global rows; padded with the pivot), cslot[n] (row -> slot or -1), slot_out = slot of the pivot
__global__ void __launch_bounds__(1024)
pchol_merge_kernel(const Cand* __restrict__ lists, int n_entries, const int64_t* __restrict__ piv_idx,
                   int64_t* __restrict__ cand_idx, int32_t* __restrict__ cslot, int* __restrict__ slot_out) {
    extern __shared__ double msm[];
    int E = 1;
    while (E < n_entries) E <<= 1;
    double* val = msm;
    double* ps = msm + E;
    double* ix = msm + 2 * E;
    const int tid = threadIdx.x;
    for (int i = tid; i < E; i += 1024) {
        if (i < n_entries) { val[i] = lists[i].val; ps[i] = lists[i].pos; ix[i] = lists[i].idx; }
        else { val[i] = -1e300; ps[i] = 1e300; ix[i] = -1.0; }
    }
    __syncthreads();
    for (int k2 = 2; k2 <= E; k2 <<= 1) {
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < E; i += 1024) {
                const int l = i ^ j;
                if (l > i) {
                    const bool desc = ((i & k2) == 0);
                    const bool i_first = cand_better(val[i], ps[i], val[l], ps[l]);  // i should precede l in descending order
                    if (desc != i_first && !(val[i] == val[l] && ps[i] == ps[l])) {
                        double t;
                        t = val[i]; val[i] = val[l]; val[l] = t;
                        t = ps[i]; ps[i] = ps[l]; ps[l] = t;
                        t = ix[i]; ix[i] = ix[l]; ix[l] = t;
                    }
                }
            }
            __syncthreads();
        }
    }
    // old candidates out
    if (tid < LA_C) {
        const int64_t old = cand_idx[tid];
        if (old >= 0) cslot[old] = -1;
    }
    __syncthreads();
    __shared__ int s_found;
    if (tid == 0) s_found = 0;
    __syncthreads();
    const int64_t pi = piv_idx[0];
    if (tid < LA_C && tid < E && val[tid] > 0.0 && (int64_t)ix[tid] == pi) s_found = 1;
    __syncthreads();
    if (tid < LA_C) {
        int64_t g = (tid < E && val[tid] > 0.0) ? (int64_t)ix[tid] : -1;
        if (!s_found && tid == LA_C - 1) g = pi;  // cannot happen unless > LA_LCAP rows tie; keep the pivot reachable
        const bool real = g >= 0;
        if (!real) g = pi;                        // padding (fewer than LA_C rows left): duplicates of the pivot
        cand_idx[tid] = g;
        if (real) cslot[g] = tid;
    }
    __syncthreads();
    if (tid == 0) *slot_out = cslot[pi];
}

// Lc[j, t] = L[cand_j, t] for t < m on the owner rank, 0 elsewhere (summed over ranks by the caller)
__global__ void pchol_gather_cand_rows_kernel(const double* __restrict__ Lt, int64_t ld, int64_t m, int64_t ld_lc,
                                              const int64_t* __restrict__ cand_idx, int64_t row0, int64_t n_local,
                                              double* __restrict__ Lc) {
    const int j = blockIdx.y;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ld_lc) return;
    const int64_t g = cand_idx[j];
    const bool owner = (g >= row0 && g < row0 + n_local);
    Lc[(int64_t)j * ld_lc + t] = (owner && t < m) ? Lt[t * ld + (g - row0)] : 0.0;
}

// ---- host-free stepping of the look-ahead build ---------------------------------------------------------------
// A step is two launches and (sharded) one collective, and the host does not wait for it:
//   prepare (1 CTA)  apply the previous step's swap to index_columns / pos, reduce the candidate partials to this rank's
//                    best row, and put {candidate, its factor row L[cand, m0:m]} into the send buffer
//   allgather        LA_MSG doubles per rank (replaces the 32-byte allgather + the allreduce of the pivot row)
//   update           every CTA picks the winner among the ranks' candidates; if its column is in the panel it runs the
//                    Schur update for the factor columns m0 .. m-1, else (a miss, or a pivot <= 0) it sets `stalled`
//                    and every later launch returns at once until the host has rebuilt the panel.
// The host launches LA_CHUNK steps, then reads the 64-byte state once.
constexpr int LA_MSG = 4 + LA_C;   // candidate (val, pos, idx, pad) + at most LA_C entries of its factor row
constexpr int LA_CHUNK = 8;
static_assert(LA_MSG <= PEER_MSG_DOUBLES, "peer message slot too small");

struct LaState {
    int stalled;        // 1: a step could not run (panel miss or non-PSD pivot); cleared by the host
    int flag;           // != 0: pivot <= 0 at step flag - 1 (incomplete_cholesky.py:62)
    int slot;           // scratch of the merge kernel
    int pad0;
    long long stall_m;  // the step that stalled
    long long miss_pi;  // its pivot row (input of the panel rebuild)
    long long pend_m;   // swap of step pend_m (pivot pend_pi) not yet applied to index_columns / pos; -1: none
    long long pend_pi;
    long long done_m;   // steps completed
    long long pad1;
    // The device owns the loop: every launch of the two step kernels has the SAME arguments (so a chunk of steps is one
    // CUDA-graph launch); the prepare kernel latches the step number and the peer epoch / slot for its update kernel.
    long long k_total;  // steps to do; launches beyond it return at once
    long long m0;       // factor columns already folded into the candidate panel (host writes it at a rebuild)
    long long cur_m;    // step of the update kernel that follows
    unsigned long long epoch;   // peer channel PEER_CH_MSG: last epoch used
    unsigned long long uses;    //   ... and rounds so far (slot parity)
    unsigned long long cur_epoch;
    long long cur_parity;
    long long pad2;
};

__device__ __forceinline__ void la_apply_pending(LaState* st, int64_t* index_columns, int32_t* pos) {
    if (st->pend_m >= 0) {
        const int64_t mp = st->pend_m, pi = st->pend_pi;
        const int32_t i_argmax = pos[pi];
        const int64_t e = index_columns[mp];
        index_columns[mp] = pi;
        index_columns[i_argmax] = e;
        pos[pi] = (int32_t)mp;
        pos[e] = i_argmax;
        st->pend_m = -1;
    }
}

// peer mode (pv_on): the message goes straight into every rank's slot `parity` and channel PEER_CH_MSG is raised --
// no collective call between the two kernels of a step
__global__ void pchol_la_prepare_kernel(const Cand* __restrict__ partials, int count, const double* __restrict__ Lt,
                                        int64_t ld, int64_t row0, int64_t* index_columns,
                                        int32_t* pos, LaState* st, double* __restrict__ send, const PeerView pv,
                                        int pv_on) {
    if (st->stalled) return;
    if (st->done_m >= st->k_total) {  // surplus launch of the last chunk: tell the update kernel there is no step
        if (threadIdx.x == 0) st->cur_m = -1;
        return;
    }
    __shared__ Cand sm[40];
    __shared__ long long s_par;
    __shared__ unsigned long long s_ep;
    const int64_t m = st->done_m, m0 = st->m0;
    if (threadIdx.x == 0) {
        la_apply_pending(st, index_columns, pos);
        st->cur_m = m;
        s_ep = ++st->epoch;
        s_par = (long long)(st->uses++ & 1ull);
        st->cur_epoch = s_ep;
        st->cur_parity = s_par;
    }
    __syncthreads();
    const int parity = (int)s_par;
    const uint64_t epoch = s_ep;
    double v = -1e300, p = 1e300, i = -1.0;
    for (int t = threadIdx.x; t < count; t += blockDim.x) {
        const Cand c = partials[t];
        if (cand_better(c.val, c.pos, v, p)) { v = c.val; p = c.pos; i = c.idx; }
    }
    block_argmax(v, p, i, sm);
    const Cand best = sm[0];
    if (threadIdx.x == 0) { send[0] = best.val; send[1] = best.pos; send[2] = best.idx; send[3] = 0.0; }
    const int64_t g = (int64_t)best.idx;
    for (int64_t t = m0 + threadIdx.x; t < m; t += blockDim.x)
        send[4 + (t - m0)] = (g >= 0) ? Lt[t * ld + (g - row0)] : 0.0;
    if (pv_on) {
        __syncthreads();
        const int len = 4 + (int)(m - m0);
        for (int r = 0; r < pv.world; ++r) {
            double* dst = pv.msg(r, parity, pv.rank);
            for (int t = threadIdx.x; t < len; t += blockDim.x) dst[t] = send[t];
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) peer_signal_all(pv, PEER_CH_MSG, epoch);
    }
}

__global__ void pchol_la_flush_kernel(LaState* st, int64_t* index_columns, int32_t* pos) {
    if (threadIdx.x == 0 && blockIdx.x == 0) la_apply_pending(st, index_columns, pos);
}

template <int MSPLIT>
__global__ void __launch_bounds__(PCHOL_THREADS)
pchol_la_update_kernel(double* __restrict__ Lt, int64_t ld, int64_t n_local, int64_t row0,
                       const double* __restrict__ panel, int64_t ld_panel, const double* __restrict__ gathered, int world,
                       const int32_t* __restrict__ cslot, double* __restrict__ diag, const int32_t* __restrict__ pos,
                       const int64_t* __restrict__ index_columns, Cand* partials, LaState* st, const PeerView pv,
                       int pv_on) {
    constexpr int COLS = PCHOL_THREADS / MSPLIT;
    // step number, slot and epoch were latched by the prepare kernel (done_m itself is advanced by block 0 of THIS
    // kernel, so it must not be read here); cur_m < 0: there was no step to prepare
    const int64_t m = st->cur_m, m0 = st->m0;
    if (m < 0) return;
    const int parity = (int)st->cur_parity;
    const uint64_t epoch = st->cur_epoch;
    __shared__ double red[MSPLIT][COLS];
    __shared__ Cand sm[40];
    __shared__ double lrow[LA_C];
    __shared__ double s_val;
    __shared__ long long s_pi;
    __shared__ int s_slot, s_rank, s_stalled;
    // the ranks' messages: the allgather buffer, or (peer mode) this rank's message slots once every rank has raised
    // the channel -- skipped while stalled, when no rank has pushed
    const int64_t msg_stride = pv_on ? PEER_MSG_DOUBLES : LA_MSG;
    if (pv_on) gathered = pv.msg(pv.rank, parity, 0);
    if (threadIdx.x == 0) {
        s_stalled = st->stalled;
        if (pv_on && !s_stalled) peer_wait_all(pv, PEER_CH_MSG, epoch);
        double v = -1e300, p = 1e300, i = -1.0;
        int wr = 0;
        for (int r = 0; r < world && !s_stalled; ++r) {
            const double* c = gathered + (int64_t)r * msg_stride;
            const double c0 = peer_ld(c), c1 = peer_ld(c + 1), c2 = peer_ld(c + 2);
            if (cand_better(c0, c1, v, p)) { v = c0; p = c1; i = c2; wr = r; }
        }
        s_val = v; s_pi = (long long)i; s_rank = wr;
        s_slot = (i >= 0.0) ? cslot[(int64_t)i] : -1;
    }
    __syncthreads();
    if (s_stalled) return;
    const int64_t pi = s_pi;
    if (s_slot < 0 || !(s_val > 0.0)) {  // panel miss / non-PSD pivot: nothing is modified, the host takes over
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            st->stall_m = m;
            st->miss_pi = pi;
            if (!(s_val > 0.0) && st->flag == 0) st->flag = (int)(m + 1);
            __threadfence();
            st->stalled = 1;
        }
        return;
    }
    for (int t = threadIdx.x; t < (int)(m - m0); t += blockDim.x) lrow[t] = peer_ld(gathered + (int64_t)s_rank * msg_stride + 4 + t);
    __syncthreads();
    const double piv = sqrt(s_val);
    const int tc = threadIdx.x % COLS, ts = threadIdx.x / COLS;
    const int64_t r = (int64_t)blockIdx.x * COLS + tc;
    double acc = 0.0;
    if (r < n_local) {
        const double* Lp = Lt + r;
        int64_t q = m0 + ts;
        for (; q + 7 * MSPLIT < m; q += 8 * MSPLIT) {
            double t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) t[u] = __ldcs(Lp + (q + u * MSPLIT) * ld);
#pragma unroll
            for (int u = 0; u < 8; ++u) acc = fma(t[u], lrow[q + u * MSPLIT - m0], acc);
        }
        for (; q < m; q += MSPLIT) acc = fma(__ldcs(Lp + q * ld), lrow[q - m0], acc);
    }
    if (MSPLIT > 1) {
        red[ts][tc] = acc;
        __syncthreads();
        if (ts == 0) {
#pragma unroll
            for (int s2 = 1; s2 < MSPLIT; ++s2) acc += red[s2][tc];
        }
    }
    double v = -1e300, p = 1e300, i = -1.0;
    if (ts == 0 && r < n_local) {
        const int64_t g = row0 + r;
        // positions after this step's swap (applied to the arrays by the next prepare kernel): the pivot goes to m,
        // the row that sat at m goes to where the pivot was
        const int64_t e = index_columns[m];
        int32_t ps = pos[g];
        if (g == e && g != pi) ps = pos[pi];
        double l;
        if (g == pi) {
            l = piv;
        } else if (ps > m) {  // still a candidate row: i_pi = index_columns[m+1:]
            const double* cp = panel + (int64_t)s_slot * ld_panel;
            l = (cp[r] - acc) / piv;
            const double dn = diag[r] - l * l;
            diag[r] = dn;
            v = dn; p = (double)ps; i = (double)g;
        } else {
            l = 0.0;
        }
        Lt[m * ld + r] = l;
    }
    block_argmax(v, p, i, sm);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = sm[0];
        if (blockIdx.x == 0) { st->pend_m = m; st->pend_pi = pi; st->done_m = m + 1; }
    }
}

static int pchol_msplit(int64_t n_local, int64_t m, int num_sms) {
    const int64_t want = (int64_t)num_sms * 1024;
    if (n_local >= want || m < 64) return 1;
    if (n_local * 4 >= want || m < 256) return 4;
    return 8;
}

struct PcholWs {
    int64_t off_col, off_lrow, off_pos, off_cands, off_gathered, off_piv, off_cand_idx, off_cslot, off_panel, off_lc,
        off_list, off_lists, off_state, off_send, off_recv, total;
};
static bool pchol_lookahead_enabled(const mlffpc_ctx* c, int64_t k, bool forced) {
    return c->pchol_lookahead && !forced && k >= 128 && (int64_t)c->comm.world * LA_LCAP <= LA_MAXE &&
           c->n_local() >= 4 * LA_C;
}
static PcholWs pchol_layout(const mlffpc_ctx* c, int64_t k) {
    auto up = [](int64_t x) { return (x + 255) / 256 * 256; };
    PcholWs w;
    int64_t o = 0;
    w.off_col = o; o = up(o + c->n_local() * 8);
    w.off_lrow = o; o = up(o + (k + 1) * 8);
    w.off_pos = o; o = up(o + c->n * 4);
    w.off_cands = o; o = up(o + 64);                 // this rank's candidate
    w.off_gathered = o; o = up(o + 64 * 1024);        // all ranks' candidates (<= 1024 ranks)
    w.off_piv = o; o = up(o + 64);                    // piv_idx (int64), piv_val (double), flag (int), slot (int)
    // look-ahead panel (allocated whenever it could be used; forced_pivots runs simply ignore it)
    const bool la = pchol_lookahead_enabled(c, k, false);
    w.off_cand_idx = o; o = up(o + (la ? LA_C * 8 : 0));
    w.off_cslot = o;    o = up(o + (la ? c->n * 4 : 0));
    w.off_panel = o;    o = up(o + (la ? (int64_t)LA_C * c->n_local() * 8 : 0));
    w.off_lc = o;       o = up(o + (la ? (int64_t)LA_C * (k + 2) * 8 : 0));
    w.off_list = o;     o = up(o + (la ? LA_LCAP * (int64_t)sizeof(Cand) : 0));
    w.off_lists = o;    o = up(o + (la ? (int64_t)c->comm.world * LA_LCAP * (int64_t)sizeof(Cand) : 0));
    w.off_state = o;    o = up(o + 256);
    w.off_send = o;     o = up(o + LA_MSG * 8);
    w.off_recv = o;     o = up(o + (int64_t)c->comm.world * LA_MSG * 8);
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

static double pchol_wall_ms() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
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

    int* slot_dev = (int*)(base + w.off_piv + 24);
    const bool la = pchol_lookahead_enabled(ctx, k, forced_pivots != nullptr);
    int64_t* cand_idx = (int64_t*)(base + w.off_cand_idx);
    int32_t* cslot = (int32_t*)(base + w.off_cslot);
    double* panel = (double*)(base + w.off_panel);
    double* Lc = (double*)(base + w.off_lc);
    Cand* my_list = (Cand*)(base + w.off_list);
    Cand* all_lists = (Cand*)(base + w.off_lists);

    MLFFPC_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), s));
    pchol_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, index_columns, pos);
    MLFFPC_LAUNCH_CHECK();
    if (la) {
        MLFFPC_CUDA(cudaMemsetAsync(cslot, 0xff, (size_t)n * 4, s));       // -1
        MLFFPC_CUDA(cudaMemsetAsync(cand_idx, 0xff, (size_t)LA_C * 8, s));  // -1
        MLFFPC_CUDA(cudaFuncSetAttribute(pchol_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         3 * LA_MAXE * (int)sizeof(double)));
    }

    // candidate partials: the update kernel writes one per CTA; size the scan the same way
    int64_t n_part = (nl + 255) / 256;
    std::vector<cudaEvent_t> ev;
    if (step_ms_host && !la) {
        ev.resize((size_t)k + 1);
        for (auto& e : ev) MLFFPC_CUDA(cudaEventCreate(&e));
        MLFFPC_CUDA(cudaEventRecord(ev[0], s));
    }

    int status = MLFFPC_OK;
    int prev_parts = 0;
    int64_t m0 = 0;        // factor columns already folded into the candidate panel
    int64_t refills = 0;
    ProfWindow pw = prof_window("pchol");
    int h_flag_la = 0;
    if (la && k > 0) {
        // ---- look-ahead build, host-free stepping (kernels above) ----------------------------------------------
        LaState* st = (LaState*)(base + w.off_state);
        double* send = (double*)(base + w.off_send);
        double* recv = (world > 1) ? (double*)(base + w.off_recv) : send;
        LaState* h_st = (LaState*)(ctx->h_scal + 24);  // pinned; sizeof(LaState) <= 160 bytes
        long long* h_val = (long long*)(ctx->h_scal + 48);  // pinned staging for single-field updates
        const int pv_on = (peer_on(ctx) && ctx->peer_pivots) ? 1 : 0;
        PeerView pv;
        if (pv_on) pv = ctx->peer->view;
        else memset(&pv, 0, sizeof(pv));
        cudaGraphExec_t gexec = nullptr;
        cudaEvent_t ev_t[2] = {nullptr, nullptr};
        do {
            LaState init;
            memset(&init, 0, sizeof(init));
            init.pend_m = -1;
            init.cur_m = -1;
            init.k_total = k;
            if (pv_on) { init.epoch = ctx->peer->epoch; init.uses = ctx->peer->uses[PEER_CH_MSG]; }
            *h_st = init;
            cudaError_t e = cudaMemcpyAsync(st, h_st, sizeof(LaState), cudaMemcpyHostToDevice, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) { status = cuda_fail(e, "pchol state init", __FILE__, __LINE__); break; }
            if (n_part > MLFFPC_MAX_PARTIALS) { set_error("pchol_build: n_local too large (%lld rows)", (long long)nl); status = MLFFPC_ERR_INVALID; break; }
            pchol_scan_kernel<<<(unsigned)n_part, 256, 0, s>>>(diag, nl, row0, pos, 0, partials);
            ++g_launches;
            // One chunk = LA_CHUNK steps = 2 LA_CHUNK launches with identical arguments (the device owns the step
            // counter).  Without a library collective between the two kernels (one GPU, or peer-memory messages) the
            // chunk is captured once and replayed as ONE graph launch; with NCCL the kernels are launched one by one.
            auto launch_chunk = [&]() -> int {
                for (int c = 0; c < LA_CHUNK; ++c) {
                    pchol_la_prepare_kernel<<<1, 256, 0, s>>>(partials, (int)n_part, Lt, ld, row0, index_columns, pos, st, send, pv, pv_on);
                    if (world > 1 && !pv_on) MLFFPC_TRY(comm_allgather(ctx->comm, send, recv, LA_MSG * sizeof(double), s));
                    // one thread per row, always: a step applies at most LA_C factor columns, and the candidate partials
                    // must keep ONE layout -- a stalled (no-op) launch leaves the previous step's partials in place
                    pchol_la_update_kernel<1><<<(unsigned)n_part, PCHOL_THREADS, 0, s>>>(Lt, ld, nl, row0, panel, nl, recv, world, cslot, diag, pos,
                                                                                         index_columns, partials, st, pv, pv_on);
                }
                return MLFFPC_OK;
            };
            if ((world == 1 || pv_on) && ctx->pchol_graph) {
                cudaGraph_t graph = nullptr;
                e = cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed);
                if (e == cudaSuccess) {
                    status = launch_chunk();
                    e = cudaStreamEndCapture(s, &graph);
                    if (e == cudaSuccess && status == MLFFPC_OK) e = cudaGraphInstantiate(&gexec, graph, 0);
                    if (graph) cudaGraphDestroy(graph);
                }
                if (e != cudaSuccess || status != MLFFPC_OK) {  // no graph: fall back to plain launches
                    cudaGetLastError();
                    gexec = nullptr;
                    status = MLFFPC_OK;
                }
            }
            if (step_ms_host) {
                e = cudaEventCreate(&ev_t[0]);
                if (e == cudaSuccess) e = cudaEventCreate(&ev_t[1]);
                if (e == cudaSuccess) e = cudaEventRecord(ev_t[0], s);
                if (e != cudaSuccess) { status = cuda_fail(e, "pchol timing events", __FILE__, __LINE__); break; }
            }
            int64_t done = 0;
            int tcur = 0;
            float carry_ms = 0.f;
            // MLFFPC_TIMING=1: host wall time of the stepping chunks against the panel rebuilds (both end in a sync)
            const char* tenv = getenv("MLFFPC_TIMING");
            const bool t_on = tenv && tenv[0] == '1';
            double t_rebuild = 0.0, t_begin = t_on ? pchol_wall_ms() : 0.0;
            int64_t n_chunks = 0;
            while (done < k && status == MLFFPC_OK) {
                ++n_chunks;
                pw.step(done);
                if (gexec) {
                    e = cudaGraphLaunch(gexec, s);
                    if (e != cudaSuccess) { status = cuda_fail(e, "pchol chunk graph", __FILE__, __LINE__); break; }
                } else {
                    status = launch_chunk();
                    if (status != MLFFPC_OK) break;
                }
                g_launches += 2 * LA_CHUNK;
                cudaError_t e2 = cudaMemcpyAsync(h_st, st, sizeof(LaState), cudaMemcpyDeviceToHost, s);
                if (e2 == cudaSuccess && step_ms_host) e2 = cudaEventRecord(ev_t[tcur ^ 1], s);
                if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(s);
                if (e2 == cudaSuccess) e2 = cudaGetLastError();
                if (e2 != cudaSuccess) { status = cuda_fail(e2, "pchol chunk", __FILE__, __LINE__); break; }
                const int64_t new_done = h_st->done_m;
                if (step_ms_host) {  // the chunk's time (incl. a preceding panel rebuild), spread over the steps it completed
                    float ms = 0.f;
                    cudaEventElapsedTime(&ms, ev_t[tcur], ev_t[tcur ^ 1]);
                    tcur ^= 1;
                    carry_ms += ms;
                    if (new_done > done) {
                        for (int64_t q = done; q < new_done; ++q) step_ms_host[q] = carry_ms / (float)(new_done - done);
                        carry_ms = 0.f;
                    }
                }
                done = new_done;
                if (h_st->flag != 0) { h_flag_la = h_st->flag; break; }
                if (!h_st->stalled) continue;
                // (2b) panel rebuild at the stalled step: top rows by residual diagonal -> merged candidate list ->
                // columns of A -> fold in L[:, :m]
                const int64_t m = h_st->stall_m;
                ++refills;
                const double t_r0 = t_on ? pchol_wall_ms() : 0.0;
                pchol_topc_kernel<<<1, 1024, 0, s>>>(diag, nl, row0, pos, m, LA_C, my_list);
                status = comm_allgather(ctx->comm, my_list, all_lists, LA_LCAP * sizeof(Cand), s);
                if (status != MLFFPC_OK) break;
                int E = 1;
                while (E < world * LA_LCAP) E <<= 1;
                pchol_merge_kernel<<<1, 1024, 3 * E * sizeof(double), s>>>(all_lists, world * LA_LCAP, (const int64_t*)&st->miss_pi,
                                                                         cand_idx, cslot, &st->slot);
                g_launches += 2;
                status = mlffpc_kernel_columns(ctx, cand_idx, LA_C, panel, nl, -1.0, nullptr, 0, (void*)s);
                if (status != MLFFPC_OK) break;
                if (m > 0) {
                    const int64_t ld_lc = (m + 1) & ~(int64_t)1;  // even pitch: the GEMM stages 16-byte copies
                    pchol_gather_cand_rows_kernel<<<dim3((unsigned)((ld_lc + 255) / 256), LA_C), 256, 0, s>>>(
                        Lt, ld, m, ld_lc, cand_idx, row0, nl, Lc);
                    ++g_launches;
                    status = comm_allreduce_sum(ctx->comm, Lc, (size_t)(LA_C * ld_lc), s);
                    if (status != MLFFPC_OK) break;
                    status = dgemm(false, LA_C, nl, m, -1.0, Lc, ld_lc, Lt, ld, 1.0, panel, nl, false, s);
                    if (status != MLFFPC_OK) break;
                }
                h_val[0] = m;
                e2 = cudaMemcpyAsync(&st->m0, h_val, sizeof(long long), cudaMemcpyHostToDevice, s);
                if (e2 == cudaSuccess) e2 = cudaMemsetAsync(&st->stalled, 0, sizeof(int), s);
                if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(s);  // h_val is reused by the next rebuild
                if (e2 != cudaSuccess) { status = cuda_fail(e2, "pchol clear stall", __FILE__, __LINE__); break; }
                if (t_on) t_rebuild += pchol_wall_ms() - t_r0;
            }
            if (t_on && ctx->comm.rank == 0) {
                const double t_all = pchol_wall_ms() - t_begin;
                fprintf(stderr, "[mlffpc timing] pchol look-ahead: %lld steps in %lld chunks %.3f ms, %lld panel rebuilds %.3f ms\n",
                        (long long)done, (long long)n_chunks, t_all - t_rebuild, (long long)refills, t_rebuild);
            }
            if (status == MLFFPC_OK && h_flag_la == 0) {
                pchol_la_flush_kernel<<<1, 32, 0, s>>>(st, index_columns, pos);
                ++g_launches;
                const cudaError_t e3 = cudaStreamSynchronize(s);
                if (e3 != cudaSuccess) status = cuda_fail(e3, "pchol end", __FILE__, __LINE__);
            }
            if (pv_on) {  // hand the channel's counters back to the host-side bookkeeping (identical on every rank)
                ctx->peer->epoch = h_st->epoch > ctx->peer->epoch ? h_st->epoch : ctx->peer->epoch;
                ctx->peer->uses[PEER_CH_MSG] = h_st->uses;
            }
        } while (0);
        if (gexec) cudaGraphExecDestroy(gexec);
        for (auto& evq : ev_t)
            if (evq) cudaEventDestroy(evq);
    }
    for (int64_t m = 0; !la && m < k && status == MLFFPC_OK; ++m) {
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
                                             pos, piv_idx, piv_val, flag, nullptr, slot_dev);
        if (forced_pivots && world > 1) {
            set_error("pchol_build: forced_pivots is a single-GPU diagnostic");
            status = MLFFPC_ERR_UNSUPPORTED;
            break;
        }
        g_launches += 3;
        // (3) pivot row of the factor (columns 0 .. m-1; m0 stays 0 in the plain build), replicated
        if (m > m0) {
            pchol_gather_row_kernel<<<(unsigned)((m - m0 + 255) / 256), 256, 0, s>>>(Lt, ld, m0, m, piv_idx, row0, nl, lrow);
            ++g_launches;
            // the owner is only known on the device: replicate with an allreduce-sum of the zero-padded row
            if (world > 1) status = comm_allreduce_sum(ctx->comm, lrow + m0, (size_t)(m - m0), s);
        }
        if (status != MLFFPC_OK) break;
        // (4) column pi of A = -K on the local rows
        status = launch_columns_device_col(ctx, piv_idx, col, -1.0, s);
        if (status != MLFFPC_OK) break;
        // (5) Schur update over the factor columns 0 .. m-1, new factor row, residual diagonal, next candidates
        const double* colsrc = col;
        const int* slot_ptr = nullptr;
        const int ms = pchol_msplit(nl, m - m0, ctx->num_sms);
        if ((nl + (256 / ms) - 1) / (256 / ms) > MLFFPC_MAX_PARTIALS) {
            set_error("pchol_build: n_local too large for the candidate buffer");
            status = MLFFPC_ERR_INVALID;
            break;
        }
        if (ms == 1) {
            prev_parts = (int)((nl + 255) / 256);
            pchol_update_kernel<1><<<prev_parts, PCHOL_THREADS, 0, s>>>(Lt, ld, m0, m, nl, row0, colsrc, slot_ptr, nl, lrow, piv_idx, piv_val, diag, pos, partials);
        } else if (ms == 4) {
            prev_parts = (int)((nl + 63) / 64);
            pchol_update_kernel<4><<<prev_parts, PCHOL_THREADS, 0, s>>>(Lt, ld, m0, m, nl, row0, colsrc, slot_ptr, nl, lrow, piv_idx, piv_val, diag, pos, partials);
        } else {
            prev_parts = (int)((nl + 31) / 32);
            pchol_update_kernel<8><<<prev_parts, PCHOL_THREADS, 0, s>>>(Lt, ld, m0, m, nl, row0, colsrc, slot_ptr, nl, lrow, piv_idx, piv_val, diag, pos, partials);
        }
        ++g_launches;
        {
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) { status = cuda_fail(e, "pchol step", __FILE__, __LINE__); break; }
        }
        if (step_ms_host) cudaEventRecord(ev[(size_t)m + 1], s);
    }
    ctx->last_pchol_refills = refills;

    pw.end();
    int h_flag = h_flag_la;
    if (status == MLFFPC_OK && !la) {
        cudaError_t e = cudaMemcpyAsync(ctx->h_scal, flag, sizeof(int), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) status = cuda_fail(e, "pchol flag readback", __FILE__, __LINE__);
        else h_flag = *(int*)ctx->h_scal;
    }
    if (step_ms_host && !la) {
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
