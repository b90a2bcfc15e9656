// Preconditioned CG driver: device-resident scalars, fused vector kernels, convergence decided on the device.
//
// Recurrence and stopping rule restate scipy 1.7.3 sparse.linalg.cg(tol=, atol=None) as called by the
// reference (solvers/iterative_solver.py:995-1005): atol = tol*||b||, probe ||A x0 - b|| <= tol first,
// on the first ||r|| <= atol after iteration 1 recompute r = b - A x once and re-test.
//
// The host never waits for a scalar inside the loop.  Iterations are launched in short batches; the kernels
// that change x, r, p or the scalars return at once when the device-side state says "frozen", which the first
// kernel to see ||r||^2 <= atol^2 sets together with the iteration number.  The host reads the 32-byte state one
// batch late (pinned memory, an event per batch), so the GPU never idles on it; the iterations launched past the
// stopping point are no-ops for the state (their operator / preconditioner kernels still run, a few milliseconds).
// ||r||^2 of iteration j travels with rho of iteration j + 1 in ONE two-element allreduce, so a sharded run
// issues two scalar collectives per iteration instead of three; the test of iteration j therefore happens in the
// p-update kernel of iteration j + 1, before anything is modified -- x, r, p and the iteration count are
// exactly those of the legacy loop.
#include <math.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "peer.cuh"

namespace mlffpc {

constexpr int VEC_THREADS = 256;
constexpr int VEC_MAX_BLOCKS = 1024;

// scalar slots in ctx->scal used by mlffpc_dot
enum { S_TMP = 4, S_COUNTER = 8 /* unsigned */ };

// device-side loop state (in the workspace; mirrored to pinned host memory once per batch)
struct PcgState {
    int frozen;        // 1: the stopping test fired (or NaN) -- vector kernels are no-ops until the host clears it
    int conv_iter;     // iteration whose residual passed the test
    int nan_flag;      // residual became NaN at conv_iter
    int last_iter;     // last iteration whose ||r||^2 has been recorded
    double last_rr;    // ||r||^2 of last_iter (global)
    double pad;
};
// red[0] = rho (r.z), red[1] = ||r||^2 of the previous update, red[2] = p.q, red[3] = rho of the previous iteration,
// red[6] = rho of this iteration summed over ranks (written by the p-update, read by the x,r-update)

// One fused scalar exchange over peer memory (peer.cuh): the producing kernel pushes into slot `parity` of channel `ch`,
// the consuming kernel waits for epoch and adds the ranks' slots in rank order.  on = 0: NCCL has already summed red[].
struct PeerXchg {
    PeerView pv;
    int on, ch, parity;
    uint64_t epoch;
};
__device__ __forceinline__ void peer_push_scalars(const PeerXchg& x, double v0, double v1) {
    for (int r = 0; r < x.pv.world; ++r) {
        double* dst = x.pv.scal(r, x.parity, x.pv.rank);
        dst[0] = v0;
        dst[1] = v1;
    }
    peer_signal_all(x.pv, x.ch, x.epoch);
}
// one thread: wait for every rank's push, return the rank-ordered sums
__device__ __forceinline__ void peer_sum_scalars(const PeerXchg& x, double& v0, double& v1) {
    peer_wait_all(x.pv, x.ch, x.epoch);
    double a = 0.0, b = 0.0;
    for (int r = 0; r < x.pv.world; ++r) {
        const double* src = x.pv.scal(x.pv.rank, x.parity, r);
        a += peer_ld(src);
        b += peer_ld(src + 1);
    }
    v0 = a;
    v1 = b;
}

// deterministic grid reduction: per-block partials, the last block to finish sums them in index order
__device__ __forceinline__ bool grid_reduce(double v, double* partials, unsigned* counter, double* result) {
    __shared__ double sm[40];
    __shared__ bool is_last;
    v = block_sum(v, sm);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = v;
        __threadfence();
        const unsigned done = atomicAdd(counter, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    double t = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) t += ((volatile double*)partials)[i];
    t = block_sum(t, sm);
    if (threadIdx.x == 0) {
        *result = t;
        *counter = 0u;
    }
    return threadIdx.x == 0;
}

// out = sum a[i] b[i]  (state == NULL: unconditional; otherwise the store is skipped while frozen).  With a peer
// exchange the last block also pushes (result, *extra) to every rank -- the collective is part of this kernel.
__global__ void dot_kernel(const double* __restrict__ a, const double* __restrict__ b, int64_t n,
                           double* partials, unsigned* counter, double* out, const PcgState* state,
                           const PeerXchg px, const double* extra) {
    if (state && state->frozen) return;
    double v = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        v = fma(a[i], b[i], v);
    __shared__ double res;
    if (grid_reduce(v, partials, counter, &res)) {
        *out = res;
        if (px.on) peer_push_scalars(px, res, extra ? *extra : 0.0);
    }
}

// Stopping test of iteration it - 1 (its ||r||^2 = red[1], global after the allreduce) and, unless it fires,
// p = z + (rho / rho_prev) p  (it == 1: p = z).  only_check = 1: the test alone (after the last iteration).
// Peer mode: (rho, ||r||^2) are combined from the ranks' pushes here (pa), and every thread stores its entries of p
// into ALL ranks' replicated search direction (pv.p_full) -- the allgather is part of this kernel; the last block to
// finish raises channel PEER_CH_P (pp).
__global__ void update_p_kernel(const double* __restrict__ z, double* __restrict__ p, int64_t n, double* red,
                                const double* red_test, PcgState* state, double atol2, int64_t it, int only_check,
                                double* __restrict__ hist, const PeerXchg pa, const PeerXchg pp, int64_t row0) {
    const bool was_frozen = state->frozen != 0;
    if (was_frozen) return;   // before any wait: the producers skipped their pushes, on every rank alike
    __shared__ double s_rho, s_rr;
    if (threadIdx.x == 0) {
        if (pa.on) peer_sum_scalars(pa, s_rho, s_rr);
        else { s_rho = red[0]; s_rr = red_test[1]; }
    }
    __syncthreads();
    const double rho = s_rho;
    const double rr = s_rr;
    if (blockIdx.x == 0 && threadIdx.x == 0 && !only_check) red[6] = rho;
    const bool test = it > 1;  // nothing to test before the first update
    const bool fire = test && (!(rr == rr) || rr <= atol2);
    if (blockIdx.x == 0 && threadIdx.x == 0 && !was_frozen && test) {
        state->last_iter = (int)(it - 1);
        state->last_rr = rr;
        if (hist) hist[it - 1] = rr;
        if (fire) {
            state->conv_iter = (int)(it - 1);
            state->nan_flag = (rr == rr) ? 0 : 1;
            __threadfence();
            state->frozen = 1;
        }
    }
    if (fire || only_check) return;
    const double beta = (it == 1) ? 0.0 : (rho / red[3]);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = (it == 1) ? z[i] : fma(beta, p[i], z[i]);
        p[i] = v;
        if (pp.on)
            for (int r = 0; r < pp.pv.world; ++r)
                if (r != pp.pv.rank) pp.pv.p_full(r)[row0 + i] = v;
    }
    if (pp.on) {
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned* cnt = pp.pv.counter(0);
            const unsigned done = atomicAdd(cnt, 1u);
            if (done == gridDim.x - 1) {
                *cnt = 0u;
                peer_signal_all(pp.pv, PEER_CH_P, pp.epoch);
            }
        }
    }
}

// alpha = rho / (p.q); x += alpha p; r -= alpha q; red[1] = sum r^2 (local); red[3] = rho
__global__ void update_xr_kernel(double* __restrict__ x, double* __restrict__ r, const double* __restrict__ p,
                                 const double* __restrict__ q, int64_t n, double* red, const PcgState* state,
                                 double* partials, unsigned* counter, const PeerXchg pb) {
    if (state->frozen) return;
    __shared__ double s_pq;
    if (threadIdx.x == 0) {
        if (pb.on) { double dummy; peer_sum_scalars(pb, s_pq, dummy); }
        else s_pq = red[2];
    }
    __syncthreads();
    const double rho = red[6];
    const double alpha = rho / s_pq;
    double v = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        x[i] = fma(alpha, p[i], x[i]);
        const double ri = fma(-alpha, q[i], r[i]);
        r[i] = ri;
        v = fma(ri, ri, v);
    }
    __shared__ double res;
    if (grid_reduce(v, partials, counter, &res)) {
        red[1] = res;
        red[3] = rho;
    }
}

// r = b - q ; *rr = sum r^2 (local)
__global__ void residual_kernel(const double* __restrict__ b, const double* __restrict__ q, double* __restrict__ r,
                                int64_t n, double* partials, unsigned* counter, double* rr) {
    double v = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double ri = b[i] - q[i];
        r[i] = ri;
        v = fma(ri, ri, v);
    }
    __shared__ double res;
    if (grid_reduce(v, partials, counter, &res)) *rr = res;
}

static inline unsigned vec_grid(int64_t n) {
    int64_t g = (n + VEC_THREADS - 1) / VEC_THREADS;
    if (g > VEC_MAX_BLOCKS) g = VEC_MAX_BLOCKS;
    if (g < 1) g = 1;
    return (unsigned)g;
}

struct PcgWs {
    int64_t n_pad, off_r, off_z, off_q, off_p, off_xg, off_u, off_mv, off_symv, off_state, off_hist, hist_len, total;
};
constexpr int64_t PCG_HIST_CAP = 1 << 20;  // device-side ||r||^2 history entries (longer runs keep the first 2^20)
static PcgWs pcg_layout(const mlffpc_ctx* c, int64_t k, bool matrix_free) {
    auto up = [](int64_t x) { return (x + 255) / 256 * 256; };
    PcgWs w;
    const int world = c->comm.world;
    const int64_t ppr = (c->M + world - 1) / world;
    w.n_pad = ppr * c->dim_i;
    const int64_t nl = c->n_local();
    int64_t o = 0;
    w.off_r = o; o = up(o + nl * 8);
    w.off_z = o; o = up(o + nl * 8);
    w.off_q = o; o = up(o + nl * 8);
    w.off_p = o; o = up(o + world * w.n_pad * 8);
    w.off_xg = o; o = up(o + (world > 1 ? world * w.n_pad * 8 : 0));
    w.off_u = o; o = up(o + (4 * k + 8) * 8);
    w.off_mv = o; o = up(o + (matrix_free ? matvec_free_ws_bytes(c) : 0));
    w.off_symv = o; o = up(o + ((!matrix_free && c->use_symv) ? symop_ws_bytes(c) : 0));
    w.off_state = o; o = up(o + 256);  // PcgState + red[8]
    w.hist_len = PCG_HIST_CAP;
    w.off_hist = o; o = up(o + w.hist_len * 8);
    w.total = o + 256;
    return w;
}

struct PcgOp {
    mlffpc_ctx* ctx;
    const double* K;
    int64_t ld_k;
    double lam;
    void* mv_ws;
    void* symv_ws;
    // q_local = A v,  A = -K + lam I;  v_full is the replicated n-vector
    int apply(const double* v_full, double* q_local, cudaStream_t s) const {
        if (K && ctx->use_symv)  // K is the symmetric tile storage of this rank (symop.cu)
            return symop_apply(ctx, K, v_full, q_local, -1.0, lam, symv_ws, nullptr, s);
        if (K)
            return launch_gemv_rows(K, ctx->n_local(), ctx->n, ld_k, v_full, q_local, -1.0, lam, ctx->row0(), s);
        return matvec_free(ctx, v_full, q_local, -1.0, lam, mv_ws, s);
    }
};

// events of the loop, destroyed on every exit path
struct EventRing {
    std::vector<cudaEvent_t> ev;
    ~EventRing() {
        for (auto e : ev) cudaEventDestroy(e);
    }
    int init(size_t count) {
        ev.reserve(count);
        for (size_t i = 0; i < count; ++i) {
            cudaEvent_t e;
            cudaError_t err = cudaEventCreate(&e);
            if (err != cudaSuccess) return cuda_fail(err, "cudaEventCreate", __FILE__, __LINE__);
            ev.push_back(e);
        }
        return MLFFPC_OK;
    }
};
struct ProfGuard {
    ProfWindow pw;
    ~ProfGuard() { pw.end(); }
};

}  // namespace mlffpc

using namespace mlffpc;

extern "C" {

int mlffpc_dot(mlffpc_ctx* ctx, const double* a, const double* b, int64_t n, double* out_host, void* stream) {
    MLFFPC_REQUIRE(ctx && a && b && out_host && n >= 0, "dot: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    unsigned* counter = (unsigned*)(ctx->scal + S_COUNTER);
    PeerXchg off;
    off.on = 0;
    dot_kernel<<<vec_grid(n), VEC_THREADS, 0, s>>>(a, b, n, ctx->partials, counter, ctx->scal + S_TMP, nullptr, off, nullptr);
    MLFFPC_LAUNCH_CHECK();
    MLFFPC_TRY(comm_allreduce_sum(ctx->comm, ctx->scal + S_TMP, 1, s));
    MLFFPC_CUDA(cudaMemcpyAsync(ctx->h_scal, ctx->scal + S_TMP, 8, cudaMemcpyDeviceToHost, s));
    MLFFPC_CUDA(cudaStreamSynchronize(s));
    *out_host = ctx->h_scal[0];
    return MLFFPC_OK;
}

int mlffpc_pcg_workspace_bytes(mlffpc_ctx* ctx, int64_t k, int matrix_free, int64_t* bytes) {
    MLFFPC_REQUIRE(ctx && bytes && k >= 0 && ctx->M > 0, "pcg_workspace_bytes: bad argument / geometry not set");
    *bytes = pcg_layout(ctx, k, matrix_free != 0).total;
    return MLFFPC_OK;
}

int mlffpc_pcg(mlffpc_ctx* ctx, const double* K_local, int64_t ld_k, double lam, const double* T,
               int64_t k, int64_t ld_t, double precon_sign, const double* Mk, const double* E, const double* b,
               double* x, double tol, int64_t maxiter, int64_t resume_iters, double* out_host, double* resid_hist_host,
               void* workspace, int64_t workspace_bytes, void* stream) {
    MLFFPC_REQUIRE(ctx && ctx->M > 0, "pcg: geometry not set");
    MLFFPC_REQUIRE(b && x && out_host && workspace, "pcg: NULL argument");
    MLFFPC_REQUIRE(lam > 0.0 && tol > 0.0 && maxiter >= 0, "pcg: bad lam/tol/maxiter");
    MLFFPC_REQUIRE(maxiter < ((int64_t)1 << 31) - 8, "pcg: maxiter too large");
    MLFFPC_REQUIRE(!K_local || ctx->use_symv || ld_k >= ctx->n, "pcg: ld_k < n");
    MLFFPC_REQUIRE(!ctx->use_symv || ctx->lay_world == ctx->comm.world, "pcg: symmetric tile layout does not match the communicator");
    MLFFPC_REQUIRE(!T || (k > 0 && ld_t >= ctx->n_local()), "pcg: bad preconditioner dimensions");
    MLFFPC_REQUIRE(!E || Mk, "pcg: the defect matrix E needs the orthonormal-form Mk");
    MLFFPC_REQUIRE(resume_iters >= 0 && resume_iters <= maxiter, "pcg: resume_iters out of range");
    const bool matrix_free = (K_local == nullptr);
    if (T && Mk) MLFFPC_TRY(ensure_reorth_scratch(ctx, k));
    const PcgWs w = pcg_layout(ctx, T ? k : 0, matrix_free);
    MLFFPC_REQUIRE(workspace_bytes >= w.total, "pcg: workspace too small (%lld < %lld)",
                   (long long)workspace_bytes, (long long)w.total);
    const int world = ctx->comm.world;
    const int64_t nl = ctx->n_local(), row0 = ctx->row0();
    if (world > 1) {
        const int64_t ppr = (ctx->M + world - 1) / world;
        MLFFPC_REQUIRE(ctx->pt0 == ctx->comm.rank * ppr, "pcg: multi-GPU needs the ceil(M/world) point partition");
    }
    cudaStream_t s = (cudaStream_t)stream;
    char* base = (char*)(((uintptr_t)workspace + 255) / 256 * 256);
    double* r = (double*)(base + w.off_r);
    double* z = (double*)(base + w.off_z);
    double* q = (double*)(base + w.off_q);
    // peer mode: the replicated search direction lives in this rank's peer buffer, where the other ranks' p-update
    // kernels store their slices directly
    const bool peer = peer_on(ctx);
    Peer* pr = peer ? ctx->peer : nullptr;
    double* p_full = peer ? (double*)((char*)pr->local + pr->view.lay.off_p) : (double*)(base + w.off_p);
    double* p = p_full + row0;
    PeerXchg off;
    off.on = 0;
    auto xchg = [&](int ch) -> PeerXchg {  // one round of a fused exchange: same epoch / slot for producer and consumer
        PeerXchg x;
        x.on = 0;
        if (peer) {
            x.pv = pr->view;
            x.on = 1;
            x.ch = ch;
            x.parity = (int)(pr->uses[ch]++ & 1);
            x.epoch = ++pr->epoch;
        }
        return x;
    };
    double* xg = (world > 1) ? (double*)(base + w.off_xg) : nullptr;
    double* u = (double*)(base + w.off_u);
    PcgState* state = (PcgState*)(base + w.off_state);
    double* red = (double*)(base + w.off_state + 64);
    double* hist = (double*)(base + w.off_hist);
    PcgOp A{ctx, K_local, ld_k, lam, (void*)(base + w.off_mv), (void*)(base + w.off_symv)};
    double* sc = ctx->scal;
    unsigned* counter = (unsigned*)(sc + S_COUNTER);
    const unsigned g = vec_grid(nl);

    auto host_scalar = [&](const double* dev, double* out) -> int {
        MLFFPC_CUDA(cudaMemcpyAsync(ctx->h_scal, dev, 8, cudaMemcpyDeviceToHost, s));
        MLFFPC_CUDA(cudaStreamSynchronize(s));
        *out = ctx->h_scal[0];
        return MLFFPC_OK;
    };
    // q = A x (x gathered when sharded), r = b - q, local ||r||^2 -> red[1]
    auto true_residual = [&]() -> int {
        const double* x_full = x;
        if (world > 1) {
            MLFFPC_CUDA(cudaMemsetAsync(xg + row0, 0, w.n_pad * 8, s));
            MLFFPC_CUDA(cudaMemcpyAsync(xg + row0, x, nl * 8, cudaMemcpyDeviceToDevice, s));
            MLFFPC_TRY(comm_allgather(ctx->comm, xg + row0, xg, w.n_pad * 8, s));
            x_full = xg;
        }
        MLFFPC_TRY(A.apply(x_full, q, s));
        residual_kernel<<<g, VEC_THREADS, 0, s>>>(b, q, r, nl, ctx->partials, counter, red + 1);
        MLFFPC_LAUNCH_CHECK();
        return MLFFPC_OK;
    };
    // global ||r|| from the local red[1] without disturbing it (the loop's allreduce expects the local value)
    auto global_resid = [&](double* out) -> int {
        MLFFPC_CUDA(cudaMemcpyAsync(sc + S_TMP, red + 1, 8, cudaMemcpyDeviceToDevice, s));
        MLFFPC_TRY(comm_allreduce_sum(ctx->comm, sc + S_TMP, 1, s));
        double rr = 0.0;
        MLFFPC_TRY(host_scalar(sc + S_TMP, &rr));
        *out = sqrt(rr);
        return MLFFPC_OK;
    };

    // resume_iters > 0: continue the run whose state (r, p, rho, x, history) the previous call left in this workspace
    // after stopping at its iteration cap -- same recurrence, no restart (checkpoint segments of Iterative.solve)
    const bool resume = resume_iters > 0;
    if (!resume) MLFFPC_CUDA(cudaMemsetAsync(base + w.off_state, 0, 256, s));
    else MLFFPC_CUDA(cudaMemsetAsync(state, 0, sizeof(PcgState), s));
    double bb = 0.0;
    dot_kernel<<<g, VEC_THREADS, 0, s>>>(b, b, nl, ctx->partials, counter, sc + S_TMP, nullptr, off, nullptr);
    MLFFPC_LAUNCH_CHECK();
    MLFFPC_TRY(comm_allreduce_sum(ctx->comm, sc + S_TMP, 1, s));
    MLFFPC_TRY(host_scalar(sc + S_TMP, &bb));
    const double bnrm2 = sqrt(bb);

    double resid = 0.0;
    if (!resume) MLFFPC_TRY(true_residual());
    MLFFPC_TRY(global_resid(&resid));
    out_host[3] = bnrm2;
    out_host[4] = out_host[5] = out_host[6] = out_host[7] = 0.0;
    if (resid_hist_host && !resume) resid_hist_host[0] = resid;
    if (!resume && resid <= tol) {  // legacy _get_atol probe
        out_host[0] = 0; out_host[1] = resid; out_host[2] = 0;
        return MLFFPC_OK;
    }
    const double atol = (bnrm2 == 0.0) ? tol : tol * bnrm2;
    const double atol2 = atol * atol;
    // (peer mode: the buffer was zeroed when it was created, and a faster rank may already be storing into it)
    if (world > 1 && !resume && !peer) MLFFPC_CUDA(cudaMemsetAsync(p_full, 0, (size_t)world * w.n_pad * 8, s));

    // batches of iterations; the state of batch i is read while batch i + 1 runs
    constexpr int BATCH = 4, SLOTS = 2;
    PcgState* h_state = (PcgState*)(ctx->h_scal + 16);  // pinned, SLOTS entries of 32 bytes
    EventRing ring;                                     // per slot: BATCH * 4 timing events + 1 "state copied" event
    MLFFPC_TRY(ring.init((size_t)SLOTS * (BATCH * 4 + 1)));
    auto ev_iter = [&](int slot, int i, int which) { return ring.ev[(size_t)slot * (BATCH * 4 + 1) + i * 4 + which]; };
    auto ev_done = [&](int slot) { return ring.ev[(size_t)slot * (BATCH * 4 + 1) + BATCH * 4]; };
    double op_ms = 0.0, pre_ms = 0.0;
    int64_t op_calls = 0;
    int slot_iters[SLOTS] = {0, 0};
    auto harvest = [&](int slot) -> int {  // wait for the slot's batch, add up its event times
        if (slot_iters[slot] == 0) return MLFFPC_OK;
        MLFFPC_CUDA(cudaEventSynchronize(ev_done(slot)));
        for (int i = 0; i < slot_iters[slot]; ++i) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ev_iter(slot, i, 0), ev_iter(slot, i, 1)) == cudaSuccess) pre_ms += ms;
            if (cudaEventElapsedTime(&ms, ev_iter(slot, i, 2), ev_iter(slot, i, 3)) == cudaSuccess) { op_ms += ms; ++op_calls; }
        }
        slot_iters[slot] = 0;
        return MLFFPC_OK;
    };

    // one iteration, launched without waiting for anything
    auto launch_iteration = [&](int64_t it, int slot, int i) -> int {
        MLFFPC_CUDA(cudaEventRecord(ev_iter(slot, i, 0), s));
        if (T) {
            MLFFPC_TRY(precon_apply(ctx, T, k, ld_t, lam, precon_sign, r, z, u, s, Mk, E));
        } else {
            MLFFPC_CUDA(cudaMemcpyAsync(z, r, nl * 8, cudaMemcpyDeviceToDevice, s));
        }
        MLFFPC_CUDA(cudaEventRecord(ev_iter(slot, i, 1), s));
        const PeerXchg xa = xchg(PEER_CH_SCAL_A), xp = xchg(PEER_CH_P), xb = xchg(PEER_CH_SCAL_B);
        dot_kernel<<<g, VEC_THREADS, 0, s>>>(r, z, nl, ctx->partials, counter, red + 0, state, xa, red + 1);
        MLFFPC_LAUNCH_CHECK();
        if (!peer) MLFFPC_TRY(comm_allreduce_sum(ctx->comm, red, 2, s));  // (rho, ||r||^2 of the previous update)
        update_p_kernel<<<g, VEC_THREADS, 0, s>>>(z, p, nl, red, red, state, atol2, it, 0, it - 1 < w.hist_len ? hist : nullptr,
                                                  xa, xp, row0);
        MLFFPC_LAUNCH_CHECK();
        if (peer) MLFFPC_TRY(peer_wait(ctx, PEER_CH_P, xp.epoch, &state->frozen, s));
        else if (world > 1) MLFFPC_TRY(comm_allgather(ctx->comm, p, p_full, w.n_pad * 8, s));
        MLFFPC_CUDA(cudaEventRecord(ev_iter(slot, i, 2), s));
        MLFFPC_TRY(A.apply(p_full, q, s));
        MLFFPC_CUDA(cudaEventRecord(ev_iter(slot, i, 3), s));
        dot_kernel<<<g, VEC_THREADS, 0, s>>>(p, q, nl, ctx->partials, counter, red + 2, state, xb, nullptr);
        MLFFPC_LAUNCH_CHECK();
        if (!peer) MLFFPC_TRY(comm_allreduce_sum(ctx->comm, red + 2, 1, s));
        update_xr_kernel<<<g, VEC_THREADS, 0, s>>>(x, r, p, q, nl, red, state, ctx->partials, counter, xb);
        MLFFPC_LAUNCH_CHECK();
        return MLFFPC_OK;
    };
    // the stopping test of the last launched iteration (its ||r||^2 is still local in red[1])
    // (on a copy: the next iteration's allreduce expects the local value in red[1])
    auto launch_check = [&](int64_t it_next) -> int {
        MLFFPC_CUDA(cudaMemcpyAsync(red + 4, red, 16, cudaMemcpyDeviceToDevice, s));
        MLFFPC_TRY(comm_allreduce_sum(ctx->comm, red + 4, 2, s));
        update_p_kernel<<<1, 32, 0, s>>>(z, p, nl, red, red + 4, state, atol2, it_next, 1, it_next - 1 < w.hist_len ? hist : nullptr,
                                         off, off, row0);
        MLFFPC_LAUNCH_CHECK();
        return MLFFPC_OK;
    };

    if (maxiter == resume_iters) {
        out_host[0] = (double)resume_iters; out_host[1] = resid; out_host[2] = 1;
        return MLFFPC_OK;
    }
    ProfGuard prof{prof_window("pcg")};
    int64_t launched = resume_iters;  // iterations launched so far
    int64_t it = 0;            // result: iterations the legacy loop would have run
    int info = (int)(maxiter > 0x7fffffff ? 0x7fffffff : maxiter);
    if (info == 0) info = 1;
    bool done = false;
    int slot = 0;
    double known_rr = resid * resid;  // latest ||r||^2 the host has seen (lags by up to two batches)
    while (!done) {
        // near the target (or in a profiling window) go one iteration at a time so that nothing runs past the stop
        const bool careful = known_rr <= 4.0 * atol2 || prof.pw.first >= 0;
        const int nb = (int)std::min<int64_t>(careful ? 1 : BATCH, maxiter - launched);
        MLFFPC_TRY(harvest(slot));  // the slot's events and host copy are about to be reused
        for (int i = 0; i < nb; ++i) {
            prof.pw.step(launched + 1);
            MLFFPC_TRY(launch_iteration(launched + 1, slot, i));
            ++launched;
        }
        if (launched == maxiter || careful) MLFFPC_TRY(launch_check(launched + 1));
        slot_iters[slot] = nb;
        MLFFPC_CUDA(cudaMemcpyAsync(&h_state[slot], state, sizeof(PcgState), cudaMemcpyDeviceToHost, s));
        MLFFPC_CUDA(cudaEventRecord(ev_done(slot), s));
        // look at the batch before this one (already finished or about to), or at this one when it was the last
        const bool last = (launched == maxiter) || careful;
        const int look = last ? slot : (slot ^ 1);
        if (last) MLFFPC_TRY(harvest(slot ^ 1));
        const bool have = slot_iters[look] > 0;
        if (have) MLFFPC_TRY(harvest(look));
        slot ^= 1;
        if (!have) continue;
        const PcgState hs = h_state[look];
        if (hs.last_iter > 0) known_rr = hs.last_rr;
        if (!hs.frozen) {
            if (launched == maxiter && last) { it = maxiter; resid = sqrt(known_rr); done = true; }
            continue;
        }
        // the test fired at iteration hs.conv_iter; everything launched after it left x, r, p untouched
        MLFFPC_TRY(harvest(0));
        MLFFPC_TRY(harvest(1));
        MLFFPC_CUDA(cudaStreamSynchronize(s));
        it = hs.conv_iter;
        if (hs.nan_flag) {
            set_error("pcg: residual became NaN at iteration %lld", (long long)it);
            return MLFFPC_ERR_LINALG;
        }
        resid = sqrt(hs.last_rr);
        if (it > 1) {  // legacy: recompute r = b - A x once on the first hit and test again
            MLFFPC_TRY(true_residual());
            MLFFPC_TRY(global_resid(&resid));
            if (it < w.hist_len) {
                const double rr_true = resid * resid;
                MLFFPC_CUDA(cudaMemcpyAsync(hist + it, &rr_true, 8, cudaMemcpyHostToDevice, s));
                MLFFPC_CUDA(cudaStreamSynchronize(s));
            }
        }
        if (resid <= atol) {
            info = 0;
            done = true;
        } else if (it >= maxiter) {
            done = true;
        } else {
            // not converged after all: continue from iteration it + 1 with the recomputed residual
            MLFFPC_CUDA(cudaMemsetAsync(state, 0, sizeof(PcgState), s));
            launched = it;
            known_rr = resid * resid;
        }
    }
    if (resid_hist_host) {
        const int64_t cnt = std::min<int64_t>(it, w.hist_len - 1);
        if (cnt > 0) {
            MLFFPC_CUDA(cudaMemcpyAsync(resid_hist_host + 1, hist + 1, (size_t)cnt * 8, cudaMemcpyDeviceToHost, s));
            MLFFPC_CUDA(cudaStreamSynchronize(s));
            for (int64_t i = 1; i <= cnt; ++i) resid_hist_host[i] = sqrt(resid_hist_host[i]);
        }
        if (it <= cnt) resid_hist_host[it] = resid;
    }
    out_host[0] = (double)it;
    out_host[1] = resid;
    out_host[2] = (double)info;
    out_host[4] = op_ms;
    out_host[5] = (double)op_calls;
    out_host[6] = pre_ms;
    return MLFFPC_OK;
}

}  // extern "C"
