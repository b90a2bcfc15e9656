// Peer-memory set-up (CUDA IPC) and the stand-alone peer collectives; the fused ones live in the kernels that
// produce / consume the data (pcg.cu, symop.cu, pchol.cu).  See peer.cuh for the protocol.
#include <string.h>

#include "common.cuh"
#include "peer.cuh"

namespace mlffpc {

static PeerLayout peer_layout(int world, int64_t k_max, int64_t n_pad) {
    auto up = [](int64_t x) { return (x + 255) / 256 * 256; };
    PeerLayout l;
    l.k_pad = (k_max + 7) / 8 * 8;
    l.n_full = (int64_t)world * n_pad;
    int64_t o = 0;
    l.off_flags = o;   o = up(o + (int64_t)PEER_NCH * PEER_MAX_RANKS * 8);
    l.off_counter = o; o = up(o + 64);
    l.off_scal = o;    o = up(o + 2 * (int64_t)PEER_MAX_RANKS * 8 * 8);
    l.off_kvec = o;    o = up(o + 2 * (int64_t)PEER_MAX_RANKS * l.k_pad * 8);
    l.off_p = o;       o = up(o + l.n_full * 8);
    l.off_yp = o;      o = up(o + l.n_full * 8);
    l.off_msg = o;     o = up(o + 2 * (int64_t)PEER_MAX_RANKS * PEER_MSG_DOUBLES * 8);
    l.total = o;
    return l;
}

// in place: w[0:k) <- sum over ranks of w, combined in rank order on every rank (one CTA)
__global__ void __launch_bounds__(1024)
peer_kvec_allreduce_kernel(const PeerView pv, double* __restrict__ w, int64_t k, int parity, uint64_t epoch) {
    for (int r = 0; r < pv.world; ++r) {
        double* dst = pv.kvec(r, parity, pv.rank);
        for (int64_t i = threadIdx.x; i < k; i += blockDim.x) dst[i] = w[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        peer_signal_all(pv, PEER_CH_KVEC, epoch);
        peer_wait_all(pv, PEER_CH_KVEC, epoch);
    }
    __syncthreads();
    for (int64_t i = threadIdx.x; i < k; i += blockDim.x) {
        double t = 0.0;
        for (int r = 0; r < pv.world; ++r) t += peer_ld(pv.kvec(pv.rank, parity, r) + i);
        w[i] = t;
    }
}

// block the stream until every rank has raised channel ch to epoch (skipped while the PCG state is frozen: the
// producers skip their stores then, on every rank alike)
__global__ void peer_wait_kernel(const PeerView pv, int ch, uint64_t epoch, const int* frozen) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        if (frozen && *frozen) return;
        peer_wait_all(pv, ch, epoch);
    }
}

bool peer_on(const mlffpc_ctx* ctx) { return ctx->peer && ctx->peer->enabled && ctx->comm.world > 1; }

int peer_allreduce_kvec(mlffpc_ctx* ctx, double* w, int64_t k, cudaStream_t s) {
    Peer* p = ctx->peer;
    const int parity = (int)(p->uses[PEER_CH_KVEC]++ & 1);
    peer_kvec_allreduce_kernel<<<1, 1024, 0, s>>>(p->view, w, k, parity, ++p->epoch);
    MLFFPC_LAUNCH_CHECK();
    return MLFFPC_OK;
}

int peer_wait(mlffpc_ctx* ctx, int ch, uint64_t epoch, const int* frozen, cudaStream_t s) {
    peer_wait_kernel<<<1, 32, 0, s>>>(ctx->peer->view, ch, epoch, frozen);
    MLFFPC_LAUNCH_CHECK();
    return MLFFPC_OK;
}

int64_t peer_kmax(const mlffpc_ctx* ctx) { return ctx->peer ? ctx->peer->k_max : 0; }
double* peer_yp_local(const mlffpc_ctx* ctx) { return (double*)((char*)ctx->peer->local + ctx->peer->view.lay.off_yp); }

// Fused reduce-scatter + finish of the symmetric tile operator: every rank's full-length partial product sits in its
// peer buffer (yp); this rank raises PEER_CH_YP, waits for the others, pulls their partials of ITS rows over NVLink
// and writes y = alpha * sum + shift * x.
__global__ void symop_finish_peer_kernel(const PeerView pv, uint64_t epoch, int64_t g_off, const double* __restrict__ x_local,
                                         double* __restrict__ y, int64_t n, double alpha, double shift) {
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) peer_signal_all(pv, PEER_CH_YP, epoch);  // the tile kernels before this one have completed
        peer_wait_all(pv, PEER_CH_YP, epoch);
    }
    __syncthreads();
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    double q = 0.0;
    for (int r = 0; r < pv.world; ++r) q += peer_ld(pv.yp(r) + g_off + t);
    y[t] = fma(shift, x_local[t], alpha * q);
}

int symop_finish_peer(mlffpc_ctx* ctx, const double* x_local, double* y_local, int64_t nl, int64_t g_off, double alpha,
                      double shift, cudaStream_t s) {
    Peer* p = ctx->peer;
    ++p->uses[PEER_CH_YP];
    symop_finish_peer_kernel<<<(unsigned)((nl + 255) / 256), 256, 0, s>>>(p->view, ++p->epoch, g_off, x_local, y_local, nl,
                                                                         alpha, shift);
    MLFFPC_LAUNCH_CHECK();
    return MLFFPC_OK;
}

void peer_destroy(mlffpc_ctx* ctx) {
    Peer* p = ctx->peer;
    if (!p) return;
    for (int r = 0; r < PEER_MAX_RANKS; ++r)
        if (p->opened[r]) cudaIpcCloseMemHandle(p->opened[r]);
    if (p->local) cudaFree(p->local);
    delete p;
    ctx->peer = nullptr;
}

}  // namespace mlffpc

using namespace mlffpc;

extern "C" {

int mlffpc_peer_export(mlffpc_ctx* ctx, int64_t k_max, void* handle64_out) {
    MLFFPC_REQUIRE(ctx && handle64_out && ctx->M > 0, "peer_export: geometry not set or NULL argument");
    MLFFPC_REQUIRE(ctx->comm.world > 1 && ctx->comm.world <= PEER_MAX_RANKS, "peer_export: needs 2..%d ranks", PEER_MAX_RANKS);
    MLFFPC_REQUIRE(k_max > 0, "peer_export: bad k_max");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    peer_destroy(ctx);
    const int world = ctx->comm.world;
    const int64_t n_pad = ((ctx->M + world - 1) / world) * ctx->dim_i;
    Peer* p = new Peer();
    p->view.rank = ctx->comm.rank;
    p->view.world = world;
    p->view.lay = peer_layout(world, k_max, n_pad);
    p->k_max = k_max;
    for (int r = 0; r < PEER_MAX_RANKS; ++r) p->view.base[r] = nullptr;
    cudaError_t e = cudaMalloc(&p->local, (size_t)p->view.lay.total);
    if (e == cudaSuccess) e = cudaMemset(p->local, 0, (size_t)p->view.lay.total);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p->local);
    if (e != cudaSuccess) {
        if (p->local) cudaFree(p->local);
        delete p;
        return cuda_fail(e, "peer buffer / cudaIpcGetMemHandle", __FILE__, __LINE__);
    }
    memcpy(handle64_out, &h, 64);
    ctx->peer = p;
    return MLFFPC_OK;
}

int mlffpc_peer_import(mlffpc_ctx* ctx, const void* handles, int count) {
    MLFFPC_REQUIRE(ctx && handles && ctx->peer, "peer_import: call peer_export first");
    Peer* p = ctx->peer;
    MLFFPC_REQUIRE(count == p->view.world, "peer_import: %d handles for %d ranks", count, p->view.world);
    for (int r = 0; r < count; ++r) {
        if (r == p->view.rank) {
            p->view.base[r] = (char*)p->local;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + 64 * r, 64);
        void* ptr = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return cuda_fail(e, "cudaIpcOpenMemHandle", __FILE__, __LINE__);
        p->opened[r] = ptr;
        p->view.base[r] = (char*)ptr;
    }
    p->enabled = true;
    return MLFFPC_OK;
}

int mlffpc_peer_disable(mlffpc_ctx* ctx) {
    MLFFPC_REQUIRE(ctx, "peer_disable: NULL context");
    peer_destroy(ctx);
    return MLFFPC_OK;
}

}  // extern "C"
