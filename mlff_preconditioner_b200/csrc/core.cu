// Context lifecycle, error plumbing, NCCL binding (dlopen).
#include <cuda_profiler_api.h>
#include <dlfcn.h>
#include <stdlib.h>
#include <stdarg.h>
#include <string.h>
#include <time.h>

#include "common.cuh"

namespace mlffpc {

static thread_local char g_err[1024] = "";
long long g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return MLFFPC_ERR_CUDA;
}

ProfWindow prof_window(const char* phase) {
    ProfWindow w;
    const char* env = getenv("MLFFPC_PROFILE");
    if (!env) return w;
    const size_t pl = strlen(phase);
    for (const char* p = env; p && *p;) {
        if (strncmp(p, phase, pl) == 0 && p[pl] == ':') {
            long long a = 0, b = 1;
            if (sscanf(p + pl + 1, "%lld:%lld", &a, &b) >= 1) { w.first = a; w.count = b; }
            break;
        }
        p = strchr(p, ',');
        if (p) ++p;
    }
    return w;
}
static double wall_ms() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
PhaseTimer::PhaseTimer(cudaStream_t stream) : s(stream), on(false), t0(0.0) {
    const char* e = getenv("MLFFPC_TIMING");
    on = e && e[0] == '1';
    if (on) { cudaStreamSynchronize(s); t0 = wall_ms(); }
}
void PhaseTimer::lap(const char* label) {
    if (!on) return;
    cudaStreamSynchronize(s);
    const double t1 = wall_ms();
    fprintf(stderr, "[mlffpc timing] %-28s %9.3f ms\n", label, t1 - t0);
    t0 = t1;
}
void ProfWindow::step(long long i) {
    if (first < 0) return;
    if (!active && i == first) { cudaDeviceSynchronize(); cudaProfilerStart(); active = true; }
    else if (active && i >= first + count) end();
}
void ProfWindow::end() {
    if (active) { cudaDeviceSynchronize(); cudaProfilerStop(); active = false; }
}

// ------------------------------------------------------------------ NCCL via dlopen
typedef int (*nccl_get_unique_id_t)(void*);
struct NcclId {
    char internal[128];
};
typedef int (*nccl_comm_init_rank_fn)(void**, int, NcclId, int);
typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_allgather_fn)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*nccl_bcast_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_reduce_scatter_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_destroy_fn)(void*);
typedef const char* (*nccl_errstr_fn)(int);

struct NcclApi {
    void* lib = nullptr;
    nccl_get_unique_id_t get_unique_id = nullptr;
    nccl_comm_init_rank_fn comm_init_rank = nullptr;
    nccl_allreduce_fn allreduce = nullptr;
    nccl_allgather_fn allgather = nullptr;
    nccl_bcast_fn bcast = nullptr;
    nccl_reduce_scatter_fn reduce_scatter = nullptr;
    nccl_destroy_fn destroy = nullptr;
    nccl_errstr_fn errstr = nullptr;
};
static NcclApi g_nccl;
static void* g_cached_comm = nullptr;  // process-lifetime communicator (see mlffpc_comm_init)
static int g_cached_rank = -1, g_cached_world = -1, g_cached_device = -1;

static int nccl_load(const char* path) {
    if (g_nccl.lib) return MLFFPC_OK;
    void* lib = dlopen(path && path[0] ? path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) {
        set_error("dlopen(%s) failed: %s", path ? path : "libnccl.so.2", dlerror());
        return MLFFPC_ERR_COMM;
    }
    g_nccl.get_unique_id = (nccl_get_unique_id_t)dlsym(lib, "ncclGetUniqueId");
    g_nccl.comm_init_rank = (nccl_comm_init_rank_fn)dlsym(lib, "ncclCommInitRank");
    g_nccl.allreduce = (nccl_allreduce_fn)dlsym(lib, "ncclAllReduce");
    g_nccl.allgather = (nccl_allgather_fn)dlsym(lib, "ncclAllGather");
    g_nccl.bcast = (nccl_bcast_fn)dlsym(lib, "ncclBroadcast");
    g_nccl.reduce_scatter = (nccl_reduce_scatter_fn)dlsym(lib, "ncclReduceScatter");
    g_nccl.destroy = (nccl_destroy_fn)dlsym(lib, "ncclCommDestroy");
    g_nccl.errstr = (nccl_errstr_fn)dlsym(lib, "ncclGetErrorString");
    if (!g_nccl.get_unique_id || !g_nccl.comm_init_rank || !g_nccl.allreduce || !g_nccl.allgather ||
        !g_nccl.bcast || !g_nccl.destroy || !g_nccl.reduce_scatter) {
        set_error("libnccl is missing a required symbol");
        return MLFFPC_ERR_COMM;
    }
    g_nccl.lib = lib;
    return MLFFPC_OK;
}

static int nccl_check(int r, const char* what) {
    if (r == 0) return MLFFPC_OK;
    set_error("NCCL error %d in %s: %s", r, what, g_nccl.errstr ? g_nccl.errstr(r) : "?");
    return MLFFPC_ERR_COMM;
}

// ncclDataType_t: ncclInt8 = 0 (ncclChar), ncclFloat64 = 8;  ncclRedOp_t: ncclSum = 0
int comm_allreduce_sum(Comm& c, double* buf, size_t count, cudaStream_t s) {
    if (c.world <= 1 || count == 0) return MLFFPC_OK;
    return nccl_check(g_nccl.allreduce(buf, buf, count, 8, 0, c.comm, s), "ncclAllReduce");
}
int comm_allgather(Comm& c, const void* send, void* recv, size_t bytes_per_rank, cudaStream_t s) {
    if (c.world <= 1) {
        if (send != recv) {
            cudaError_t e = cudaMemcpyAsync(recv, send, bytes_per_rank, cudaMemcpyDeviceToDevice, s);
            if (e != cudaSuccess) return cuda_fail(e, "allgather self copy", __FILE__, __LINE__);
        }
        return MLFFPC_OK;
    }
    return nccl_check(g_nccl.allgather(send, recv, bytes_per_rank, 0, c.comm, s), "ncclAllGather");
}
int comm_reduce_scatter_sum(Comm& c, const double* send, double* recv, size_t count_per_rank, cudaStream_t s) {
    if (c.world <= 1) {
        cudaError_t e = cudaMemcpyAsync(recv, send, count_per_rank * 8, cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) return cuda_fail(e, "reduce_scatter self copy", __FILE__, __LINE__);
        return MLFFPC_OK;
    }
    return nccl_check(g_nccl.reduce_scatter(send, recv, count_per_rank, 8, 0, c.comm, s), "ncclReduceScatter");
}
int comm_broadcast(Comm& c, void* buf, size_t bytes, int root, cudaStream_t s) {
    if (c.world <= 1 || bytes == 0) return MLFFPC_OK;
    return nccl_check(g_nccl.bcast(buf, buf, bytes, 0, root, c.comm, s), "ncclBroadcast");
}

}  // namespace mlffpc

using namespace mlffpc;

extern "C" {

int mlffpc_version(void) { return 200; }

int64_t mlffpc_launch_count(void) { return (int64_t)g_launches; }

const char* mlffpc_last_error(void) { return g_err; }

int mlffpc_create(mlffpc_ctx** out, int device) {
    MLFFPC_REQUIRE(out != nullptr, "mlffpc_create: out is NULL");
    int count = 0;
    MLFFPC_CUDA(cudaGetDeviceCount(&count));
    MLFFPC_REQUIRE(device >= 0 && device < count, "mlffpc_create: device %d out of range (%d devices)", device, count);
    MLFFPC_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    MLFFPC_CUDA(cudaGetDeviceProperties(&prop, device));
    MLFFPC_REQUIRE(prop.major >= 10, "mlffpc_create: this library is built for sm_100a only (device is sm_%d%d)",
                   prop.major, prop.minor);
    // Scratch of the Gram kernels comes from the stream-ordered allocator; keep what it has mapped instead of handing
    // it back at every synchronisation (the default release threshold is 0: ~0.4 s of re-mapping per solve at k = 4839)
    {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    mlffpc_ctx* c = new mlffpc_ctx();
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    MLFFPC_CUDA(cudaMalloc(&c->scal, MLFFPC_NUM_SCAL * sizeof(double)));
    MLFFPC_CUDA(cudaMemset(c->scal, 0, MLFFPC_NUM_SCAL * sizeof(double)));
    MLFFPC_CUDA(cudaMallocHost(&c->h_scal, MLFFPC_NUM_SCAL * sizeof(double)));
    MLFFPC_CUDA(cudaMalloc(&c->partials, MLFFPC_MAX_PARTIALS * 4 * sizeof(double)));
    *out = c;
    return MLFFPC_OK;
}

int mlffpc_destroy(mlffpc_ctx* ctx) {
    if (!ctx) return MLFFPC_OK;
    if (ctx->comm.comm && ctx->comm.comm != g_cached_comm && g_nccl.destroy) g_nccl.destroy(ctx->comm.comm);
    if (ctx->scal) cudaFree(ctx->scal);
    if (ctx->h_scal) cudaFreeHost(ctx->h_scal);
    if (ctx->partials) cudaFree(ctx->partials);
    if (ctx->reorth_scratch) cudaFree(ctx->reorth_scratch);
    if (ctx->rows_ws) cudaFree(ctx->rows_ws);
    peer_destroy(ctx);
    delete ctx;
    return MLFFPC_OK;
}

int mlffpc_comm_unique_id(const char* libnccl_path, void* id128_out) {
    MLFFPC_REQUIRE(id128_out != nullptr, "comm_unique_id: out is NULL");
    MLFFPC_TRY(nccl_load(libnccl_path));
    return nccl_check(g_nccl.get_unique_id(id128_out), "ncclGetUniqueId");
}

int mlffpc_comm_init(mlffpc_ctx* ctx, const char* libnccl_path, const void* id128, int rank, int world) {
    MLFFPC_REQUIRE(ctx && id128, "comm_init: NULL argument");
    MLFFPC_REQUIRE(world >= 1 && rank >= 0 && rank < world, "comm_init: bad rank %d / world %d", rank, world);
    ctx->comm.rank = rank;
    ctx->comm.world = world;
    ctx->lay_rank = rank;
    ctx->lay_world = world;
    if (world == 1) return MLFFPC_OK;
    MLFFPC_TRY(nccl_load(libnccl_path));
    MLFFPC_CUDA(cudaSetDevice(ctx->device));
    // One communicator per process and (rank, world, device): creating one costs seconds, and a driver that
    // builds an engine per solve (Iterative.solve) would pay it every time.  Every rank takes the same branch.
    if (g_cached_comm && g_cached_rank == rank && g_cached_world == world && g_cached_device == ctx->device) {
        ctx->comm.comm = g_cached_comm;
        return MLFFPC_OK;
    }
    NcclId id;
    memcpy(id.internal, id128, 128);
    MLFFPC_TRY(nccl_check(g_nccl.comm_init_rank(&ctx->comm.comm, world, id, rank), "ncclCommInitRank"));
    if (!g_cached_comm) {
        g_cached_comm = ctx->comm.comm;
        g_cached_rank = rank; g_cached_world = world; g_cached_device = ctx->device;
    }
    return MLFFPC_OK;
}

int mlffpc_set_option(mlffpc_ctx* ctx, const char* name, int64_t value) {
    MLFFPC_REQUIRE(ctx && name, "set_option: NULL argument");
    const std::string nm(name);
    if (nm == "symmetric_gemv") { ctx->use_symv = value != 0; return MLFFPC_OK; }
    if (nm == "tgemv_msplit") { ctx->tgemv_msplit = (value == 1 || value == 4 || value == 8) ? (int)value : 0; return MLFFPC_OK; }
    if (nm == "precon_reorth") {
        ctx->precon_reorth = value != 0;
        return MLFFPC_OK;
    }
    if (nm == "assemble_legacy") { ctx->assemble_legacy = value != 0; return MLFFPC_OK; }
    if (nm == "pairs_kernel") { ctx->pairs_kernel = (value >= 1 && value <= 3) ? (int)value : 0; return MLFFPC_OK; }
    if (nm == "peer_pivots") { ctx->peer_pivots = value != 0; return MLFFPC_OK; }
    if (nm == "peer_kvec") { ctx->peer_kvec = value != 0; return MLFFPC_OK; }
    if (nm == "tma_rows") { ctx->tma_rows = value != 0 ? 1 : 0; return MLFFPC_OK; }
    if (nm == "defect_mode") { ctx->defect_mode = value == 2 ? 2 : 1; return MLFFPC_OK; }
    if (nm == "symop_multi") { ctx->symop_multi = value != 0 ? 1 : 0; return MLFFPC_OK; }
    if (nm == "gram_fold") { ctx->gram_fold = (value == 1 || value == 2 || value == 4) ? (int)value : 0; return MLFFPC_OK; }
    if (nm == "gram_mode") { ctx->gram_mode = value != 0 ? 1 : 0; return MLFFPC_OK; }
    if (nm == "syrk_chunk") { ctx->syrk_chunk = value > 0 ? value : 0; return MLFFPC_OK; }
    if (nm == "precon_accuracy") { ctx->precon_accuracy = (int)value; return MLFFPC_OK; }
    if (nm == "pchol_graph") { ctx->pchol_graph = value != 0; return MLFFPC_OK; }
    if (nm == "pchol_lookahead") { ctx->pchol_lookahead = value != 0; return MLFFPC_OK; }
    if (nm == "layout_world") {
        MLFFPC_REQUIRE(value >= 1 && value <= 1024, "set_option: layout_world out of range");
        ctx->lay_world = (int)value;
        if (ctx->lay_rank >= ctx->lay_world) ctx->lay_rank = 0;
        return MLFFPC_OK;
    }
    if (nm == "layout_rank") {
        MLFFPC_REQUIRE(value >= 0 && value < ctx->lay_world, "set_option: layout_rank out of range");
        ctx->lay_rank = (int)value;
        return MLFFPC_OK;
    }
    set_error("set_option: unknown option '%s'", name);
    return MLFFPC_ERR_INVALID;
}

int mlffpc_allreduce_sum(mlffpc_ctx* ctx, double* buf, int64_t count, void* stream) {
    MLFFPC_REQUIRE(ctx && (buf || count == 0) && count >= 0, "allreduce_sum: bad argument");
    return comm_allreduce_sum(ctx->comm, buf, (size_t)count, (cudaStream_t)stream);
}

int mlffpc_allgather(mlffpc_ctx* ctx, const void* send, void* recv, int64_t bytes_per_rank, void* stream) {
    MLFFPC_REQUIRE(ctx && send && recv && bytes_per_rank >= 0, "allgather: bad argument");
    return comm_allgather(ctx->comm, send, recv, (size_t)bytes_per_rank, (cudaStream_t)stream);
}

}  // extern "C"
