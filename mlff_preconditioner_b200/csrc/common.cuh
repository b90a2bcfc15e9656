// Shared declarations for libmlffpc (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/mlffpc.h"

namespace mlffpc {

// ---- error plumbing: every extern "C" entry returns a status, message via mlffpc_last_error() ----
void set_error(const char* fmt, ...);
extern long long g_launches;  // kernels launched by this library (bench.py's gpu_launches)
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

// Profiler windows (ncu --profile-from-start off): the environment variable MLFFPC_PROFILE holds
// "<phase>:<first>:<count>[,<phase>:...]" with phase in {pcg, pchol, woodbury, syrk, potrf, trsm, assemble}; the library brackets
// iterations/steps [first, first+count) of that phase with cudaProfilerStart/Stop.  Unset = no effect.
struct ProfWindow {
    long long first = -1, count = 0;
    bool active = false;
    void step(long long i);  // call at the start of iteration/step i
    void end();              // call when the phase ends
};
ProfWindow prof_window(const char* phase);

// MLFFPC_TIMING=1: the factorisation entry points synchronise after every sub-step and print its wall time to stderr
// (development aid; unset = no synchronisation, no output)
struct PhaseTimer {
    cudaStream_t s;
    bool on;
    double t0;
    explicit PhaseTimer(cudaStream_t stream);
    void lap(const char* label);
};

#define MLFFPC_CUDA(call)                                                          \
    do {                                                                           \
        cudaError_t _e = (call);                                                   \
        if (_e != cudaSuccess) return mlffpc::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define MLFFPC_LAUNCH_CHECK()                                                      \
    do {                                                                           \
        ++mlffpc::g_launches;                                                      \
        cudaError_t _e = cudaGetLastError();                                       \
        if (_e != cudaSuccess) return mlffpc::cuda_fail(_e, "kernel launch", __FILE__, __LINE__); \
    } while (0)

#define MLFFPC_REQUIRE(cond, ...)                                                  \
    do {                                                                           \
        if (!(cond)) {                                                             \
            mlffpc::set_error(__VA_ARGS__);                                        \
            return MLFFPC_ERR_INVALID;                                             \
        }                                                                          \
    } while (0)

#define MLFFPC_TRY(call)                                                           \
    do {                                                                           \
        int _s = (call);                                                           \
        if (_s != MLFFPC_OK) return _s;                                            \
    } while (0)

// ---- NCCL, bound at run time with dlopen (comm.cu) ----
struct Comm {
    void* lib = nullptr;
    void* comm = nullptr;  // ncclComm_t
    int rank = 0;
    int world = 1;
};
int comm_allreduce_sum(Comm& c, double* buf, size_t count, cudaStream_t s);
int comm_allgather(Comm& c, const void* send, void* recv, size_t bytes_per_rank, cudaStream_t s);
int comm_broadcast(Comm& c, void* buf, size_t bytes, int root, cudaStream_t s);
int comm_reduce_scatter_sum(Comm& c, const double* send, double* recv, size_t count_per_rank, cudaStream_t s);

struct Peer;  // NVLink peer-memory collectives (peer.cuh)

}  // namespace mlffpc

// The opaque context.  Holds no large allocations: geometry tables live in caller-owned workspace.
struct mlffpc_ctx {
    int device = 0;
    int num_sms = 148;
    // geometry (device pointers into caller buffers / workspace)
    int64_t M = 0;
    int N = 0, S = 0, D = 0, dim_i = 0;
    int64_t n = 0;
    double sig = 0.0;
    const double* R_desc = nullptr;    // [M, D]
    const double* R_d_desc = nullptr;  // [M, D, 3]
    const int32_t* desc_perms = nullptr;  // [S, D]  pi_p(d)
    const int32_t* atom_perms = nullptr;  // [S, N]  P_p
    int32_t* atom_perms_inv = nullptr;    // [S, N]  (workspace)
    int32_t* pair_a = nullptr;            // [D] larger atom index   (workspace)
    int32_t* pair_b = nullptr;            // [D] smaller atom index  (workspace)
    double* Xp = nullptr;                 // [M*S, D] permuted descriptors (workspace; == R_desc if S == 1)
    // row-block shard: local rows are points [pt0, pt1)
    int64_t pt0 = 0, pt1 = 0;
    int64_t n_local() const { return (pt1 - pt0) * (int64_t)dim_i; }
    int64_t row0() const { return pt0 * (int64_t)dim_i; }
    mlffpc::Comm comm;
    mlffpc::Peer* peer = nullptr;  // mapped peer buffers (mlffpc_peer_export / _import); NULL: NCCL only
    // partition used by the symmetric tile operator; follows the communicator unless overridden by the
    // options "layout_rank"/"layout_world" (rank emulation on one GPU, tests only)
    int lay_rank = 0, lay_world = 1;
    // option "precon_reorth" (experimental, off): project the complement twice in the orthonormal-form apply; the
    // scratch (n_local + k doubles) is allocated by the library when the option is switched on
    bool precon_reorth = false;
    double* reorth_scratch = nullptr;
    int64_t reorth_scratch_len = 0;
    bool assemble_legacy = false;  // option "assemble_legacy": one CTA per 3N x 3N block (the first-generation kernel)
    int gram_mode = 1;             // option "gram_mode": 1 (default) = Gram matrices with (hi, lo) accumulation of the k-tile
                                   // products (gramdd.cu); 0 = one running fp64 sum per entry (the round-1 kernel)
    int trsm_order = -1;           // set by a factorisation for its TRSM calls: 1 right-looking, 0 left-looking, -1 by column count
    int symop_multi = 1;           // option "symop_multi": 1 = all tiles of a rank in one persistent launch, 0 = tile by tile
    int gram_fold = 0;             // option "gram_fold": k-tiles (16 columns) per (hi, lo) fold: 1 (default; 0 = default), 2 or 4
    int defect_mode = 1;           // option "defect_mode": E = Q Q^T - I from 1 = the DMMA kernel with (hi, lo) k-tile folding,
                                   // 2 = exact products and sums on the FP64 vector pipe (~7x slower; reference for tests)
    int64_t syrk_chunk = 0;        // option "syrk_chunk": > 0 = Gram matrices by column chunks with Kahan-summed partials (diagnostics)
    int tgemv_msplit = 0;          // option "tgemv_msplit": force the row split of T^T u (1, 4, 8; 0 = auto)
    int precon_accuracy = 0;       // option "precon_accuracy": 1 = Kahan-compensated T r and T^T u (diagnostics)
    bool pchol_graph = true;       // option "pchol_graph": replay a chunk of pivot steps as one CUDA graph when no NCCL call sits inside
    bool pchol_lookahead = true;   // option "pchol_lookahead": candidate-panel (blocked) pivoted Cholesky
    long long last_pchol_refills = 0;  // panel rebuilds of the last mlffpc_pchol_build (diagnostics)
    bool tma_attr_symv = false, tma_attr_rows = false, tma_attr_multi = false;  // cudaFuncSetAttribute done for this context's device
    double* rows_ws = nullptr;     // scratch of the TMA row-strip GEMV (precon.cu), owned by the context
    int64_t rows_ws_len = 0;
    int pairs_kernel = 0;          // option "pairs_kernel": 0 = by descriptor length (default), 1 = 64 x 64 tiles, 2 = 128 x 64 + cp.async ring, 3 = 128 x 32, two CTAs per SM
    int pairs2_attr = 0;           // bit per instantiation of mv_pairs2_kernel whose shared-memory attribute is set
    bool peer_pivots = true;       // option "peer_pivots": pivot-step message over peer memory when mapped
    bool peer_kvec = true;         // option "peer_kvec": k-vector allreduce of the apply over peer memory when mapped
    int tma_rows = 1;              // option "tma_rows": 1 = T r of the preconditioner apply on the TMA row-strip kernel
    bool use_symv = false;  // option "symmetric_gemv": the assembled operator is the symmetric tile storage (symop.cu)
    // small persistent device scratch owned by the ctx (scalars / partial reductions, a few KB)
    double* scal = nullptr;    // device scalars
    double* h_scal = nullptr;  // pinned host mirror
    double* partials = nullptr;  // [MLFFPC_MAX_PARTIALS * 4]
};

#define MLFFPC_MAX_PARTIALS 65536
#define MLFFPC_NUM_SCAL 64

namespace mlffpc {

// ---- device helpers ----
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum, result valid in every thread; sm must hold >= 33 doubles
__device__ __forceinline__ double block_sum(double v, double* sm) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sm[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = (lane < nw) ? sm[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) sm[32] = t;
    }
    __syncthreads();
    return sm[32];
}

// index of pair (x, y), x != y, in np.tril_indices(N, -1) order: a(a-1)/2 + b with a > b
__device__ __forceinline__ int pair_index(int x, int y) {
    const int a = x > y ? x : y, b = x > y ? y : x;
    return a * (a - 1) / 2 + b;
}

// HBM-bound matrix-vector kernels (gemv.cu)
int launch_gemv_rows(const double* K, int64_t n_rows, int64_t n_cols, int64_t ld, const double* x,
                     double* y, double alpha, double shift, int64_t x_off, cudaStream_t s,
                     bool compensated = false);
int launch_tgemv_cols(const double* T, int64_t k, int64_t n_cols, int64_t ld, const double* w,
                      double* out, int post, const double* r, double sign_over_lam, int num_sms,
                      cudaStream_t s, bool compensated = false, int force_msplit = 0, const double* w2 = nullptr,
                      double sign = 1.0);
// symmetric operator (symop.cu, symtma.cu)
int64_t symv_ws_bytes(int64_t n);
int launch_symv(mlffpc_ctx* ctx, const double* K, int64_t n, int64_t ld, const double* x, double* y, double alpha,
                double shift, void* workspace, cudaStream_t s);
int64_t symop_storage_elems(const mlffpc_ctx* ctx);
int64_t symop_ws_bytes(const mlffpc_ctx* ctx);
int symop_apply(mlffpc_ctx* ctx, const double* Ksym, const double* x_full, double* y_local, double alpha,
                double shift, void* workspace, double* partial_out, cudaStream_t s);
int64_t symv_tma_ws_doubles(int64_t nr, int64_t nc);
int64_t rows_tma_ws_doubles(int64_t nr, int64_t nc, int num_sms);
bool rows_tma_usable(const double* A, int64_t ld, int64_t nr, int64_t nc);
int rows_gemv_tma(mlffpc_ctx* ctx, const double* A, int64_t nr, int64_t nc, int64_t ld, const double* x, double* y,
                  double alpha, double* wsd, cudaStream_t s);
bool symv_tma_usable(const double* K, int64_t ld, int64_t nr);
int symv_tile_tma(mlffpc_ctx* ctx, const double* K, int64_t ld, int64_t nr, int64_t nc, int diag, int packed,
                  const double* xr, const double* xc, double* wsd, double* out_c, double* out_r,
                  const double* x_shift, double alpha, double shift, cudaStream_t s);
// all tiles of a rank in one persistent launch (symtma.cu); tile 0 = the diagonal tile
struct SymTileIn {
    const double* K;
    int64_t ld, nr, nc, row_off;
    int diag, packed;
    const double* xr;
    const double* xc;
    double* out_c;
};
int symv_tiles_tma(mlffpc_ctx* ctx, int ntiles, const SymTileIn* in, double* wsd, cudaStream_t s);
// explicit rectangle K[points i_pt0:i_pt1, points j_pt0:j_pt1] -> out; packed = 0: row-major with leading
// dimension ld; packed = 1 (square diagonal tiles): the band layout of symlayout.cuh, entries right of a
// band's pitch are not stored
int assemble_tile(mlffpc_ctx* ctx, int64_t i_pt0, int64_t i_pt1, int64_t j_pt0, int64_t j_pt1, double* out,
                  int64_t ld, int packed, cudaStream_t s);
// one column of scale*K on the local rows, column index read from device memory (geometry.cu)
int launch_columns_device_col(mlffpc_ctx* ctx, const int64_t* col_dev, double* out, double scale,
                              cudaStream_t s);
// matrix-free operator (matvec.cu)
int64_t matvec_free_ws_bytes(const mlffpc_ctx* ctx);
int matvec_free(mlffpc_ctx* ctx, const double* v, double* y_local, double alpha, double shift,
                void* workspace, cudaStream_t s);
// preconditioner apply (precon.cu)
int precon_apply(mlffpc_ctx* ctx, const double* T, int64_t k, int64_t ld, double lam, double sign,
                 const double* r, double* z, double* u, cudaStream_t s, const double* Mk = nullptr,
                 const double* E = nullptr);

int ensure_reorth_scratch(mlffpc_ctx* ctx, int64_t k);
// peer-memory collectives (peer.cu)
bool peer_on(const mlffpc_ctx* ctx);
int peer_allreduce_kvec(mlffpc_ctx* ctx, double* w, int64_t k, cudaStream_t s);
int peer_wait(mlffpc_ctx* ctx, int ch, uint64_t epoch, const int* frozen, cudaStream_t s);
void peer_destroy(mlffpc_ctx* ctx);
int64_t peer_kmax(const mlffpc_ctx* ctx);
double* peer_yp_local(const mlffpc_ctx* ctx);
int symop_finish_peer(mlffpc_ctx* ctx, const double* x_local, double* y_local, int64_t nl, int64_t g_off, double alpha,
                      double shift, cudaStream_t s);
// W = X X^T + shift I (or X X^T - I) with extended-precision accumulation, summed over ranks (gramdd.cu)
int gram_dd(mlffpc_ctx* ctx, const double* X, int64_t m, int64_t n_cols, int64_t ldx, double* out, int64_t ld_out,
            double shift, bool minus_identity, cudaStream_t s, bool exact = false);
// internal dense building blocks (dense.cu), all on `s`
// nsplit > 1: split-K, slice z writes its partial product to C + z * c_zstride (beta applies to every slice)
int dgemm(bool transB, int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda,
          const double* B, int64_t ldb, double beta, double* C, int64_t ldc, bool lower_only,
          cudaStream_t s, int nsplit = 1, int64_t c_zstride = 0);
int dgemm_split_k(int64_t m, int64_t n, int64_t k, int num_sms);

}  // namespace mlffpc
