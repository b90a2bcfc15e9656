// Symmetric assembled operator: every stored entry of K is read once and used twice.
//
// K is symmetric (K_ji = K_ij^T; reference train.py:207-210 `exploit_sym` relies on the same fact), so a
// matvec needs only one triangle.  The G x G grid of point blocks (the row-block partition of SURVEY 8e)
// is dealt out so that every unordered pair of blocks belongs to exactly one rank and every rank reads
// the same number of bytes -- half of its row block:
//     rank g:  the diagonal tile (g, g)                      (lower triangle read, by 32-row strips)
//              tiles (g, (g+d) % G) for d = 1 .. (G-1)/2      (read fully, used for rows AND columns)
//              if G is even, half of tile (g, (g+G/2) % G)    (g < G/2: the first half of its own points;
//                                                             g >= G/2: the second half of the partner's)
// One tile pass accumulates  y[rows] += A x[cols]  in registers and writes the per-strip column sums
// (A^T x[rows]) with plain coalesced stores (deterministic, no atomics); a second small kernel adds the
// strip partials.  Across ranks the full-length partial results are combined with one reduce-scatter
// (n doubles) -- the only extra collective next to the allgather of the search direction.
// HBM traffic per matvec and rank: 4 n n_local (1 + 2/32) bytes instead of 8 n n_local.
// The diagonal tile is stored packed (bands of 256 rows holding only the columns left of and including
// their diagonal blocks, symlayout.cuh), so the storage is half of the row block as well.  The tile passes
// run on the persistent TMA kernel of symtma.cu; symv_tile_kernel below is the register-staged fallback
// for a caller-owned square K whose base or pitch is not 16-byte aligned (mlffpc_symv).
#include <vector>

#include "common.cuh"
#include "symlayout.cuh"

namespace mlffpc {

constexpr int SYMV_THREADS = 256;
constexpr int SYMV_TR = 32;

__device__ __forceinline__ double2 ld_stream2(const double* p) {
    return __ldcs(reinterpret_cast<const double2*>(p));
}

// One CTA = one strip of TR rows of a tile.  Columns [0, n2) of the strip are two-sided (n2 = r0 for the
// diagonal tile, = nc for an off-diagonal tile); the diagonal tile also applies its TR x TR diagonal
// block one-sided.  K: tile base (row-major, even ld, 16-byte aligned); xr = x at the tile's first row,
// xc = x at the tile's first column (8-byte aligned only: c0 may be odd).
template <int TR>
__global__ void __launch_bounds__(SYMV_THREADS, (TR <= 32 ? 2 : 1))
symv_tile_kernel(const double* __restrict__ K, int64_t ld, int64_t nr, int64_t nc, int diag,
                 const double* __restrict__ xr, const double* __restrict__ xc, double* __restrict__ y1,
                 double* __restrict__ ws, int64_t ld_ws) {
    __shared__ double xs[TR];
    __shared__ double red[TR][SYMV_THREADS / 32];
    __shared__ double dsum[TR];
    const int tid = threadIdx.x;
    const int64_t s = (int64_t)gridDim.x - 1 - blockIdx.x;  // diagonal tile: longest strips first
    const int64_t r0 = s * TR;
    const int rows = (int)((nr - r0 < TR) ? (nr - r0) : TR);
    if (tid < TR) {
        xs[tid] = (tid < rows) ? xr[r0 + tid] : 0.0;
        dsum[tid] = 0.0;
    }
    __syncthreads();

    double acc[TR];
#pragma unroll
    for (int i = 0; i < TR; ++i) acc[i] = 0.0;
    const double* base = K + r0 * ld;
    const int64_t last = (int64_t)(rows - 1) * ld;  // rows past the end re-read the last valid row (xs = 0 there)

    const int64_t n2 = diag ? r0 : nc;  // r0 is even, so column pairs never straddle the diagonal block
    const int64_t nv = n2 >> 1;
    for (int64_t c2 = tid; c2 < nv; c2 += SYMV_THREADS) {
        double2 xv;
        xv.x = __ldg(xc + 2 * c2);
        xv.y = __ldg(xc + 2 * c2 + 1);
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
                const double xrow = xs[b * 8 + i];
                acc[b * 8 + i] = fma(kv[i].y, xv.y, fma(kv[i].x, xv.x, acc[b * 8 + i]));
                cacc.x = fma(kv[i].x, xrow, cacc.x);
                cacc.y = fma(kv[i].y, xrow, cacc.y);
            }
        }
        *reinterpret_cast<double2*>(ws + s * ld_ws + 2 * c2) = cacc;
    }
    if (!diag && (n2 & 1)) {  // odd tile width: last column, one thread per row
        const int64_t c = n2 - 1;
        if (tid < rows) {
            const double kx = base[(int64_t)tid * ld + c];
            dsum[tid] = kx * __ldg(xc + c);
            red[tid][0] = kx * xs[tid];
        }
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int i = 0; i < rows; ++i) t += red[i][0];
            ws[s * ld_ws + c] = t;
        }
        __syncthreads();
    }

    if (diag) {  // diagonal block, one-sided: 8 lanes per row
        const int cp = tid & 7;
        for (int r = tid >> 3; r < TR; r += SYMV_THREADS / 8) {
            double v = 0.0;
            if (r < rows)
                for (int c = cp; c < rows; c += 8) v = fma(base[(int64_t)r * ld + r0 + c], xs[c], v);
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
    if (tid < rows) {
        double v = dsum[tid];
#pragma unroll
        for (int i = 0; i < SYMV_THREADS / 32; ++i) v += red[tid][i];
        y1[r0 + tid] = v;
    }
}

// Adds the strip partials of one tile.
//   columns:  out_c[c] = post( [diag: y1[c]] + sum_{s >= s_min(c)} ws[s, c] ),  s_min = c/TR + 1 (diag) or 0
//   rows (off-diagonal tiles only):  out_r[r] += y1[r]
// post(v) = alpha v + shift x_shift[c]  when x_shift != NULL (single-rank finish), else v.
template <int TR>
__global__ void symv_reduce_kernel(const double* __restrict__ y1, const double* __restrict__ ws, int64_t ld_ws,
                                   int64_t nr, int64_t nc, int64_t nstrips, int diag, double* __restrict__ out_c,
                                   double* __restrict__ out_r, const double* __restrict__ x_shift, double alpha,
                                   double shift) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nc) {
        const int64_t c = t;
        double acc = diag ? y1[c] : 0.0;
        int64_t s = diag ? (c / TR + 1) : 0;
        const double* p = ws + c;
        for (; s + 7 < nstrips; s += 8) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcs(p + (s + u) * ld_ws);
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += v[u];
        }
        for (; s < nstrips; ++s) acc += __ldcs(p + s * ld_ws);
        if (x_shift) {
            acc *= alpha;
            if (shift != 0.0) acc = fma(shift, x_shift[c], acc);
        }
        out_c[c] = acc;
    } else if (!diag && t < nc + nr) {
        const int64_t r = t - nc;
        out_r[r] += y1[r];
    }
}

__global__ void symop_finish_kernel(const double* __restrict__ q, const double* __restrict__ x_local,
                                    double* __restrict__ y, int64_t n, double alpha, double shift) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) y[t] = fma(shift, x_local[t], alpha * q[t]);
}

// ---- tile plan -------------------------------------------------------------------------------------
struct SymTile {
    int64_t i_pt0, i_pt1, j_pt0, j_pt1;  // point ranges (rows, columns)
    int64_t nr, nc, ld, off;             // rows, columns, leading dimension (0: packed), element offset
    int diag;                            // diagonal tiles use the packed band layout of symlayout.cuh
};

static inline int64_t up32(int64_t x) { return (x + 31) / 32 * 32; }

static std::vector<SymTile> symop_plan(const mlffpc_ctx* c, int64_t* total_elems) {
    std::vector<SymTile> tiles;
    const int W = c->lay_world, g = c->lay_rank;
    const int64_t M = c->M, di = c->dim_i;
    const int64_t ppr = (M + W - 1) / W;
    auto blk0 = [&](int b) { return (int64_t)b * ppr < M ? (int64_t)b * ppr : M; };
    auto blk1 = [&](int b) { return (int64_t)(b + 1) * ppr < M ? (int64_t)(b + 1) * ppr : M; };
    int64_t off = 0;
    auto add = [&](int64_t i0, int64_t i1, int64_t j0, int64_t j1, int diag) {
        if (i1 <= i0 || j1 <= j0) return;
        SymTile t;
        t.i_pt0 = i0; t.i_pt1 = i1; t.j_pt0 = j0; t.j_pt1 = j1;
        t.nr = (i1 - i0) * di; t.nc = (j1 - j0) * di;
        t.ld = diag ? 0 : ((t.nc + 1) & ~(int64_t)1);
        t.off = off; t.diag = diag;
        off = up32(off + (diag ? st_packed_elems(t.nr) : t.nr * t.ld));
        tiles.push_back(t);
    };
    add(blk0(g), blk1(g), blk0(g), blk1(g), 1);
    for (int d = 1; d <= (W - 1) / 2; ++d) {
        const int h = (g + d) % W;
        add(blk0(g), blk1(g), blk0(h), blk1(h), 0);
    }
    if (W > 1 && W % 2 == 0) {
        const int h = (g + W / 2) % W;
        if (g < W / 2) {
            const int64_t half = (blk1(g) - blk0(g) + 1) / 2;
            add(blk0(g), blk0(g) + half, blk0(h), blk1(h), 0);
        } else {
            const int64_t half = (blk1(h) - blk0(h) + 1) / 2;
            add(blk0(g), blk1(g), blk0(h) + half, blk1(h), 0);
        }
    }
    if (total_elems) *total_elems = off;
    return tiles;
}

struct SymWs {
    int64_t n_pad, off_tile, off_yp, off_q, total;
};
static SymWs symop_ws_layout(const mlffpc_ctx* c, const std::vector<SymTile>& tiles) {
    auto up = [](int64_t x) { return (x + 255) / 256 * 256; };
    SymWs w;
    const int W = c->lay_world;
    w.n_pad = ((c->M + W - 1) / W) * c->dim_i;
    int64_t sum_ws = 0;  // the one-launch pass keeps the partials of all tiles at once
    for (const auto& t : tiles) sum_ws += symv_tma_ws_doubles(t.nr, t.nc);
    int64_t o = 0;
    w.off_tile = o; o = up(o + sum_ws * 8);
    w.off_yp = o; o = up(o + (int64_t)W * w.n_pad * 8);
    w.off_q = o;  o = up(o + w.n_pad * 8);
    w.total = o + 512;
    return w;
}

int64_t symop_storage_elems(const mlffpc_ctx* ctx) {
    int64_t e = 0;
    symop_plan(ctx, &e);
    return e;
}
int64_t symop_ws_bytes(const mlffpc_ctx* ctx) {
    return symop_ws_layout(ctx, symop_plan(ctx, nullptr)).total;
}

// register-staged fallback: strip pass + partial reduction on a square row-major K
static int symv_square_fallback(const double* K, int64_t ld, int64_t n, const double* x, double* y1, double* ws,
                                double* y, double alpha, double shift, cudaStream_t s) {
    const int64_t nstrips = (n + SYMV_TR - 1) / SYMV_TR;
    const int64_t ld_ws = (n + 1) & ~(int64_t)1;
    symv_tile_kernel<SYMV_TR><<<(unsigned)nstrips, SYMV_THREADS, 0, s>>>(K, ld, n, n, 1, x, x, y1, ws, ld_ws);
    MLFFPC_LAUNCH_CHECK();
    symv_reduce_kernel<SYMV_TR><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(y1, ws, ld_ws, n, n, nstrips, 1, y, nullptr,
                                                                          x, alpha, shift);
    MLFFPC_LAUNCH_CHECK();
    return MLFFPC_OK;
}

int64_t symv_ws_bytes(int64_t n) {
    const int64_t nstrips = (n + SYMV_TR - 1) / SYMV_TR;
    const int64_t ld_ws = (n + 1) & ~(int64_t)1;
    const int64_t fallback = nstrips * ld_ws + n + 128;
    const int64_t tma = symv_tma_ws_doubles(n, n);
    return (fallback > tma ? fallback : tma) * 8 + 512;
}

// y = alpha * K x + shift * x for one symmetric square K (only the lower triangle by strips is read)
int launch_symv(mlffpc_ctx* ctx, const double* K, int64_t n, int64_t ld, const double* x, double* y, double alpha,
                double shift, void* workspace, cudaStream_t s) {
    double* wsd = (double*)(((uintptr_t)workspace + 255) / 256 * 256);
    if (symv_tma_usable(K, ld, n))
        return symv_tile_tma(ctx, K, ld, n, n, 1, 0, x, x, wsd, y, nullptr, x, alpha, shift, s);
    MLFFPC_REQUIRE(ld % 2 == 0 && ((uintptr_t)K % 16 == 0),
                   "symv: K must be 16-byte aligned with an even leading dimension");
    return symv_square_fallback(K, ld, n, x, wsd, wsd + ((n + 63) / 32 * 32), y, alpha, shift, s);
}

// The sharded symmetric operator on this rank's tiles.  With partial_out != NULL the full-length partial
// result (world * n_pad doubles) is left there and no collective is issued (rank emulation for tests).
int symop_apply(mlffpc_ctx* ctx, const double* Ksym, const double* x_full, double* y_local, double alpha,
                double shift, void* workspace, double* partial_out, cudaStream_t s) {
    const std::vector<SymTile> tiles = symop_plan(ctx, nullptr);
    const SymWs w = symop_ws_layout(ctx, tiles);
    char* base = (char*)(((uintptr_t)workspace + 255) / 256 * 256);
    double* wt = (double*)(base + w.off_tile);
    const int W = ctx->lay_world;
    const int64_t di = ctx->dim_i;
    if (W == 1 && !partial_out) {
        const SymTile& t = tiles[0];
        return symv_tile_tma(ctx, Ksym + t.off, 0, t.nr, t.nc, 1, 1, x_full, x_full, wt, y_local, nullptr, x_full,
                             alpha, shift, s);
    }
    // with mapped peer buffers the partial products are written where the other ranks can pull them
    const bool peer = !partial_out && peer_on(ctx) && ctx->comm.world == W;
    double* yp = partial_out ? partial_out : (peer ? peer_yp_local(ctx) : (double*)(base + w.off_yp));
    MLFFPC_CUDA(cudaMemsetAsync(yp, 0, (size_t)W * w.n_pad * 8, s));
    if (ctx->symop_multi && tiles.size() <= 8) {
        SymTileIn in[8];
        for (size_t i = 0; i < tiles.size(); ++i) {
            const SymTile& t = tiles[i];
            in[i].K = Ksym + t.off; in[i].ld = t.ld; in[i].nr = t.nr; in[i].nc = t.nc;
            in[i].row_off = (t.i_pt0 - tiles[0].i_pt0) * di;
            in[i].diag = t.diag; in[i].packed = t.diag;
            in[i].xr = x_full + t.i_pt0 * di; in[i].xc = x_full + t.j_pt0 * di;
            in[i].out_c = yp + t.j_pt0 * di;
        }
        MLFFPC_TRY(symv_tiles_tma(ctx, (int)tiles.size(), in, wt, s));
    } else {
        for (const auto& t : tiles) {
            const double* xr = x_full + t.i_pt0 * di;
            const double* xc = x_full + t.j_pt0 * di;
            MLFFPC_TRY(symv_tile_tma(ctx, Ksym + t.off, t.ld, t.nr, t.nc, t.diag, t.diag, xr, xc, wt, yp + t.j_pt0 * di,
                                     yp + t.i_pt0 * di, nullptr, 1.0, 0.0, s));
        }
    }
    if (partial_out) return MLFFPC_OK;
    MLFFPC_REQUIRE(ctx->comm.world == W, "symop_apply: the tile layout (%d ranks) needs a communicator of that size", W);
    const int64_t nl = ctx->n_local();
    if (peer) return symop_finish_peer(ctx, x_full + ctx->row0(), y_local, nl, ctx->row0(), alpha, shift, s);
    double* q = (double*)(base + w.off_q);
    MLFFPC_TRY(comm_reduce_scatter_sum(ctx->comm, yp, q, (size_t)w.n_pad, s));
    symop_finish_kernel<<<(unsigned)((nl + 255) / 256), 256, 0, s>>>(q, x_full + ctx->row0(), y_local, nl, alpha, shift);
    MLFFPC_LAUNCH_CHECK();
    return MLFFPC_OK;
}

int symop_assemble(mlffpc_ctx* ctx, double* Ksym, cudaStream_t s) {
    const std::vector<SymTile> tiles = symop_plan(ctx, nullptr);
    ProfWindow pw = prof_window("assemble");
    pw.step(pw.first);
    int st = MLFFPC_OK;
    for (const auto& t : tiles) {
        st = assemble_tile(ctx, t.i_pt0, t.i_pt1, t.j_pt0, t.j_pt1, Ksym + t.off, t.ld, t.diag ? 1 : 0, s);
        if (st != MLFFPC_OK) break;
    }
    pw.end();
    return st;
}

}  // namespace mlffpc

using namespace mlffpc;

extern "C" {

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
    return launch_symv(ctx, K, n, ld, x, y, alpha, shift, workspace, (cudaStream_t)stream);
}

int mlffpc_symop_storage_elems(mlffpc_ctx* ctx, int64_t* elems) {
    MLFFPC_REQUIRE(ctx && elems && ctx->M > 0, "symop_storage_elems: geometry not set");
    *elems = symop_storage_elems(ctx);
    return MLFFPC_OK;
}

int mlffpc_symop_tiles(mlffpc_ctx* ctx, int64_t* out, int64_t max_tiles, int64_t* n_tiles) {
    MLFFPC_REQUIRE(ctx && n_tiles && ctx->M > 0, "symop_tiles: geometry not set");
    const std::vector<SymTile> tiles = symop_plan(ctx, nullptr);
    *n_tiles = (int64_t)tiles.size();
    if (out) {
        MLFFPC_REQUIRE(max_tiles >= (int64_t)tiles.size(), "symop_tiles: output too small");
        for (size_t i = 0; i < tiles.size(); ++i) {
            const SymTile& t = tiles[i];
            int64_t* o = out + 8 * i;
            o[0] = t.i_pt0; o[1] = t.i_pt1; o[2] = t.j_pt0; o[3] = t.j_pt1; o[4] = t.ld; o[5] = t.off; o[6] = t.diag; o[7] = 0;
        }
    }
    return MLFFPC_OK;
}

int mlffpc_symop_assemble(mlffpc_ctx* ctx, double* Ksym, void* stream) {
    MLFFPC_REQUIRE(ctx && Ksym && ctx->M > 0, "symop_assemble: geometry not set or NULL output");
    MLFFPC_REQUIRE((uintptr_t)Ksym % 16 == 0, "symop_assemble: storage must be 16-byte aligned");
    return symop_assemble(ctx, Ksym, (cudaStream_t)stream);
}

int mlffpc_symop_workspace_bytes(mlffpc_ctx* ctx, int64_t* bytes) {
    MLFFPC_REQUIRE(ctx && bytes && ctx->M > 0, "symop_workspace_bytes: geometry not set");
    *bytes = symop_ws_bytes(ctx);
    return MLFFPC_OK;
}

int mlffpc_symop_apply(mlffpc_ctx* ctx, const double* Ksym, const double* x_full, double* y_local, double alpha,
                       double shift, void* workspace, int64_t workspace_bytes, double* partial_out, void* stream) {
    MLFFPC_REQUIRE(ctx && ctx->M > 0, "symop_apply: geometry not set");
    MLFFPC_REQUIRE(Ksym && x_full && (y_local || partial_out) && workspace, "symop_apply: NULL argument");
    MLFFPC_REQUIRE(workspace_bytes >= symop_ws_bytes(ctx), "symop_apply: workspace too small");
    return symop_apply(ctx, Ksym, x_full, y_local, alpha, shift, workspace, partial_out, (cudaStream_t)stream);
}

}  // extern "C"
