// Matrix-free Hessian-kernel operator  y = alpha (K v)_local + shift v_local.
//
// Reference math (torchtools.py:216-263, predict.py:185-229), q = sqrt5/sig:
//   beta_j = J_j v_j;   Delta = x_i - x_j^(p);  rho^ = q|Delta|;  e^ = 5/(3 sig^2) exp(-rho^)
//   f_i = sum_jp [ e^ q^2 (Delta . beta_jp) Delta - e^ (1 + rho^) beta_jp ];   (K v)_i = J_i^T f_i
// Three stages, nothing of size [B, S*M, D] is ever materialised:
//   1. prepare:  Bmat = [[X^(p), 1], [beta^(p), 0]]                      (2MS x (D+1))
//   2. pairs:    C1 = e^ q^2 (Delta.beta), C2 = e^ (1+rho^)  by direct differences (no Gram cancellation)
//   3. DMMA GEMM G = [C1 | C2] Bmat;  f_i = G[i,D] x_i - G[i,:D];  epilogue J_i^T f_i
#include "common.cuh"

namespace mlffpc {

__global__ void mv_prepare_kernel(int64_t MS, int S, int D, int N, int64_t ldb,
                                  const double* __restrict__ Xp, const double* __restrict__ R_d_desc,
                                  const int32_t* __restrict__ desc_perms, const int32_t* __restrict__ pair_a,
                                  const int32_t* __restrict__ pair_b, const double* __restrict__ v,
                                  double* __restrict__ Bmat) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = MS * (D + 1);
    if (t >= total) return;
    const int64_t jp = t / (D + 1);
    const int d = (int)(t % (D + 1));
    if (d == D) {
        Bmat[jp * ldb + D] = 1.0;
        Bmat[(MS + jp) * ldb + D] = 0.0;
        return;
    }
    const int p = (int)(jp % S);
    const int64_t j = jp / S;
    const int e = desc_perms[p * D + d];
    const int a = pair_a[e], b = pair_b[e];
    const double* g = R_d_desc + (j * D + e) * 3;
    const double* vj = v + j * 3 * N;
    const double beta = g[0] * (vj[3 * b] - vj[3 * a]) + g[1] * (vj[3 * b + 1] - vj[3 * a + 1]) +
                        g[2] * (vj[3 * b + 2] - vj[3 * a + 2]);
    Bmat[jp * ldb + d] = Xp[jp * D + d];
    Bmat[(MS + jp) * ldb + d] = beta;
}

constexpr int PT = 64;    // pair tile edge
constexpr int PDC = 16;   // descriptor chunk

__global__ void __launch_bounds__(256)
mv_pairs_kernel(const double* __restrict__ Xq, int64_t Ml, const double* __restrict__ Bmat, int64_t ldb,
                int64_t MS, int D, double q, double pref, double* __restrict__ Cmat, int64_t ldc) {
    __shared__ double xq[PDC][PT + 2], xj[PDC][PT + 2], bj[PDC][PT + 2];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int64_t i0 = (int64_t)blockIdx.y * PT, j0 = (int64_t)blockIdx.x * PT;
    double s2[4][4], tt[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) s2[a][b] = tt[a][b] = 0.0;

    const int lrow = tid >> 2, ld4 = (tid & 3) * 4;
    for (int dc = 0; dc < D; dc += PDC) {
        const bool iok = (i0 + lrow) < Ml, jok = (j0 + lrow) < MS;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int d = dc + ld4 + u;
            const bool dok = d < D;
            xq[ld4 + u][lrow] = (iok && dok) ? Xq[(i0 + lrow) * D + d] : 0.0;
            xj[ld4 + u][lrow] = (jok && dok) ? Bmat[(j0 + lrow) * ldb + d] : 0.0;
            bj[ld4 + u][lrow] = (jok && dok) ? Bmat[(MS + j0 + lrow) * ldb + d] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int dd = 0; dd < PDC; ++dd) {
            double xi[4], xv[4], bv[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) xi[a] = xq[dd][ty * 4 + a];
#pragma unroll
            for (int b = 0; b < 4; ++b) { xv[b] = xj[dd][tx + 16 * b]; bv[b] = bj[dd][tx + 16 * b]; }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const double dl = xi[a] - xv[b];
                    s2[a][b] = fma(dl, dl, s2[a][b]);
                    tt[a][b] = fma(dl, bv[b], tt[a][b]);
                }
        }
        __syncthreads();
    }
    const double q2 = q * q;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int64_t i = i0 + ty * 4 + a;
        if (i >= Ml) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int64_t j = j0 + tx + 16 * b;
            if (j >= MS) continue;
            const double rho = q * sqrt(s2[a][b]);
            const double e = pref * exp(-rho);
            Cmat[i * ldc + j] = e * q2 * tt[a][b];
            Cmat[i * ldc + MS + j] = e * (1.0 + rho);
        }
    }
}

// CTA per local point: G = sum of the split-K partials; f = G[i,D] x_i - G[i,:D];  y = alpha J_i^T f + shift v_local
__global__ void mv_epilogue_kernel(int N, int D, int64_t pt0, const double* __restrict__ R_desc,
                                   const double* __restrict__ R_d_desc, double* __restrict__ G,
                                   int64_t ldg, int nsplit, int64_t zstride, const double* __restrict__ v,
                                   double* __restrict__ y, double alpha, double shift) {
    const int64_t il = blockIdx.x, i = pt0 + il;
    const double* xi = R_desc + i * D;
    const double* gi = R_d_desc + i * D * 3;
    double* Gi = G + il * ldg;
    if (nsplit > 1) {  // fixed-order sum of the partial products into slice 0
        for (int d = threadIdx.x; d <= D; d += blockDim.x) {
            double t = Gi[d];
            for (int z = 1; z < nsplit; ++z) t += Gi[(int64_t)z * zstride + d];
            Gi[d] = t;
        }
        __syncthreads();
    }
    const double r1 = Gi[D];
    const int dim_i = 3 * N;
    for (int r = threadIdx.x; r < dim_i; r += blockDim.x) {
        const int A = r / 3, c = r % 3;
        double acc = 0.0;
        for (int B = 0; B < N; ++B) {
            if (B == A) continue;
            const int d = pair_index(A, B);
            const double f = fma(r1, xi[d], -Gi[d]);
            const double t = gi[d * 3 + c] * f;
            acc += (A < B) ? t : -t;
        }
        double out = alpha * acc;
        if (shift != 0.0) out = fma(shift, v[i * dim_i + r], out);
        y[il * dim_i + r] = out;
    }
}

struct MvWs {
    int64_t ldb, off_bmat, off_cmat, off_g, total;
    int nsplit;
};
static MvWs mv_layout(const mlffpc_ctx* c) {
    auto up = [](int64_t x) { return (x + 255) / 256 * 256; };
    MvWs w;
    const int64_t MS = c->M * c->S, Ml = c->pt1 - c->pt0;
    w.ldb = (c->D + 2) & ~(int64_t)1;
    int64_t o = 0;
    w.off_bmat = o; o = up(o + 2 * MS * w.ldb * 8);
    w.off_cmat = o; o = up(o + Ml * 2 * MS * 8);
    // the contraction G = [C1|C2] [X; beta] has few output tiles and a very long k: split k until the grid
    // covers the GPU about twice
    const int64_t tiles = ((Ml + 127) / 128) * ((c->D + 1 + 127) / 128);
    int64_t ns = (2 * (int64_t)c->num_sms + tiles - 1) / tiles;
    const int64_t max_by_k = (2 * MS) / 512;  // keep >= 512 columns of k per slice
    if (ns > max_by_k) ns = max_by_k;
    if (ns > 32) ns = 32;
    if (ns < 1) ns = 1;
    w.nsplit = (int)ns;
    w.off_g = o;    o = up(o + ns * Ml * w.ldb * 8);
    w.total = o + 256;
    return w;
}

int64_t matvec_free_ws_bytes(const mlffpc_ctx* ctx) { return mv_layout(ctx).total; }

int matvec_free(mlffpc_ctx* ctx, const double* v, double* y_local, double alpha, double shift,
                void* workspace, cudaStream_t s) {
    const MvWs w = mv_layout(ctx);
    char* base = (char*)(((uintptr_t)workspace + 255) / 256 * 256);
    double* Bmat = (double*)(base + w.off_bmat);
    double* Cmat = (double*)(base + w.off_cmat);
    double* G = (double*)(base + w.off_g);
    const int64_t MS = ctx->M * ctx->S, Ml = ctx->pt1 - ctx->pt0;
    const int D = ctx->D, N = ctx->N;
    const double q = sqrt(5.0) / ctx->sig, pref = 5.0 / (3.0 * ctx->sig * ctx->sig);

    const int64_t total = MS * (D + 1);
    mv_prepare_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(MS, ctx->S, D, N, w.ldb, ctx->Xp, ctx->R_d_desc,
                                                                     ctx->desc_perms, ctx->pair_a, ctx->pair_b, v, Bmat);
    MLFFPC_LAUNCH_CHECK();
    dim3 grid((unsigned)((MS + PT - 1) / PT), (unsigned)((Ml + PT - 1) / PT));
    MLFFPC_REQUIRE(grid.y <= 65535, "matvec_free: too many local points for this launch shape");
    mv_pairs_kernel<<<grid, 256, 0, s>>>(ctx->R_desc + ctx->pt0 * D, Ml, Bmat, w.ldb, MS, D, q, pref, Cmat, 2 * MS);
    MLFFPC_LAUNCH_CHECK();
    MLFFPC_TRY(dgemm(false, Ml, D + 1, 2 * MS, 1.0, Cmat, 2 * MS, Bmat, w.ldb, 0.0, G, w.ldb, false, s, w.nsplit,
                     Ml * w.ldb));
    int block = 32;
    while (block < 3 * N && block < 256) block <<= 1;
    mv_epilogue_kernel<<<(unsigned)Ml, block, 0, s>>>(N, D, ctx->pt0, ctx->R_desc, ctx->R_d_desc, G, w.ldb, w.nsplit,
                                                     Ml * w.ldb, v, y_local, alpha, shift);
    MLFFPC_LAUNCH_CHECK();
    return MLFFPC_OK;
}

}  // namespace mlffpc

using namespace mlffpc;

extern "C" {

int mlffpc_matvec_free_workspace_bytes(mlffpc_ctx* ctx, int64_t* bytes) {
    MLFFPC_REQUIRE(ctx && bytes && ctx->M > 0, "matvec_free_workspace_bytes: geometry not set");
    *bytes = matvec_free_ws_bytes(ctx);
    return MLFFPC_OK;
}

int mlffpc_matvec_free(mlffpc_ctx* ctx, const double* v, double* y_local, double alpha, double shift,
                       void* workspace, int64_t workspace_bytes, void* stream) {
    MLFFPC_REQUIRE(ctx && ctx->M > 0, "matvec_free: geometry not set");
    MLFFPC_REQUIRE(v && y_local && workspace, "matvec_free: NULL argument");
    MLFFPC_REQUIRE(workspace_bytes >= matvec_free_ws_bytes(ctx), "matvec_free: workspace too small");
    return matvec_free(ctx, v, y_local, alpha, shift, workspace, (cudaStream_t)stream);
}

}  // extern "C"
