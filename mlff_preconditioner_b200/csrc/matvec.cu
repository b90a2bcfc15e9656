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
                                  const double* __restrict__ beta_in, double* __restrict__ Bmat) {
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
    double beta;
    if (beta_in) {  // beta_j = J_j alpha_j given (model['R_d_desc_alpha'], train.py:640-645)
        beta = beta_in[j * D + e];
    } else {
        const int a = pair_a[e], b = pair_b[e];
        const double* g = R_d_desc + (j * D + e) * 3;
        const double* vj = v + j * 3 * N;
        beta = g[0] * (vj[3 * b] - vj[3 * a]) + g[1] * (vj[3 * b + 1] - vj[3 * a + 1]) +
               g[2] * (vj[3 * b + 2] - vj[3 * a + 2]);
    }
    Bmat[jp * ldb + d] = Xp[jp * D + d];
    Bmat[(MS + jp) * ldb + d] = beta;
}

// ---- epilogue arithmetic of a pair -------------------------------------------------------------------------
// rho^ = q sqrt(s2), e^ = pref exp(-rho^).  With D = 36 the generic sqrt / exp (slow-path calls, range checks) were as
// many instructions per pair as the descriptor loop itself (ncu, profiles/r02l): these versions assume what holds
// here -- s2 >= 0 finite, the exponent <= 0 -- and stay within 2 ulp (parity bound of the operator: 1e-10).
__device__ __forceinline__ double pair_sqrt(double x) {
    // Newton on y ~ 1/sqrt(x) from the 64-bit rsqrt seed: g -> sqrt(x), h -> 1/(2 sqrt(x))
    const double xs = fmax(x, 1e-280);  // x = 0 (a point against itself): sqrt = 1e-140 ~ 0 instead of 0 * inf
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(xs));
    double g = xs * y, h = 0.5 * y;
    double r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-g, g, xs);       // last correction on the residual x - g^2
    return fma(r, h, g);
}
__device__ __forceinline__ double pair_exp_neg(double x) {  // exp(x) for x <= 0
    if (x < -700.0) return 0.0;  // below 1e-304: no contribution (and the exponent arithmetic below stays normal)
    const double SHIFT = 6755399441055744.0;  // 1.5 * 2^52: rounds to nearest integer in the low word
    const double t = fma(x, 1.4426950408889634074, SHIFT);
    const int n = __double2loint(t);
    const double nf = t - SHIFT;
    double r = fma(nf, -6.93147180369123816490e-01, x);   // Cody-Waite: ln2 = hi + lo
    r = fma(nf, -1.90821492927058770002e-10, r);
    // |r| <= ln2 / 2: Taylor to r^13 (remainder 4e-18)
    double p = 1.6059043836821613e-10;                   // 1/13!
    p = fma(p, r, 2.08767569878681e-09);                  // 1/12!
    p = fma(p, r, 2.505210838544172e-08);                 // 1/11!
    p = fma(p, r, 2.755731922398589e-07);                 // 1/10!
    p = fma(p, r, 2.7557319223985893e-06);                // 1/9!
    p = fma(p, r, 2.48015873015873e-05);                  // 1/8!
    p = fma(p, r, 1.984126984126984e-04);                 // 1/7!
    p = fma(p, r, 1.388888888888889e-03);                 // 1/6!
    p = fma(p, r, 8.333333333333333e-03);                 // 1/5!
    p = fma(p, r, 4.1666666666666664e-02);                // 1/4!
    p = fma(p, r, 1.6666666666666666e-01);                // 1/3!
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));  // p * 2^n, n >= -1010
}

constexpr int PT = 64;    // pair tile edge
constexpr int PDC = 16;   // descriptor chunk

__global__ void __launch_bounds__(256)
mv_pairs_kernel(const double* __restrict__ Xq, int64_t Ml, const double* __restrict__ Bmat, int64_t ldb,
                int64_t MS, int D, double q, double pref, double* __restrict__ Cmat, int64_t ldc,
                double* __restrict__ Epart) {
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
        // a ragged last chunk stops at the next multiple of 4 (D = 36: 16 + 16 + 4 instead of 48 descriptor steps)
        const int dd_end = (D - dc >= PDC) ? PDC : ((D - dc + 3) & ~3);
#pragma unroll 4
        for (int dd = 0; dd < dd_end; ++dd) {
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
    const bool interior = (i0 + PT <= Ml) && (j0 + PT <= MS);  // no per-pair bounds tests in full tiles
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int64_t i = i0 + ty * 4 + a;
        double* crow = Cmat + i * ldc + j0 + tx;
        double esum = 0.0;  // energy: sum_j e^ (1 + rho^) (Delta . beta)   (torchtools.py:268, predict.py:207)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            if (!interior && (i >= Ml || j0 + tx + 16 * b >= MS)) continue;
            const double rho = q * pair_sqrt(s2[a][b]);
            const double e = pref * pair_exp_neg(-rho);
            const double c2 = fma(e, rho, e);
            crow[16 * b] = e * q2 * tt[a][b];
            crow[MS + 16 * b] = c2;
            esum = fma(c2, tt[a][b], esum);
        }
        if (Epart) {  // the 16 threads of a row sit in one half-warp
            esum += __shfl_xor_sync(0xffffffffu, esum, 8);
            esum += __shfl_xor_sync(0xffffffffu, esum, 4);
            esum += __shfl_xor_sync(0xffffffffu, esum, 2);
            esum += __shfl_xor_sync(0xffffffffu, esum, 1);
            if (tx == 0 && i < Ml) Epart[i * gridDim.x + blockIdx.x] = esum;
        }
    }
}

// Second-generation pair kernel: 128 queries x (16 NB) training rows per CTA, 8 x NB pairs per thread, descriptor chunks
// of 16 staged by cp.async into a 3-stage ring (transposed to [d][row] so that a thread's 8 query values and its NB
// training values are 128-bit shared loads).  Same arithmetic per pair as mv_pairs_kernel, same outputs.
//   NB = 4 (222 registers, 8 warps per SM): 8 LDS.128 for 96 FP64 instructions per descriptor; ncu (r02l, D = 210): FP64
//          pipe 68 % busy, the rest are fixed-latency waits that two warps per scheduler cannot cover
//   NB = 2 (<= 128 registers, two CTAs = 16 warps per SM): 6 LDS.128 for 48 FP64 instructions, twice the warps
constexpr int P2_TI = 128, P2_DC = 16, P2_STAGES = 3, P2_THREADS = 256;
constexpr int P2_QS = P2_TI + 2;                                          // padded row length (even: 16-byte aligned pairs)
template <int NB> struct P2Cfg {
    static constexpr int TJ = 16 * NB, JS = TJ + 2;
    static constexpr int STAGE = P2_DC * (P2_QS + 2 * JS);                // doubles per stage
    static constexpr size_t SMEM = (size_t)P2_STAGES * STAGE * sizeof(double);
};

__device__ __forceinline__ void p2_cp8(void* dst, const void* src, bool ok) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int sz = ok ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(sz));
}

template <int NB>
__global__ void __launch_bounds__(P2_THREADS, NB == 2 ? 2 : 1)
mv_pairs2_kernel(const double* __restrict__ Xq, int64_t Ml, const double* __restrict__ Bmat, int64_t ldb,
                 int64_t MS, int D, double q, double pref, double* __restrict__ Cmat, int64_t ldc,
                 double* __restrict__ Epart, int n_epart) {
    using C = P2Cfg<NB>;
    extern __shared__ double p2_smem[];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int64_t i0 = (int64_t)blockIdx.y * P2_TI, j0 = (int64_t)blockIdx.x * C::TJ;
    double s2[8][NB], tt[8][NB];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < NB; ++b) s2[a][b] = tt[a][b] = 0.0;

    const int nchunks = (D + P2_DC - 1) / P2_DC;
    auto load_stage = [&](int stage, int c) {
        double* xq = p2_smem + (size_t)stage * C::STAGE;
        double* xj = xq + P2_DC * P2_QS;
        double* bj = xj + P2_DC * C::JS;
        const int d0 = c * P2_DC;
        // queries: 128 rows x 16 d; a thread copies 8 elements (row = e / 16 keeps the global reads of a warp in rows of 16)
        for (int e = tid; e < P2_TI * P2_DC; e += P2_THREADS) {
            const int r = e / P2_DC, d = e % P2_DC;
            const bool ok = (i0 + r < Ml) && (d0 + d < D);
            p2_cp8(xq + d * P2_QS + r, ok ? (Xq + (i0 + r) * D + d0 + d) : Xq, ok);
        }
        for (int e = tid; e < C::TJ * P2_DC; e += P2_THREADS) {
            const int r = e / P2_DC, d = e % P2_DC;
            const bool ok = (j0 + r < MS) && (d0 + d < D);
            p2_cp8(xj + d * C::JS + r, ok ? (Bmat + (j0 + r) * ldb + d0 + d) : Bmat, ok);
            p2_cp8(bj + d * C::JS + r, ok ? (Bmat + (MS + j0 + r) * ldb + d0 + d) : Bmat, ok);
        }
    };
#pragma unroll
    for (int st = 0; st < P2_STAGES - 1; ++st) {
        if (st < nchunks) load_stage(st, st);
        asm volatile("cp.async.commit_group;");
    }
    for (int c = 0; c < nchunks; ++c) {
        asm volatile("cp.async.wait_group %0;" ::"n"(P2_STAGES - 2));
        __syncthreads();
        {
            const int nc = c + P2_STAGES - 1;
            if (nc < nchunks) load_stage(nc % P2_STAGES, nc);
            asm volatile("cp.async.commit_group;");
        }
        const double* xq = p2_smem + (size_t)(c % P2_STAGES) * C::STAGE;
        const double* xj = xq + P2_DC * P2_QS;
        const double* bj = xj + P2_DC * C::JS;
        // the last chunk may hold fewer than 16 descriptors: stop at the next multiple of 4 (zero-filled beyond D)
        const int dd_end = (D - c * P2_DC >= P2_DC) ? P2_DC : ((D - c * P2_DC + 3) & ~3);
#pragma unroll 4
        for (int dd = 0; dd < dd_end; ++dd) {
            double xi[8], xv[NB], bv[NB];
            const double2* q2 = reinterpret_cast<const double2*>(xq + dd * P2_QS + ty * 8);
#pragma unroll
            for (int a = 0; a < 4; ++a) { const double2 t = q2[a]; xi[2 * a] = t.x; xi[2 * a + 1] = t.y; }
#pragma unroll
            for (int h = 0; h < NB / 2; ++h) {  // columns j0 + 32 h + 2 tx + {0, 1}
                const double2 t0 = *reinterpret_cast<const double2*>(xj + dd * C::JS + 32 * h + 2 * tx);
                const double2 u0 = *reinterpret_cast<const double2*>(bj + dd * C::JS + 32 * h + 2 * tx);
                xv[2 * h] = t0.x; xv[2 * h + 1] = t0.y;
                bv[2 * h] = u0.x; bv[2 * h + 1] = u0.y;
            }
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const double dl = xi[a] - xv[b];
                    s2[a][b] = fma(dl, dl, s2[a][b]);
                    tt[a][b] = fma(dl, bv[b], tt[a][b]);
                }
        }
    }
    asm volatile("cp.async.wait_group 0;");
    const double q2c = q * q;
    const bool interior = (i0 + P2_TI <= Ml) && (j0 + C::TJ <= MS);
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int64_t i = i0 + ty * 8 + a;
        double* crow = Cmat + i * ldc + j0 + 2 * tx;
        double esum = 0.0;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int jo = (b >> 1) * 32 + (b & 1);  // column offset next to 2 tx
            if (!interior && (i >= Ml || j0 + 2 * tx + jo >= MS)) continue;
            const double rho = q * pair_sqrt(s2[a][b]);
            const double e = pref * pair_exp_neg(-rho);
            const double c2 = fma(e, rho, e);
            crow[jo] = e * q2c * tt[a][b];
            crow[MS + jo] = c2;
            esum = fma(c2, tt[a][b], esum);
        }
        if (Epart) {  // the 16 threads of a row sit in one half-warp
            esum += __shfl_xor_sync(0xffffffffu, esum, 8);
            esum += __shfl_xor_sync(0xffffffffu, esum, 4);
            esum += __shfl_xor_sync(0xffffffffu, esum, 2);
            esum += __shfl_xor_sync(0xffffffffu, esum, 1);
            if (tx == 0 && i < Ml) Epart[i * n_epart + blockIdx.x] = esum;
        }
    }
}

// columns per energy partial: the narrowest pair tile (every kernel writes one partial per tile of its own width; the
// caller sizes the buffer for 32-column tiles)
constexpr int P_EPART_COLS = 32;

// pair stage.  Option "pairs_kernel": 1 = first kernel (64 x 64 tiles), 2 = 128 x 64 tiles, 3 = 128 x 32 tiles with two
// CTAs per SM, 0 = by descriptor length (measured, profiles/r02l-m)
static int launch_pairs(mlffpc_ctx* ctx, const double* Xq, int64_t Mq, const double* Bmat, int64_t ldb, int64_t MS, int D,
                        double q, double pref, double* Cmat, double* Epart, int* n_epart_out, cudaStream_t s) {
    int which = ctx->pairs_kernel;
    if (which == 0) which = D < 64 ? 1 : 3;
    if (which == 1) {
        dim3 grid((unsigned)((MS + PT - 1) / PT), (unsigned)((Mq + PT - 1) / PT));
        MLFFPC_REQUIRE(grid.y <= 65535, "pairs: too many query points for this launch shape");
        mv_pairs_kernel<<<grid, 256, 0, s>>>(Xq, Mq, Bmat, ldb, MS, D, q, pref, Cmat, 2 * MS, Epart);
        MLFFPC_LAUNCH_CHECK();
        if (n_epart_out) *n_epart_out = (int)grid.x;
        return MLFFPC_OK;
    }
    if (which == 2) {
        using C = P2Cfg<4>;
        dim3 grid((unsigned)((MS + C::TJ - 1) / C::TJ), (unsigned)((Mq + P2_TI - 1) / P2_TI));
        MLFFPC_REQUIRE(grid.y <= 65535, "pairs: too many query points for this launch shape");
        if (!(ctx->pairs2_attr & 1)) {
            MLFFPC_CUDA(cudaFuncSetAttribute(mv_pairs2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
            ctx->pairs2_attr |= 1;
        }
        mv_pairs2_kernel<4><<<grid, P2_THREADS, C::SMEM, s>>>(Xq, Mq, Bmat, ldb, MS, D, q, pref, Cmat, 2 * MS, Epart, (int)grid.x);
        MLFFPC_LAUNCH_CHECK();
        if (n_epart_out) *n_epart_out = (int)grid.x;
        return MLFFPC_OK;
    }
    using C = P2Cfg<2>;
    dim3 grid((unsigned)((MS + C::TJ - 1) / C::TJ), (unsigned)((Mq + P2_TI - 1) / P2_TI));
    MLFFPC_REQUIRE(grid.y <= 65535, "pairs: too many query points for this launch shape");
    if (!(ctx->pairs2_attr & 2)) {
        MLFFPC_CUDA(cudaFuncSetAttribute(mv_pairs2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        ctx->pairs2_attr |= 2;
    }
    mv_pairs2_kernel<2><<<grid, P2_THREADS, C::SMEM, s>>>(Xq, Mq, Bmat, ldb, MS, D, q, pref, Cmat, 2 * MS, Epart, (int)grid.x);
    MLFFPC_LAUNCH_CHECK();
    if (n_epart_out) *n_epart_out = (int)grid.x;
    return MLFFPC_OK;
}

// CTA per local point: G = sum of the split-K partials; f = G[i,D] x_i - G[i,:D];  y = alpha J_i^T f + shift v_local
// R_desc / R_d_desc / v are indexed by pt0 + il (training mode: the context's tables; prediction: the query tables with
// pt0 = 0).  Epart != NULL: E[il] = sum of the pair kernel's per-tile energy partials (fixed order).
__global__ void mv_epilogue_kernel(int N, int D, int64_t pt0, const double* __restrict__ R_desc,
                                   const double* __restrict__ R_d_desc, double* __restrict__ G,
                                   int64_t ldg, int nsplit, int64_t zstride, const double* __restrict__ v,
                                   double* __restrict__ y, double alpha, double shift,
                                   const double* __restrict__ Epart, int n_epart, double* __restrict__ E_out) {
    const int64_t il = blockIdx.x, i = pt0 + il;
    if (Epart && threadIdx.x == 0) {
        double e = 0.0;
        for (int c = 0; c < n_epart; ++c) e += Epart[il * n_epart + c];
        E_out[il] = alpha * e;
    }
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
    // the contraction G = [C1|C2] [X; beta] has few output tiles and a very long k: split k into whole waves
    const int64_t ns = dgemm_split_k(Ml, c->D + 1, 2 * MS, c->num_sms);
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
                                                                     ctx->desc_perms, ctx->pair_a, ctx->pair_b, v, nullptr, Bmat);
    MLFFPC_LAUNCH_CHECK();
    MLFFPC_TRY(launch_pairs(ctx, ctx->R_desc + ctx->pt0 * D, Ml, Bmat, w.ldb, MS, D, q, pref, Cmat, nullptr, nullptr, s));
    MLFFPC_TRY(dgemm(false, Ml, D + 1, 2 * MS, 1.0, Cmat, 2 * MS, Bmat, w.ldb, 0.0, G, w.ldb, false, s, w.nsplit,
                     Ml * w.ldb));
    int block = 32;
    while (block < 3 * N && block < 256) block <<= 1;
    mv_epilogue_kernel<<<(unsigned)Ml, block, 0, s>>>(N, D, ctx->pt0, ctx->R_desc, ctx->R_d_desc, G, w.ldb, w.nsplit,
                                                     Ml * w.ldb, v, y_local, alpha, shift, nullptr, 0, nullptr);
    MLFFPC_LAUNCH_CHECK();
    return MLFFPC_OK;
}

// ---- prediction for arbitrary query geometries (GDMLPredict.predict, predict.py:997-1110; torchtools.py:172-272) ----
// Descriptor x_d = 1 / |r_a - r_b| and its compressed Jacobian g_d = (r_a - r_b) / |r_a - r_b|^3 for the pairs
// d <-> (a_d > b_d) in np.tril_indices order (utils/desc.py:112-200, :292-358).  One thread per (geometry, pair).
__global__ void desc_from_r_kernel(const double* __restrict__ R, int64_t B, int N, int D, double* __restrict__ R_desc,
                                   double* __restrict__ R_d_desc) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * D) return;
    const int64_t m = t / D;
    const int d = (int)(t % D);
    int a = (int)((1.0 + sqrt(1.0 + 8.0 * (double)d)) * 0.5);
    while (a * (a - 1) / 2 > d) --a;
    while ((a + 1) * a / 2 <= d) ++a;
    const int b = d - a * (a - 1) / 2;
    const double* ra = R + (m * N + a) * 3;
    const double* rb = R + (m * N + b) * 3;
    const double dx = ra[0] - rb[0], dy = ra[1] - rb[1], dz = ra[2] - rb[2];
    // no FMA contraction: the same roundings as the host's sum of squares
    const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    const double r = sqrt(r2);
    const double r3 = __dmul_rn(__dmul_rn(r, r), r);
    R_desc[t] = 1.0 / r;
    R_d_desc[t * 3 + 0] = dx / r3;
    R_d_desc[t * 3 + 1] = dy / r3;
    R_d_desc[t * 3 + 2] = dz / r3;
}

// beta[m, d] = g[m, d, :] . (v[m, b_d, :] - v[m, a_d, :])   (desc.d_desc_dot_vec, utils/desc.py:394-405; the
// model's R_d_desc_alpha, train.py:640-645)
__global__ void jv_kernel(int64_t M, int D, int N, const double* __restrict__ R_d_desc, const int32_t* __restrict__ pair_a,
                          const int32_t* __restrict__ pair_b, const double* __restrict__ v, double* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= M * D) return;
    const int64_t m = t / D;
    const int d = (int)(t % D);
    const int a = pair_a[d], b = pair_b[d];
    const double* g = R_d_desc + t * 3;
    const double* vm = v + m * 3 * N;
    out[t] = g[0] * (vm[3 * b] - vm[3 * a]) + g[1] * (vm[3 * b + 1] - vm[3 * a + 1]) + g[2] * (vm[3 * b + 2] - vm[3 * a + 2]);
}

struct PredWs {
    int64_t ldb, off_bmat, off_cmat, off_g, off_epart, total;
    int nsplit, ncb;
};
static PredWs pred_layout(const mlffpc_ctx* c, int64_t B) {
    auto up = [](int64_t x) { return (x + 255) / 256 * 256; };
    PredWs w;
    const int64_t MS = c->M * c->S;
    w.ldb = (c->D + 2) & ~(int64_t)1;
    int64_t o = 0;
    w.off_bmat = o; o = up(o + 2 * MS * w.ldb * 8);
    w.off_cmat = o; o = up(o + B * 2 * MS * 8);
    const int64_t ns = dgemm_split_k(B, c->D + 1, 2 * MS, c->num_sms);
    w.nsplit = (int)ns;
    w.off_g = o; o = up(o + ns * B * w.ldb * 8);
    w.ncb = (int)((MS + P_EPART_COLS - 1) / P_EPART_COLS);  // capacity: the narrowest pair tile
    w.off_epart = o; o = up(o + B * w.ncb * 8);
    w.total = o + 256;
    return w;
}

int predict(mlffpc_ctx* ctx, const double* Rq_desc, const double* Rq_d_desc, int64_t B, const double* v,
            const double* beta, double* F_out, double* E_out, void* workspace, cudaStream_t s) {
    const PredWs w = pred_layout(ctx, B);
    char* base = (char*)(((uintptr_t)workspace + 255) / 256 * 256);
    double* Bmat = (double*)(base + w.off_bmat);
    double* Cmat = (double*)(base + w.off_cmat);
    double* G = (double*)(base + w.off_g);
    double* Epart = (double*)(base + w.off_epart);
    const int64_t MS = ctx->M * ctx->S;
    const int D = ctx->D, N = ctx->N;
    const double q = sqrt(5.0) / ctx->sig, pref = 5.0 / (3.0 * ctx->sig * ctx->sig);
    const int64_t total = MS * (D + 1);
    mv_prepare_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(MS, ctx->S, D, N, w.ldb, ctx->Xp, ctx->R_d_desc,
                                                                     ctx->desc_perms, ctx->pair_a, ctx->pair_b, v, beta, Bmat);
    MLFFPC_LAUNCH_CHECK();
    int n_epart = 0;
    MLFFPC_TRY(launch_pairs(ctx, Rq_desc, B, Bmat, w.ldb, MS, D, q, pref, Cmat, E_out ? Epart : nullptr, &n_epart, s));
    MLFFPC_TRY(dgemm(false, B, D + 1, 2 * MS, 1.0, Cmat, 2 * MS, Bmat, w.ldb, 0.0, G, w.ldb, false, s, w.nsplit, B * w.ldb));
    int block = 32;
    while (block < 3 * N && block < 256) block <<= 1;
    mv_epilogue_kernel<<<(unsigned)B, block, 0, s>>>(N, D, 0, Rq_desc, Rq_d_desc, G, w.ldb, w.nsplit, B * w.ldb, nullptr,
                                                    F_out, 1.0, 0.0, E_out ? Epart : nullptr, n_epart, E_out);
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

int mlffpc_desc_from_r(mlffpc_ctx* ctx, const double* R, int64_t B, int N, double* R_desc, double* R_d_desc, void* stream) {
    MLFFPC_REQUIRE(ctx && R && R_desc && R_d_desc && B >= 0 && N >= 2, "desc_from_r: bad argument");
    const int D = N * (N - 1) / 2;
    const int64_t total = B * D;
    if (total == 0) return MLFFPC_OK;
    desc_from_r_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(R, B, N, D, R_desc, R_d_desc);
    MLFFPC_LAUNCH_CHECK();
    return MLFFPC_OK;
}

int mlffpc_d_desc_dot_vec(mlffpc_ctx* ctx, const double* v, double* out, void* stream) {
    MLFFPC_REQUIRE(ctx && ctx->M > 0 && v && out, "d_desc_dot_vec: geometry not set or NULL argument");
    MLFFPC_REQUIRE(ctx->R_d_desc, "d_desc_dot_vec: geometry was set without R_d_desc");
    const int64_t total = ctx->M * ctx->D;
    jv_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ctx->M, ctx->D, ctx->N, ctx->R_d_desc,
                                                                                  ctx->pair_a, ctx->pair_b, v, out);
    MLFFPC_LAUNCH_CHECK();
    return MLFFPC_OK;
}

int mlffpc_predict_workspace_bytes(mlffpc_ctx* ctx, int64_t B, int64_t* bytes) {
    MLFFPC_REQUIRE(ctx && bytes && B > 0 && ctx->M > 0, "predict_workspace_bytes: bad argument / geometry not set");
    *bytes = pred_layout(ctx, B).total;
    return MLFFPC_OK;
}

int mlffpc_predict(mlffpc_ctx* ctx, const double* Rq_desc, const double* Rq_d_desc, int64_t B, const double* v,
                   const double* beta, double* F_out, double* E_out, void* workspace, int64_t workspace_bytes,
                   void* stream) {
    MLFFPC_REQUIRE(ctx && ctx->M > 0, "predict: geometry not set");
    MLFFPC_REQUIRE(Rq_desc && Rq_d_desc && F_out && workspace && B > 0, "predict: bad argument");
    MLFFPC_REQUIRE((v != nullptr) != (beta != nullptr), "predict: give exactly one of v (alphas) and beta (R_d_desc_alpha)");
    MLFFPC_REQUIRE(beta || ctx->R_d_desc, "predict: coefficients v need the training Jacobians (geometry was set without R_d_desc)");
    MLFFPC_REQUIRE(workspace_bytes >= pred_layout(ctx, B).total, "predict: workspace too small");
    return predict(ctx, Rq_desc, Rq_d_desc, B, v, beta, F_out, E_out, workspace, (cudaStream_t)stream);
}

int mlffpc_matvec_free(mlffpc_ctx* ctx, const double* v, double* y_local, double alpha, double shift,
                       void* workspace, int64_t workspace_bytes, void* stream) {
    MLFFPC_REQUIRE(ctx && ctx->M > 0, "matvec_free: geometry not set");
    MLFFPC_REQUIRE(v && y_local && workspace, "matvec_free: NULL argument");
    MLFFPC_REQUIRE(ctx->R_d_desc, "matvec_free: geometry was set without R_d_desc");
    MLFFPC_REQUIRE(workspace_bytes >= matvec_free_ws_bytes(ctx), "matvec_free: workspace too small");
    return matvec_free(ctx, v, y_local, alpha, shift, workspace, (cudaStream_t)stream);
}

}  // extern "C"
