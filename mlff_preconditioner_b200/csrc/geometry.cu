// Kernel-entry generation from descriptors: geometry tables, -diag(K), explicit rows, column panels.
//
// Block formula (reference train.py:162-204):
//   Delta_p = x_i - x_j^(p),  rho^ = (sqrt5/sig)|Delta_p|,  e^ = 5/(3 sig^2) exp(-rho^)
//   a_p = e^ * 5/sig^2,  b_p = e^ * (1 + rho^)
//   K_ij = sum_p [ a_p u_ip u_jp^T - b_p J_i^T J_j^(p) ],  u_ip = J_i^T Delta_p,  u_jp = J_j^(p)T Delta_p
// with the sparse Jacobian J[d, b_d] = +g_d, J[d, a_d] = -g_d (utils/desc.py:444-462) never inflated.
#include "common.cuh"
#include "symlayout.cuh"

namespace mlffpc {

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

struct GeoWs {
    int64_t off_pair_a, off_pair_b, off_pinv, off_xp, total;
};
static GeoWs geo_layout(int64_t M, int N, int S) {
    const int64_t D = (int64_t)N * (N - 1) / 2;
    GeoWs w;
    int64_t o = 0;
    w.off_pair_a = o; o = align_up(o + D * 4, 256);
    w.off_pair_b = o; o = align_up(o + D * 4, 256);
    w.off_pinv = o;   o = align_up(o + (int64_t)S * N * 4, 256);
    w.off_xp = o;     o = align_up(o + (S > 1 ? M * S * D * 8 : 0), 256);
    w.total = o + 256;
    return w;
}

__global__ void geo_tables_kernel(int N, int S, int D, const int32_t* __restrict__ atom_perms,
                                  int32_t* pair_a, int32_t* pair_b, int32_t* pinv) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < D) {
        // invert d = a(a-1)/2 + b
        int a = (int)floor((1.0 + sqrt(1.0 + 8.0 * (double)t)) * 0.5);
        while (a * (a - 1) / 2 > t) --a;
        while ((a + 1) * a / 2 <= t) ++a;
        pair_a[t] = a;
        pair_b[t] = t - a * (a - 1) / 2;
    }
    if (t < S * N) {
        const int p = t / N, x = t % N;
        pinv[p * N + atom_perms[p * N + x]] = x;
    }
}

__global__ void geo_permute_kernel(int64_t total, int S, int D, const double* __restrict__ R_desc,
                                   const int32_t* __restrict__ desc_perms, double* __restrict__ Xp) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int d = (int)(t % D);
    const int64_t jp = t / D;
    const int p = (int)(jp % S);
    const int64_t j = jp / S;
    Xp[t] = R_desc[j * D + desc_perms[p * D + d]];
}

// ---- shared device pieces -------------------------------------------------------------

struct GeoView {
    int N, S, D, dim_i;
    double q, pref;  // sqrt5/sig, 5/(3 sig^2)
    const double* R_desc;
    const double* R_d_desc;
    const double* Xp;
    const int32_t* desc_perms;
    const int32_t* P;
    const int32_t* Pinv;
    const int32_t* pair_a;  // [D] larger atom of descriptor d
    const int32_t* pair_b;  // [D] smaller atom
};

static GeoView make_view(const mlffpc_ctx* c) {
    GeoView g;
    g.N = c->N; g.S = c->S; g.D = c->D; g.dim_i = c->dim_i;
    g.q = sqrt(5.0) / c->sig;
    g.pref = 5.0 / (3.0 * c->sig * c->sig);
    g.R_desc = c->R_desc; g.R_d_desc = c->R_d_desc; g.Xp = c->Xp;
    g.desc_perms = c->desc_perms; g.P = c->atom_perms; g.Pinv = c->atom_perms_inv;
    g.pair_a = c->pair_a; g.pair_b = c->pair_b;
    return g;
}

// a_p, b_p for the pair (i, j, p) -> sm_ab[2p], sm_ab[2p+1]; all threads participate
__device__ __forceinline__ void pair_coeffs(const GeoView& g, const double* xi, const double* xjp_base,
                                            double* sm_ab, double* sm_red) {
    for (int p = 0; p < g.S; ++p) {
        const double* xjp = xjp_base + (int64_t)p * g.D;
        double s2 = 0.0;
        for (int d = threadIdx.x; d < g.D; d += blockDim.x) {
            const double dl = xi[d] - xjp[d];
            s2 = fma(dl, dl, s2);
        }
        s2 = block_sum(s2, sm_red);
        if (threadIdx.x == 0) {
            const double rho = g.q * sqrt(s2);
            const double e = g.pref * exp(-rho);
            sm_ab[2 * p] = e * g.q * g.q;
            sm_ab[2 * p + 1] = e * (1.0 + rho);
        }
    }
    __syncthreads();
}

// u_ip[(A,c)] = sum_{B != A} sgn(A,B) g_i[pair(A,B), c] Delta_p[pair(A,B)]
__device__ __forceinline__ double u_i_entry(const GeoView& g, const double* gi, const double* xi,
                                            const double* xjp, int A, int c) {
    double acc = 0.0;
    for (int B = 0; B < g.N; ++B) {
        if (B == A) continue;
        const int d = pair_index(A, B);
        const double dl = xi[d] - xjp[d];
        const double t = gi[d * 3 + c] * dl;
        acc += (A < B) ? t : -t;
    }
    return acc;
}

// u_jp[(A2,c2)] = sum_d J_j^(p)[d, (A2,c2)] Delta_p[d]
__device__ __forceinline__ double u_j_entry(const GeoView& g, const double* gj, const double* xi,
                                            const double* xjp, const int32_t* P, const int32_t* Pinv,
                                            int A2, int c2) {
    const int App = Pinv[A2];
    double acc = 0.0;
    for (int B = 0; B < g.N; ++B) {
        if (B == App) continue;
        const int d = pair_index(App, B);
        const int PB = P[B];
        const int e = pair_index(A2, PB);
        const double dl = xi[d] - xjp[d];
        const double t = gj[e * 3 + c2] * dl;
        acc += (A2 < PB) ? t : -t;
    }
    return acc;
}

// (J_i^T J_j^(p))[(A,c),(A2,c2)]
__device__ __forceinline__ double jtj_entry(const GeoView& g, const double* gi, const double* gj,
                                            const int32_t* P, const int32_t* Pinv, int A, int c, int A2,
                                            int c2) {
    const int App = Pinv[A2];
    if (App != A) {
        const int d = pair_index(A, App);
        const int PA = P[A];
        const int e = pair_index(PA, A2);
        const double t = gi[d * 3 + c] * gj[e * 3 + c2];
        return ((A < App) == (A2 < PA)) ? t : -t;
    }
    double acc = 0.0;
    for (int B = 0; B < g.N; ++B) {
        if (B == A) continue;
        const int d = pair_index(A, B);
        const int PB = P[B];
        const int e = pair_index(A2, PB);
        const double t = gi[d * 3 + c] * gj[e * 3 + c2];
        acc += ((A < B) == (A2 < PB)) ? t : -t;
    }
    return acc;
}

// ---- explicit block (i, j): one CTA -----------------------------------------------------
// dynamic smem: ab[2S] | red[33] | u_i[S*dim_i] | u_j[S*dim_i]
template <bool DIAG_ONLY>
__global__ void assemble_block_kernel(GeoView g, int64_t pt0, int64_t j_pt0, int packed,
                                      double* __restrict__ out, int64_t ld) {
    extern __shared__ double sm[];
    double* sm_ab = sm;
    double* sm_red = sm_ab + 2 * g.S;
    double* sm_ui = sm_red + 40;
    double* sm_uj = sm_ui + g.S * g.dim_i;

    const int64_t il = blockIdx.x;       // local row point
    const int64_t i = pt0 + il;
    const int64_t jl = DIAG_ONLY ? il : (int64_t)blockIdx.y;
    const int64_t j = DIAG_ONLY ? i : j_pt0 + jl;
    // packed diagonal tile: band b of 256 rows keeps columns [0, 256 (b + 1)); skip point blocks right of that
    if (!DIAG_ONLY && packed && jl * g.dim_i >= st_band_pitch(((il + 1) * g.dim_i - 1) / ST_BAND_ROWS)) return;
    const double* xi = g.R_desc + i * g.D;
    const double* gi = g.R_d_desc + i * g.D * 3;
    const double* gj = g.R_d_desc + j * g.D * 3;
    const double* xjp_base = g.Xp + j * g.S * (int64_t)g.D;

    pair_coeffs(g, xi, xjp_base, sm_ab, sm_red);

    for (int t = threadIdx.x; t < g.S * g.dim_i; t += blockDim.x) {
        const int p = t / g.dim_i, r = t % g.dim_i;
        const double* xjp = xjp_base + (int64_t)p * g.D;
        sm_ui[t] = u_i_entry(g, gi, xi, xjp, r / 3, r % 3);
        sm_uj[t] = u_j_entry(g, gj, xi, xjp, g.P + p * g.N, g.Pinv + p * g.N, r / 3, r % 3);
    }
    __syncthreads();

    if (DIAG_ONLY) {
        for (int r = threadIdx.x; r < g.dim_i; r += blockDim.x) {
            double val = 0.0;
            for (int p = 0; p < g.S; ++p) {
                const double G = jtj_entry(g, gi, gj, g.P + p * g.N, g.Pinv + p * g.N, r / 3, r % 3, r / 3, r % 3);
                val += sm_ab[2 * p] * sm_ui[p * g.dim_i + r] * sm_uj[p * g.dim_i + r] - sm_ab[2 * p + 1] * G;
            }
            out[il * g.dim_i + r] = -val;
        }
    } else {
        const int nent = g.dim_i * g.dim_i;
        double* blk = out + (il * g.dim_i) * ld + jl * g.dim_i;
        for (int t = threadIdx.x; t < nent; t += blockDim.x) {
            const int r = t / g.dim_i, r2 = t % g.dim_i;
            double* dst = blk + (int64_t)r * ld + r2;
            if (packed) {
                const int64_t row = il * g.dim_i + r, col = jl * g.dim_i + r2, b = row / ST_BAND_ROWS;
                if (col >= st_band_pitch(b)) continue;
                dst = out + st_band_off(b) + (row - b * ST_BAND_ROWS) * st_band_pitch(b) + col;
            }
            double val = 0.0;
            for (int p = 0; p < g.S; ++p) {
                const double G = jtj_entry(g, gi, gj, g.P + p * g.N, g.Pinv + p * g.N, r / 3, r % 3, r2 / 3, r2 % 3);
                val += sm_ab[2 * p] * sm_ui[p * g.dim_i + r] * sm_uj[p * g.dim_i + r2] - sm_ab[2 * p + 1] * G;
            }
            *dst = val;
        }
    }
}


// ---- explicit rows: one CTA = one row point i x up to JT column points ------------------------------------
// Phase 1 builds, for every (column point jj, permutation p) =: q of the CTA, in shared memory
//   dl[q][d]        = Delta_p = x_i - x_j^(p)                       -> a_p, b_p  (ab[q])
//   ui[q][(A,c)]    = (J_i^T Delta)[(A,c)],   uj[q][(A2,c2)] = (J_j^(p)T Delta)[(A2,c2)]
//   prod[q][d][c,c2] = sigma_p(d) g_i[d,c] g_j[e_p(d),c2]   with e_p(d) = pair(P[a_d], P[b_d]), sigma = +-1:
//                      the 3x3 block of J_i^T J_j^(p) for the atom pair d = (A, Pinv[A2])
//   gd[q][A2][c,c2]  = -sum_{B != A} prod[q][pair(A,B)][c,c2],  A = Pinv[A2]: the diagonal atom blocks
//                      (translation invariance: every row of J sums to zero)
// (all threads busy: JT * S * D * 9 products).  Phase 2 gives every thread one OUTPUT COLUMN (point jj, atom A2,
// component c2) and lets it walk down the 3N rows, so a warp stores 32 consecutive doubles of one row of K --
// full-sector coalesced writes, the HBM-write roofline of assembly; per entry it needs two shared loads and two
// FMAs, no data-dependent branch or loop.
// dynamic smem (doubles): xi[D] | gi[3D] | gj[JT][3D] | dl[JT S][D] | ui[JT S][3N] | uj[JT S][3N] | ab[JT S][2]
//                         | gd[JT S][N][9] | prod[JT S][D][9]   then ints: P[S][N] | Pinv[S][N]
constexpr int ASM_THREADS = 256;

template <bool S1>
__global__ void __launch_bounds__(ASM_THREADS)
assemble_rows_kernel(GeoView g, int64_t i_pt0, int64_t j_pt0, int64_t n_jpts, int JT, int packed,
                     double* __restrict__ out, int64_t ld) {
    extern __shared__ double asm_sm[];
    const int D = g.D, di = g.dim_i, S = S1 ? 1 : g.S, N = g.N;
    double* xi = asm_sm;
    double* gi = xi + D;
    double* gj = gi + 3 * D;
    double* dl = gj + (size_t)JT * 3 * D;
    double* ui = dl + (size_t)JT * S * D;
    double* uj = ui + (size_t)JT * S * di;
    double* ab = uj + (size_t)JT * S * di;
    double* gd = ab + (size_t)JT * S * 2;
    double* prod = gd + (size_t)JT * S * N * 9;
    int* P = (int*)(prod + (size_t)JT * S * D * 9);
    int* Pinv = P + S * N;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t il = blockIdx.x, i = i_pt0 + il;
    const int64_t jb = (int64_t)blockIdx.y * JT;
    const int nj = (int)((n_jpts - jb < JT) ? (n_jpts - jb) : JT);
    const int nq = nj * S;
    // packed diagonal tile: band b of 256 rows keeps columns [0, 256 (b + 1)); skip CTAs entirely right of that
    if (packed && jb * di >= st_band_pitch(((il + 1) * di - 1) / ST_BAND_ROWS)) return;

    for (int t = tid; t < D; t += ASM_THREADS) xi[t] = g.R_desc[i * D + t];
    for (int t = tid; t < 3 * D; t += ASM_THREADS) gi[t] = g.R_d_desc[i * D * 3 + t];
    for (int t = tid; t < S * N; t += ASM_THREADS) { P[t] = g.P[t]; Pinv[t] = g.Pinv[t]; }
    for (int jj = 0; jj < nj; ++jj) {
        const double* src = g.R_d_desc + (j_pt0 + jb + jj) * D * 3;
        for (int e = tid; e < 3 * D; e += ASM_THREADS) gj[(size_t)jj * 3 * D + e] = src[e];
    }
    __syncthreads();
    for (int q = 0; q < nq; ++q) {
        const double* xj = g.Xp + ((j_pt0 + jb + q / S) * S + q % S) * (int64_t)D;
        for (int d = tid; d < D; d += ASM_THREADS) dl[(size_t)q * D + d] = xi[d] - xj[d];
    }
    // prod[q][d][c,c2]: one thread per (q, d), 9 products each
    for (int t = tid; t < nq * D; t += ASM_THREADS) {
        const int d = t % D, q = t / D, p = q % S, jj = q / S;
        const int a = g.pair_a[d], b = g.pair_b[d];  // a > b
        const int Pa = P[p * N + a], Pb = P[p * N + b];
        const int e = pair_index(Pa, Pb);
        const double sg = (Pa < Pb) ? 1.0 : -1.0;
        const double* gje = gj + (size_t)jj * 3 * D + e * 3;
        const double g0 = sg * gi[d * 3], g1 = sg * gi[d * 3 + 1], g2 = sg * gi[d * 3 + 2];
        double* o = prod + (size_t)t * 9;
        o[0] = g0 * gje[0]; o[1] = g0 * gje[1]; o[2] = g0 * gje[2];
        o[3] = g1 * gje[0]; o[4] = g1 * gje[1]; o[5] = g1 * gje[2];
        o[6] = g2 * gje[0]; o[7] = g2 * gje[1]; o[8] = g2 * gje[2];
    }
    __syncthreads();
    // a_p, b_p: one warp per q
    for (int q = warp; q < nq; q += ASM_THREADS / 32) {
        double s2 = 0.0;
        for (int d = lane; d < D; d += 32) { const double v = dl[(size_t)q * D + d]; s2 = fma(v, v, s2); }
        s2 = warp_sum(s2);
        if (lane == 0) {
            const double rho = g.q * sqrt(s2);
            const double e = g.pref * exp(-rho);
            ab[2 * q] = e * g.q * g.q;
            ab[2 * q + 1] = e * (1.0 + rho);
        }
    }
    // u_i[(A,c)] = sum_{B != A} sgn(A,B) g_i[pair(A,B), c] Delta[pair(A,B)]
    // u_j[(A2,c2)] = sum_{B != App} sgn(A2, P[B]) g_j[pair(A2, P[B]), c2] Delta[pair(App, B)],  App = Pinv[A2]
    for (int t = tid; t < nq * di; t += ASM_THREADS) {
        const int r = t % di, q = t / di, p = q % S, jj = q / S;
        const int A = r / 3, c = r % 3;
        const double* dq = dl + (size_t)q * D;
        const double* gjj = gj + (size_t)jj * 3 * D;
        const int* Pp = P + p * N;
        const int App = Pinv[p * N + A];
        double a1 = 0.0, a2 = 0.0;
        for (int B = 0; B < N; ++B) {
            if (B != A) {
                const int d = pair_index(A, B);
                const double t1 = gi[d * 3 + c] * dq[d];
                a1 += (A < B) ? t1 : -t1;
            }
            if (B != App) {
                const int d = pair_index(App, B);
                const int PB = Pp[B];
                const double t2 = gjj[pair_index(A, PB) * 3 + c] * dq[d];
                a2 += (A < PB) ? t2 : -t2;
            }
        }
        ui[t] = a1;
        uj[t] = a2;
    }
    // gd[q][A2][c,c2] = -sum_{B != A} prod[q][pair(A,B)][c,c2],  A = Pinv[A2]
    for (int t = tid; t < nq * N * 9; t += ASM_THREADS) {
        const int cc = t % 9, A2 = (t / 9) % N, q = t / (9 * N), p = q % S;
        const int A = Pinv[p * N + A2];
        const double* pq = prod + (size_t)q * D * 9 + cc;
        double acc = 0.0;
        for (int B = 0; B < N; ++B)
            if (B != A) acc -= pq[pair_index(A, B) * 9];
        gd[t] = acc;
    }
    __syncthreads();

    // phase 2: thread <-> output column
    for (int col = tid; col < nj * di; col += ASM_THREADS) {
        const int jj = col / di, r2 = col % di, A2 = r2 / 3, c2 = r2 % 3;
        const int64_t gcol = (jb + jj) * di + r2;  // column inside the tile
        // S1: everything that does not depend on the row lives in registers
        const int q1 = jj;
        const int App1 = Pinv[A2];
        const double auj1 = ab[2 * q1] * uj[(size_t)q1 * di + r2], b1 = ab[2 * q1 + 1];
        const double* prod1 = prod + (size_t)q1 * D * 9 + c2;
        const double* gd1 = gd + ((size_t)q1 * N + A2) * 9 + c2;
        const double* ui1 = ui + (size_t)q1 * di;
        for (int A = 0; A < N; ++A) {
            const double* pr1 = (A != App1) ? (prod1 + pair_index(A, App1) * 9) : gd1;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int r = 3 * A + c;
                double val;
                if (S1) {
                    val = fma(auj1, ui1[r], -b1 * pr1[c * 3]);
                } else {
                    val = 0.0;
                    for (int p = 0; p < S; ++p) {
                        const int q = jj * S + p;
                        const int App = Pinv[p * N + A2];
                        const double G = (App != A) ? prod[((size_t)q * D + pair_index(A, App)) * 9 + c * 3 + c2]
                                                    : gd[((size_t)q * N + A2) * 9 + c * 3 + c2];
                        val += ab[2 * q] * ui[(size_t)q * di + r] * uj[(size_t)q * di + r2] - ab[2 * q + 1] * G;
                    }
                }
                const int64_t row = il * di + r;
                if (packed) {
                    const int64_t b = row / ST_BAND_ROWS;
                    if (gcol < st_band_pitch(b)) out[st_band_off(b) + (row - b * ST_BAND_ROWS) * st_band_pitch(b) + gcol] = val;
                } else {
                    out[row * ld + gcol] = val;
                }
            }
        }
    }
}

static size_t assemble_rows_smem(const GeoView& g, int JT) {
    const size_t dbl = (size_t)g.D + 3 * g.D + (size_t)JT * 3 * g.D + (size_t)JT * g.S * g.D +
                       2 * (size_t)JT * g.S * g.dim_i + (size_t)JT * g.S * 2 + (size_t)JT * g.S * g.N * 9 +
                       (size_t)JT * g.S * g.D * 9;
    return dbl * sizeof(double) + 2 * (size_t)g.S * g.N * sizeof(int) + 16;
}

// ---- one column restricted to the local rows: CTA per (local point i, column c) -------------
// out[c, il*dim_i + r] = scale * K[(i, r), cols[c]]
// dynamic smem: ab[2S] | red[40] | uj[S] | ui[S*dim_i]
__global__ void column_kernel(GeoView g, int64_t pt0, const int64_t* __restrict__ cols,
                              double* __restrict__ out, int64_t ld, double scale) {
    extern __shared__ double sm[];
    double* sm_ab = sm;
    double* sm_red = sm_ab + 2 * g.S;
    double* sm_uj = sm_red + 40;

    const int64_t il = blockIdx.x;
    const int64_t i = pt0 + il;
    const int64_t col = cols[blockIdx.y];
    const int64_t j = col / g.dim_i;
    const int r2 = (int)(col % g.dim_i);
    const int A2 = r2 / 3, c2 = r2 % 3;
    const double* xi = g.R_desc + i * g.D;
    const double* gi = g.R_d_desc + i * g.D * 3;
    const double* gj = g.R_d_desc + j * g.D * 3;
    const double* xjp_base = g.Xp + j * g.S * (int64_t)g.D;

    pair_coeffs(g, xi, xjp_base, sm_ab, sm_red);

    // u_jp[(A2,c2)] : N-1 terms, block-reduced
    for (int p = 0; p < g.S; ++p) {
        const double* xjp = xjp_base + (int64_t)p * g.D;
        const int32_t* P = g.P + p * g.N;
        const int App = g.Pinv[p * g.N + A2];
        double acc = 0.0;
        for (int B = threadIdx.x; B < g.N; B += blockDim.x) {
            if (B == App) continue;
            const int d = pair_index(App, B);
            const int PB = P[B];
            const int e = pair_index(A2, PB);
            const double t = gj[e * 3 + c2] * (xi[d] - xjp[d]);
            acc += (A2 < PB) ? t : -t;
        }
        acc = block_sum(acc, sm_red);
        if (threadIdx.x == 0) sm_uj[p] = acc;
    }
    __syncthreads();

    double* orow = out + (int64_t)blockIdx.y * ld + il * g.dim_i;
    for (int r = threadIdx.x; r < g.dim_i; r += blockDim.x) {
        const int A = r / 3, c = r % 3;
        double val = 0.0;
        for (int p = 0; p < g.S; ++p) {
            const double* xjp = xjp_base + (int64_t)p * g.D;
            const double ui = u_i_entry(g, gi, xi, xjp, A, c);
            const double G = jtj_entry(g, gi, gj, g.P + p * g.N, g.Pinv + p * g.N, A, c, A2, c2);
            val += sm_ab[2 * p] * ui * sm_uj[p] - sm_ab[2 * p + 1] * G;
        }
        orow[r] = scale * val;
    }
}

static int pick_block(int work) {
    int b = 32;
    while (b < work && b < 256) b <<= 1;
    return b;
}

}  // namespace mlffpc

using namespace mlffpc;

extern "C" {

int mlffpc_geometry_workspace_bytes(int64_t M, int N, int S, int64_t* bytes) {
    MLFFPC_REQUIRE(bytes && M > 0 && N >= 2 && S >= 1, "geometry_workspace_bytes: bad argument");
    *bytes = geo_layout(M, N, S).total;
    return MLFFPC_OK;
}

int mlffpc_set_geometry(mlffpc_ctx* ctx, int64_t M, int N, int S, const double* R_desc,
                        const double* R_d_desc, const int32_t* desc_perms, const int32_t* atom_perms,
                        double sig, int64_t pt0, int64_t pt1, void* workspace, int64_t workspace_bytes,
                        void* stream) {
    // R_d_desc may be NULL for a prediction-only context (mlffpc_predict with beta); everything that needs the
    // training Jacobians checks for it
    MLFFPC_REQUIRE(ctx && R_desc && desc_perms && atom_perms && workspace, "set_geometry: NULL argument");
    MLFFPC_REQUIRE(M > 0 && N >= 2 && S >= 1 && sig > 0, "set_geometry: bad sizes (M=%lld N=%d S=%d sig=%g)",
                   (long long)M, N, S, sig);
    MLFFPC_REQUIRE(0 <= pt0 && pt0 < pt1 && pt1 <= M, "set_geometry: bad shard [%lld, %lld) of %lld",
                   (long long)pt0, (long long)pt1, (long long)M);
    const GeoWs w = geo_layout(M, N, S);
    MLFFPC_REQUIRE(workspace_bytes >= w.total, "set_geometry: workspace too small (%lld < %lld)",
                   (long long)workspace_bytes, (long long)w.total);
    MLFFPC_REQUIRE((int64_t)3 * N * M < ((int64_t)1 << 40), "set_geometry: system too large");
    cudaStream_t s = (cudaStream_t)stream;
    char* base = (char*)(((uintptr_t)workspace + 255) / 256 * 256);
    ctx->M = M; ctx->N = N; ctx->S = S; ctx->D = N * (N - 1) / 2; ctx->dim_i = 3 * N;
    ctx->n = (int64_t)3 * N * M;
    ctx->sig = sig;
    ctx->R_desc = R_desc; ctx->R_d_desc = R_d_desc;
    ctx->desc_perms = desc_perms; ctx->atom_perms = atom_perms;
    ctx->pair_a = (int32_t*)(base + w.off_pair_a);
    ctx->pair_b = (int32_t*)(base + w.off_pair_b);
    ctx->atom_perms_inv = (int32_t*)(base + w.off_pinv);
    ctx->pt0 = pt0; ctx->pt1 = pt1;
    const int D = ctx->D;
    const int tt = (D > S * N ? D : S * N);
    geo_tables_kernel<<<(tt + 255) / 256, 256, 0, s>>>(N, S, D, atom_perms, ctx->pair_a, ctx->pair_b,
                                                       ctx->atom_perms_inv);
    MLFFPC_LAUNCH_CHECK();
    if (S > 1) {
        ctx->Xp = (double*)(base + w.off_xp);
        const int64_t total = M * S * D;
        geo_permute_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(total, S, D, R_desc, desc_perms, ctx->Xp);
        MLFFPC_LAUNCH_CHECK();
    } else {
        ctx->Xp = const_cast<double*>(R_desc);
    }
    return MLFFPC_OK;
}

int mlffpc_kernel_diag(mlffpc_ctx* ctx, double* out, void* stream) {
    MLFFPC_REQUIRE(ctx && out && ctx->M > 0, "kernel_diag: geometry not set or NULL output");
    MLFFPC_REQUIRE(ctx->R_d_desc, "kernel_diag: geometry was set without R_d_desc");
    GeoView g = make_view(ctx);
    const size_t smem = (size_t)(2 * g.S + 40 + 2 * g.S * g.dim_i) * sizeof(double);
    MLFFPC_REQUIRE(smem <= 200 * 1024, "kernel_diag: S*3N = %d too large for shared memory", g.S * g.dim_i);
    MLFFPC_CUDA(cudaFuncSetAttribute(assemble_block_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int block = pick_block(g.D < g.dim_i ? g.dim_i : g.D);
    assemble_block_kernel<true><<<dim3((unsigned)(ctx->pt1 - ctx->pt0)), block, smem, (cudaStream_t)stream>>>(
        g, ctx->pt0, 0, 0, out, 0);
    MLFFPC_LAUNCH_CHECK();
    return MLFFPC_OK;
}

int mlffpc_kernel_assemble(mlffpc_ctx* ctx, double* K_out, int64_t ld, void* stream) {
    MLFFPC_REQUIRE(ctx && K_out && ctx->M > 0, "kernel_assemble: geometry not set or NULL output");
    MLFFPC_REQUIRE(ld >= ctx->n, "kernel_assemble: ld %lld < n %lld", (long long)ld, (long long)ctx->n);
    ProfWindow pw = prof_window("assemble");
    pw.step(pw.first);
    const int st = assemble_tile(ctx, ctx->pt0, ctx->pt1, 0, ctx->M, K_out, ld, 0, (cudaStream_t)stream);
    pw.end();
    return st;
}

int mlffpc_kernel_columns_workspace_bytes(mlffpc_ctx* ctx, int64_t b, int64_t* bytes) {
    MLFFPC_REQUIRE(ctx && bytes && b >= 0, "kernel_columns_workspace_bytes: bad argument");
    *bytes = 256;  // the column kernel keeps everything in shared memory
    return MLFFPC_OK;
}

int mlffpc_kernel_columns(mlffpc_ctx* ctx, const int64_t* cols, int64_t b, double* out, int64_t ld,
                          double scale, void* workspace, int64_t workspace_bytes, void* stream) {
    (void)workspace; (void)workspace_bytes;
    MLFFPC_REQUIRE(ctx && ctx->M > 0, "kernel_columns: geometry not set");
    if (b == 0) return MLFFPC_OK;
    MLFFPC_REQUIRE(cols && out && b > 0, "kernel_columns: NULL argument");
    MLFFPC_REQUIRE(ctx->R_d_desc, "kernel_columns: geometry was set without R_d_desc");
    MLFFPC_REQUIRE(ld >= ctx->n_local(), "kernel_columns: ld %lld < n_local %lld", (long long)ld, (long long)ctx->n_local());
    GeoView g = make_view(ctx);
    const size_t smem = (size_t)(2 * g.S + 40 + g.S) * sizeof(double);
    MLFFPC_REQUIRE(smem <= 200 * 1024, "kernel_columns: S too large for shared memory");
    MLFFPC_CUDA(cudaFuncSetAttribute(column_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int block = pick_block(g.D < g.dim_i ? g.dim_i : (g.D < 128 ? g.D : 128));
    const int64_t ml = ctx->pt1 - ctx->pt0;
    for (int64_t c0 = 0; c0 < b; c0 += 65535) {
        const int64_t nb = (b - c0 < 65535) ? (b - c0) : 65535;
        column_kernel<<<dim3((unsigned)ml, (unsigned)nb), block, smem, (cudaStream_t)stream>>>(
            g, ctx->pt0, cols + c0, out + c0 * ld, ld, scale);
        MLFFPC_LAUNCH_CHECK();
    }
    return MLFFPC_OK;
}

}  // extern "C"

namespace mlffpc {
int assemble_tile(mlffpc_ctx* ctx, int64_t i_pt0, int64_t i_pt1, int64_t j_pt0, int64_t j_pt1, double* out,
                  int64_t ld, int packed, cudaStream_t s) {
    MLFFPC_REQUIRE(0 <= i_pt0 && i_pt0 < i_pt1 && i_pt1 <= ctx->M && 0 <= j_pt0 && j_pt0 < j_pt1 && j_pt1 <= ctx->M,
                   "assemble_tile: bad point ranges");
    MLFFPC_REQUIRE(packed ? (i_pt0 == j_pt0 && i_pt1 == j_pt1) : (ld >= (j_pt1 - j_pt0) * ctx->dim_i),
                   "assemble_tile: ld too small / packed layout needs a square diagonal tile");
    MLFFPC_REQUIRE(ctx->R_d_desc, "assemble_tile: geometry was set without R_d_desc");
    GeoView g = make_view(ctx);
    const int64_t ni = i_pt1 - i_pt0, nj = j_pt1 - j_pt0;
    // row-walking kernel: JT column points per CTA so that one CTA covers ~256 output columns
    int JT = ASM_THREADS / g.dim_i;
    if (JT > 8) JT = 8;
    if (JT < 1) JT = 1;
    while (JT > 1 && assemble_rows_smem(g, JT) > 72 * 1024) --JT;  // keep >= 3 CTAs per SM when possible
    const size_t smem_rows = assemble_rows_smem(g, JT);
    const int64_t gy = (nj + JT - 1) / JT;
    if (!ctx->assemble_legacy && smem_rows <= 200 * 1024 && gy <= 65535) {
        if (g.S == 1) {
            MLFFPC_CUDA(cudaFuncSetAttribute(assemble_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows));
            assemble_rows_kernel<true><<<dim3((unsigned)ni, (unsigned)gy), ASM_THREADS, smem_rows, s>>>(g, i_pt0, j_pt0, nj, JT,
                                                                                                    packed, out, ld);
        } else {
            MLFFPC_CUDA(cudaFuncSetAttribute(assemble_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows));
            assemble_rows_kernel<false><<<dim3((unsigned)ni, (unsigned)gy), ASM_THREADS, smem_rows, s>>>(g, i_pt0, j_pt0, nj, JT,
                                                                                                     packed, out, ld);
        }
        MLFFPC_LAUNCH_CHECK();
        return MLFFPC_OK;
    }
    // fallback (very large molecules: the per-point Jacobian does not fit in shared memory): one CTA per block
    MLFFPC_REQUIRE(!packed || nj <= 65535, "assemble_tile: packed tiles are limited to 65535 points");
    const size_t smem = (size_t)(2 * g.S + 40 + 2 * g.S * g.dim_i) * sizeof(double);
    MLFFPC_REQUIRE(smem <= 200 * 1024, "kernel_assemble: S*3N = %d too large for shared memory", g.S * g.dim_i);
    MLFFPC_CUDA(cudaFuncSetAttribute(assemble_block_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int block = pick_block(g.dim_i * g.dim_i);
    for (int64_t j0 = j_pt0; j0 < j_pt1; j0 += 65535) {  // grid.y limit
        const int64_t njc = (j_pt1 - j0 < 65535) ? (j_pt1 - j0) : 65535;
        assemble_block_kernel<false><<<dim3((unsigned)ni, (unsigned)njc), block, smem, s>>>(
            g, i_pt0, j0, packed, packed ? out : out + (j0 - j_pt0) * g.dim_i, ld);
        MLFFPC_LAUNCH_CHECK();
    }
    return MLFFPC_OK;
}

int launch_columns_device_col(mlffpc_ctx* ctx, const int64_t* col_dev, double* out, double scale,
                              cudaStream_t s) {
    return mlffpc_kernel_columns(ctx, col_dev, 1, out, ctx->n_local(), scale, nullptr, 0, (void*)s);
}
}  // namespace mlffpc
