// Persistent, TMA-fed pass over one tile of the symmetric operator (see symop.cu for the tile plan).
//
// One CTA per SM walks a contiguous range of 32-row x 256-column units (64 KB each) in strip-major order.
// A producer lane issues one cp.async.bulk.tensor.2d per unit into a 3-stage shared-memory ring (mbarrier
// full/empty pairs); four consumer warps turn every landed unit into 32 row partial sums (registers, carried
// across the units of a strip) and 256 column partial sums (one coalesced store per unit).  Nothing but the
// TMA touches K, so ~128 KB per SM are in flight without spending registers on it; ragged edges are
// zero-filled by the tensor map instead of being clamped in the inner loop.  Deterministic: every partial
// has exactly one writer and the sums are added in a fixed order by symv_tma_reduce_kernel.
//
// Diagonal tiles come in two layouts:
//   square : row-major [nr, ld], only columns < r0 + 32 of strip r0 are read (mlffpc_symv on a caller's K)
//   packed : bands of 256 rows; band b stores columns [0, 256 (b + 1)) with that pitch, one tensor map per
//            band (symop storage: half the footprint, which also keeps the stream inside the TLB reach)
// In both, strip s owns units j = 0 .. s/8: unit j covers columns [256 j, 256 j + 256); the last unit holds
// the remaining two-sided columns (r0 % 256 of them) followed by the strip's 32 x 32 diagonal block.
#include <cuda.h>

#include <mutex>
#include <vector>

#include "common.cuh"
#include "symlayout.cuh"

namespace mlffpc {

constexpr int STR_COLS = 64, STR_SPLIT = 4;  // reduce kernels: columns per block, strip split
constexpr int ST_STAGES = 3;
constexpr int ST_CONSUMERS = 128;  // thread t owns columns (2t, 2t+1) of a unit
constexpr int ST_THREADS = ST_CONSUMERS + 32;
constexpr int ST_STAGE_BYTES = ST_ROWS * ST_COLS * 8;
constexpr size_t ST_SMEM = (size_t)ST_STAGES * ST_STAGE_BYTES + 1024 /* alignment slack */ + 4096 /* small arrays */;

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c_inner,
                                            int c_outer, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(ST_CONSUMERS) : "memory"); }

struct SymTmaArgs {
    int64_t nr, nc, nstrips, units_total, units_per_cta, ld_ws;
    int diag, packed;
    const CUtensorMap* tmaps;  // global memory: one map (square / rectangular) or one per 256-row band (packed)
    const double* xr;
    const double* xc;
    double* ws;       // [nstrips, ld_ws] per-unit column sums
    double* rowpart;  // [nstrips, 2, 32] per-CTA row sums of a strip (slot 0: the CTA that starts the strip)
};

__global__ void __launch_bounds__(ST_THREADS, 1) symv_tma_kernel(const SymTmaArgs a) {
    extern __shared__ unsigned char st_smem_raw[];
    unsigned char* sm = (unsigned char*)(((uintptr_t)st_smem_raw + 1023) & ~(uintptr_t)1023);
    double* tiles = (double*)sm;                                   // [STAGES][32][256]
    unsigned char* small = sm + (size_t)ST_STAGES * ST_STAGE_BYTES;
    uint64_t* full_bar = (uint64_t*)small;                         // [STAGES]
    uint64_t* empty_bar = full_bar + ST_STAGES;                    // [STAGES]
    double* xs = (double*)(small + 64);                            // [32]
    double* dsum = xs + ST_ROWS;                                   // [32]
    double* red = dsum + ST_ROWS;                                  // [32][4]

    const int tid = threadIdx.x;
    const int64_t u0 = (int64_t)blockIdx.x * a.units_per_cta;
    int64_t u1 = u0 + a.units_per_cta;
    if (u1 > a.units_total) u1 = a.units_total;
    if (u0 >= u1) return;

    if (tid == 0) {
        for (int i = 0; i < ST_STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], ST_CONSUMERS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // first unit of this CTA: the strip s with units_before(s) <= u0 < units_before(s + 1)
    int64_t lo = 0, hi = a.nstrips;
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (st_units_before(mid, a.diag, a.nc) <= u0) lo = mid; else hi = mid;
    }
    int64_t s = lo;
    int64_t j = u0 - st_units_before(s, a.diag, a.nc);
    int64_t nj = st_units_in_strip(s, a.diag, a.nc);

    if (tid >= ST_CONSUMERS) {
        // ------------------------------------------------------------------ producer warp
        if (tid == ST_CONSUMERS) {
            uint64_t policy;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t u = u0; u < u1; ++u) {
                const CUtensorMap* map = a.packed ? (a.tmaps + s / ST_BAND_STRIPS) : a.tmaps;
                const int row = a.packed ? (int)((s % ST_BAND_STRIPS) * ST_ROWS) : (int)(s * ST_ROWS);
                mbar_wait(&empty_bar[stage], phase ^ 1);
                mbar_expect_tx(&full_bar[stage], ST_STAGE_BYTES);
                tma_load_2d(tiles + (size_t)stage * (ST_ROWS * ST_COLS), map, &full_bar[stage], (int)(j * ST_COLS), row,
                            policy);
                if (++stage == ST_STAGES) { stage = 0; phase ^= 1; }
                if (++j == nj) { ++s; j = 0; nj = st_units_in_strip(s, a.diag, a.nc); }
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    const int lane = tid & 31, warp = tid >> 5;
    double acc[ST_ROWS];
    int stage = 0;
    uint32_t phase = 0;
    bool fresh = true;  // the strip's x slice / accumulators need (re)loading
    for (int64_t u = u0; u < u1; ++u) {
        const int64_t r0 = s * ST_ROWS;
        const int rows = (int)((a.nr - r0 < ST_ROWS) ? (a.nr - r0) : ST_ROWS);
        if (fresh) {
            if (tid < ST_ROWS) {
                xs[tid] = (tid < rows) ? a.xr[r0 + tid] : 0.0;
                dsum[tid] = 0.0;
            }
#pragma unroll
            for (int i = 0; i < ST_ROWS; ++i) acc[i] = 0.0;
            consumer_bar();
            fresh = false;
        }
        const double* tile = tiles + (size_t)stage * (ST_ROWS * ST_COLS);
        const int64_t col0 = j * ST_COLS;
        const int64_t n2 = a.diag ? r0 : a.nc;  // two-sided columns of this strip: [0, n2)
        const int w = (int)((n2 - col0 < ST_COLS) ? (n2 - col0) : ST_COLS);
        const bool okx = 2 * tid < w, oky = 2 * tid + 1 < w;
        double2 xv = make_double2(0.0, 0.0);
        if (okx) xv.x = __ldg(a.xc + col0 + 2 * tid);
        if (oky) xv.y = __ldg(a.xc + col0 + 2 * tid + 1);
        mbar_wait(&full_bar[stage], phase);
        if (okx) {
            double2 cacc = make_double2(0.0, 0.0);
            const double2* tp = reinterpret_cast<const double2*>(tile) + tid;
#pragma unroll
            for (int i = 0; i < ST_ROWS; ++i) {
                double2 kv = tp[i * (ST_COLS / 2)];
                if (!oky) kv.y = 0.0;  // padding column of an odd-width tile: never used
                const double xrow = xs[i];
                acc[i] = fma(kv.y, xv.y, fma(kv.x, xv.x, acc[i]));
                cacc.x = fma(kv.x, xrow, cacc.x);
                cacc.y = fma(kv.y, xrow, cacc.y);
            }
            *reinterpret_cast<double2*>(a.ws + s * a.ld_ws + col0 + 2 * tid) = cacc;
        }
        if (a.diag && j == nj - 1) {
            // the strip's diagonal block sits right after the two-sided columns of its last unit: one-sided,
            // 4 threads per row
            const int r = tid >> 2, cp = tid & 3;
            double v = 0.0;
            if (r < rows)
                for (int c = cp; c < rows; c += 4) v = fma(tile[r * ST_COLS + w + c], xs[c], v);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            if (cp == 0) dsum[r] = v;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == ST_STAGES) { stage = 0; phase ^= 1; }

        if (j + 1 == nj || u + 1 == u1) {
            // flush this CTA's share of the strip's row sums
#pragma unroll
            for (int i = 0; i < ST_ROWS; ++i) {
                const double v = warp_sum(acc[i]);
                if (lane == 0) red[i * 4 + warp] = v;
            }
            consumer_bar();
            if (tid < ST_ROWS) {
                const int64_t first_unit = st_units_before(s, a.diag, a.nc);
                const int slot = (first_unit / a.units_per_cta == (int64_t)blockIdx.x) ? 0 : 1;
                const double t = dsum[tid] + ((red[tid * 4 + 0] + red[tid * 4 + 1]) + (red[tid * 4 + 2] + red[tid * 4 + 3]));
                a.rowpart[(s * 2 + slot) * ST_ROWS + tid] = t;
            }
            consumer_bar();
            fresh = true;
        }
        if (++j == nj) { ++s; j = 0; nj = st_units_in_strip(s, a.diag, a.nc); }
    }
}

// ---- all tiles of a rank in ONE persistent launch -----------------------------------------------------------
// Sharded, a rank owns the diagonal tile of its row block plus (world - 1) / 2 (+ half a) off-diagonal tiles.  Launched
// one by one they cost three launches and one pipeline fill / drain each -- at 8 ranks 15 launches for ~1 ms of
// streaming.  Here the unit lists of the tiles are concatenated: CTA b walks global units [b upc, (b + 1) upc) and
// crosses tile boundaries like it crosses strip boundaries; one reduce kernel then writes every output entry once
// (diagonal-tile columns also collect the row sums of the off-diagonal tiles, in tile order: same bits as the
// tile-by-tile passes, which added them in that order).
constexpr int SYM_MAX_TILES = 8;
struct SymTileDev {
    int64_t nr, nc, nstrips, unit_base, ld_ws, row_off;  // row_off: first row of this tile relative to the diagonal tile
    int diag, packed;
    const CUtensorMap* tmaps;
    const double* xr;
    const double* xc;
    double* ws;
    double* rowpart;
    double* out_c;
    int64_t blk_base, ncb;  // reduce kernel: first block and number of 64-column blocks of this tile
};
struct SymMultiArgs {
    int ntiles;
    int64_t units_total, units_per_cta;
    SymTileDev t[SYM_MAX_TILES];
};

__global__ void __launch_bounds__(ST_THREADS, 1) symv_tma_multi_kernel(const __grid_constant__ SymMultiArgs a) {
    extern __shared__ unsigned char st_smem_raw[];
    unsigned char* sm = (unsigned char*)(((uintptr_t)st_smem_raw + 1023) & ~(uintptr_t)1023);
    double* tiles = (double*)sm;                                   // [STAGES][32][256]
    unsigned char* small = sm + (size_t)ST_STAGES * ST_STAGE_BYTES;
    uint64_t* full_bar = (uint64_t*)small;                         // [STAGES]
    uint64_t* empty_bar = full_bar + ST_STAGES;                    // [STAGES]
    double* xs = (double*)(small + 64);                            // [32]
    double* dsum = xs + ST_ROWS;                                   // [32]
    double* red = dsum + ST_ROWS;                                  // [32][4]
    SymTileDev* T = (SymTileDev*)(red + ST_ROWS * 4);              // [SYM_MAX_TILES] (dynamic indexing: keep it out of local memory)

    const int tid = threadIdx.x;
    const int64_t u0 = (int64_t)blockIdx.x * a.units_per_cta;
    int64_t u1 = u0 + a.units_per_cta;
    if (u1 > a.units_total) u1 = a.units_total;
    if (u0 >= u1) return;

    if (tid < a.ntiles) T[tid] = a.t[tid];
    if (tid == 0) {
        for (int i = 0; i < ST_STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], ST_CONSUMERS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int ntiles = a.ntiles;
    SymCursor c;
    sym_cursor_init(T, ntiles, u0, c);

    if (tid >= ST_CONSUMERS) {
        if (tid == ST_CONSUMERS) {  // producer lane
            uint64_t policy;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t u = u0; u < u1; ++u) {
                const SymTileDev& t = T[c.ti];
                const CUtensorMap* map = t.packed ? (t.tmaps + c.s / ST_BAND_STRIPS) : t.tmaps;
                const int row = t.packed ? (int)((c.s % ST_BAND_STRIPS) * ST_ROWS) : (int)(c.s * ST_ROWS);
                mbar_wait(&empty_bar[stage], phase ^ 1);
                mbar_expect_tx(&full_bar[stage], ST_STAGE_BYTES);
                tma_load_2d(tiles + (size_t)stage * (ST_ROWS * ST_COLS), map, &full_bar[stage], (int)(c.j * ST_COLS), row, policy);
                if (++stage == ST_STAGES) { stage = 0; phase ^= 1; }
                sym_cursor_next(T, ntiles, c);
            }
        }
        return;
    }

    const int lane = tid & 31, warp = tid >> 5;
    double acc[ST_ROWS];
    int stage = 0;
    uint32_t phase = 0;
    bool fresh = true;
    for (int64_t u = u0; u < u1; ++u) {
        const SymTileDev& t = T[c.ti];
        const int64_t s = c.s, j = c.j, nj = c.nj;
        const int64_t r0 = s * ST_ROWS;
        const int rows = (int)((t.nr - r0 < ST_ROWS) ? (t.nr - r0) : ST_ROWS);
        if (fresh) {
            if (tid < ST_ROWS) {
                xs[tid] = (tid < rows) ? t.xr[r0 + tid] : 0.0;
                dsum[tid] = 0.0;
            }
#pragma unroll
            for (int i = 0; i < ST_ROWS; ++i) acc[i] = 0.0;
            consumer_bar();
            fresh = false;
        }
        const double* tile = tiles + (size_t)stage * (ST_ROWS * ST_COLS);
        const int64_t col0 = j * ST_COLS;
        const int64_t n2 = t.diag ? r0 : t.nc;
        const int w = (int)((n2 - col0 < ST_COLS) ? (n2 - col0) : ST_COLS);
        const bool okx = 2 * tid < w, oky = 2 * tid + 1 < w;
        double2 xv = make_double2(0.0, 0.0);
        if (okx) xv.x = __ldg(t.xc + col0 + 2 * tid);
        if (oky) xv.y = __ldg(t.xc + col0 + 2 * tid + 1);
        mbar_wait(&full_bar[stage], phase);
        if (okx) {
            double2 cacc = make_double2(0.0, 0.0);
            const double2* tp = reinterpret_cast<const double2*>(tile) + tid;
#pragma unroll
            for (int i = 0; i < ST_ROWS; ++i) {
                double2 kv = tp[i * (ST_COLS / 2)];
                if (!oky) kv.y = 0.0;
                const double xrow = xs[i];
                acc[i] = fma(kv.y, xv.y, fma(kv.x, xv.x, acc[i]));
                cacc.x = fma(kv.x, xrow, cacc.x);
                cacc.y = fma(kv.y, xrow, cacc.y);
            }
            *reinterpret_cast<double2*>(t.ws + s * t.ld_ws + col0 + 2 * tid) = cacc;
        }
        if (t.diag && j == nj - 1) {
            const int r = tid >> 2, cp = tid & 3;
            double v = 0.0;
            if (r < rows)
                for (int cc = cp; cc < rows; cc += 4) v = fma(tile[r * ST_COLS + w + cc], xs[cc], v);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            if (cp == 0) dsum[r] = v;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == ST_STAGES) { stage = 0; phase ^= 1; }

        if (j + 1 == nj || u + 1 == u1) {
#pragma unroll
            for (int i = 0; i < ST_ROWS; ++i) {
                const double v = warp_sum(acc[i]);
                if (lane == 0) red[i * 4 + warp] = v;
            }
            consumer_bar();
            if (tid < ST_ROWS) {
                const int64_t first_unit = t.unit_base + st_units_before(s, t.diag, t.nc);
                const int slot = sym_row_slot(first_unit, a.units_per_cta, (int64_t)blockIdx.x);
                const double v = dsum[tid] + ((red[tid * 4 + 0] + red[tid * 4 + 1]) + (red[tid * 4 + 2] + red[tid * 4 + 3]));
                t.rowpart[(s * 2 + slot) * ST_ROWS + tid] = v;
            }
            consumer_bar();
            fresh = true;
        }
        sym_cursor_next(T, ntiles, c);
    }
}

// one output entry per thread group: column sums of its tile (+ for the diagonal tile: its own row sums and those of
// the off-diagonal tiles, tile order)
__global__ void __launch_bounds__(STR_COLS * STR_SPLIT)
symv_tma_multi_reduce_kernel(const __grid_constant__ SymMultiArgs a) {
    __shared__ double red[STR_SPLIT][STR_COLS];
    const int tc = threadIdx.x % STR_COLS, ts = threadIdx.x / STR_COLS;
    int ti = 0;
    while (ti + 1 < a.ntiles && a.t[ti + 1].blk_base <= (int64_t)blockIdx.x) ++ti;
    const SymTileDev& t = a.t[ti];
    const int64_t c = ((int64_t)blockIdx.x - t.blk_base) * STR_COLS + tc;
    double acc = 0.0;
    if (c < t.nc) {
        int64_t s = (t.diag ? (c / ST_ROWS + 1) : 0) + ts;
        const double* p = t.ws + c;
        for (; s + 7 * STR_SPLIT < t.nstrips; s += 8 * STR_SPLIT) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcs(p + (s + u * STR_SPLIT) * t.ld_ws);
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += v[u];
        }
        for (; s < t.nstrips; s += STR_SPLIT) acc += __ldcs(p + s * t.ld_ws);
    }
    red[ts][tc] = acc;
    __syncthreads();
    if (ts != 0 || c >= t.nc) return;
    acc = (red[0][tc] + red[1][tc]) + (red[2][tc] + red[3][tc]);
    if (t.diag) {
        const int64_t sr = c / ST_ROWS, i = c % ST_ROWS;
        acc += t.rowpart[(sr * 2) * ST_ROWS + i] + t.rowpart[(sr * 2 + 1) * ST_ROWS + i];
        for (int q = 0; q < a.ntiles; ++q) {
            if (q == ti) continue;
            const SymTileDev& o = a.t[q];
            const int64_t r = c - o.row_off;
            if (r >= 0 && r < o.nr) {
                const int64_t so = r / ST_ROWS, io = r % ST_ROWS;
                acc += o.rowpart[(so * 2) * ST_ROWS + io] + o.rowpart[(so * 2 + 1) * ST_ROWS + io];
            }
        }
    }
    t.out_c[c] = acc;
}

// ---- one-sided variant: y = A x for a wide matrix with FEW rows (the k x n_local preconditioner factor) ---------
// Same pipeline (persistent CTAs, TMA-fed 3-stage ring of 32 x 256 units, row sums in registers), no column sums.
// The register-staged gemv_rows kernel needs rows / 4 CTAs' worth of loads in flight and reaches ~0.76 of the HBM
// rate on 4839 rows; here the TMA keeps ~128 KB per SM in flight regardless of the row count.  A strip of 32 rows may
// be shared by several CTAs (few rows, many SMs): CTA b writes its partial row sums into slot b - first_cta(strip);
// a small kernel adds the slots in order (deterministic).
struct RowsTmaArgs {
    int64_t nr, nc, nstrips, ups, units_total, units_per_cta;
    int nslots;
    const CUtensorMap* tmap;
    const double* x;
    double* rowpart;  // [nstrips, nslots, 32]
};

__global__ void __launch_bounds__(ST_THREADS, 1) rows_tma_kernel(const RowsTmaArgs a) {
    extern __shared__ unsigned char st_smem_raw[];
    unsigned char* sm = (unsigned char*)(((uintptr_t)st_smem_raw + 1023) & ~(uintptr_t)1023);
    double* tiles = (double*)sm;
    unsigned char* small = sm + (size_t)ST_STAGES * ST_STAGE_BYTES;
    uint64_t* full_bar = (uint64_t*)small;
    uint64_t* empty_bar = full_bar + ST_STAGES;
    double* red = (double*)(small + 64);  // [32][4]

    const int tid = threadIdx.x;
    const int64_t u0 = (int64_t)blockIdx.x * a.units_per_cta;
    int64_t u1 = u0 + a.units_per_cta;
    if (u1 > a.units_total) u1 = a.units_total;
    if (u0 >= u1) return;
    if (tid == 0) {
        for (int i = 0; i < ST_STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], ST_CONSUMERS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int64_t s = u0 / a.ups, j = u0 % a.ups;

    if (tid >= ST_CONSUMERS) {
        if (tid == ST_CONSUMERS) {
            uint64_t policy;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t u = u0; u < u1; ++u) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                mbar_expect_tx(&full_bar[stage], ST_STAGE_BYTES);
                tma_load_2d(tiles + (size_t)stage * (ST_ROWS * ST_COLS), a.tmap, &full_bar[stage], (int)(j * ST_COLS),
                            (int)(s * ST_ROWS), policy);
                if (++stage == ST_STAGES) { stage = 0; phase ^= 1; }
                if (++j == a.ups) { ++s; j = 0; }
            }
        }
        return;
    }

    const int lane = tid & 31, warp = tid >> 5;
    double acc[ST_ROWS];
#pragma unroll
    for (int i = 0; i < ST_ROWS; ++i) acc[i] = 0.0;
    int stage = 0;
    uint32_t phase = 0;
    for (int64_t u = u0; u < u1; ++u) {
        const int64_t col0 = j * ST_COLS;
        const int w = (int)((a.nc - col0 < ST_COLS) ? (a.nc - col0) : ST_COLS);
        double2 xv = make_double2(0.0, 0.0);   // columns past nc are zero-filled by the tensor map; x = 0 there as well
        if (2 * tid < w) xv.x = __ldg(a.x + col0 + 2 * tid);
        if (2 * tid + 1 < w) xv.y = __ldg(a.x + col0 + 2 * tid + 1);
        mbar_wait(&full_bar[stage], phase);
        const double2* tp = reinterpret_cast<const double2*>(tiles + (size_t)stage * (ST_ROWS * ST_COLS)) + tid;
#pragma unroll
        for (int i = 0; i < ST_ROWS; ++i) {
            const double2 kv = tp[i * (ST_COLS / 2)];
            acc[i] = fma(kv.y, xv.y, fma(kv.x, xv.x, acc[i]));
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == ST_STAGES) { stage = 0; phase ^= 1; }
        if (j + 1 == a.ups || u + 1 == u1) {
#pragma unroll
            for (int i = 0; i < ST_ROWS; ++i) {
                const double v = warp_sum(acc[i]);
                if (lane == 0) red[i * 4 + warp] = v;
                acc[i] = 0.0;
            }
            consumer_bar();
            if (tid < ST_ROWS) {
                const int slot = (int)((int64_t)blockIdx.x - (s * a.ups) / a.units_per_cta);
                a.rowpart[(s * a.nslots + slot) * ST_ROWS + tid] =
                    (red[tid * 4 + 0] + red[tid * 4 + 1]) + (red[tid * 4 + 2] + red[tid * 4 + 3]);
            }
            consumer_bar();
        }
        if (++j == a.ups) { ++s; j = 0; }
    }
}

__global__ void rows_tma_reduce_kernel(const double* __restrict__ rowpart, int nslots, int64_t nr, double* __restrict__ y,
                                       double alpha) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nr) return;
    const int64_t s = r / ST_ROWS, i = r % ST_ROWS;
    double t = 0.0;
    for (int q = 0; q < nslots; ++q) t += rowpart[(s * nslots + q) * ST_ROWS + i];
    y[r] = alpha * t;
}

// columns: out_c[c] = post( [diag: rowsum(c)] + sum_{s >= s_min(c)} ws[s, c] );  rows (off-diagonal): out_r[r] += rowsum(r)
__global__ void __launch_bounds__(STR_COLS * STR_SPLIT)
symv_tma_reduce_kernel(const double* __restrict__ rowpart, const double* __restrict__ ws, int64_t ld_ws, int64_t nr,
                       int64_t nc, int64_t nstrips, int diag, double* __restrict__ out_c, double* __restrict__ out_r,
                       const double* __restrict__ x_shift, double alpha, double shift) {
    __shared__ double red[STR_SPLIT][STR_COLS];
    const int tc = threadIdx.x % STR_COLS, ts = threadIdx.x / STR_COLS;
    const int64_t ncb = (nc + STR_COLS - 1) / STR_COLS;
    if ((int64_t)blockIdx.x >= ncb) {  // row part of an off-diagonal tile
        const int64_t r = ((int64_t)blockIdx.x - ncb) * (STR_COLS * STR_SPLIT) + threadIdx.x;
        if (r < nr) {
            const int64_t sr = r / ST_ROWS, i = r % ST_ROWS;
            out_r[r] += rowpart[(sr * 2) * ST_ROWS + i] + rowpart[(sr * 2 + 1) * ST_ROWS + i];
        }
        return;
    }
    const int64_t c = (int64_t)blockIdx.x * STR_COLS + tc;
    double acc = 0.0;
    if (c < nc) {
        int64_t s = (diag ? (c / ST_ROWS + 1) : 0) + ts;
        const double* p = ws + c;
        for (; s + 7 * STR_SPLIT < nstrips; s += 8 * STR_SPLIT) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcs(p + (s + u * STR_SPLIT) * ld_ws);
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += v[u];
        }
        for (; s < nstrips; s += STR_SPLIT) acc += __ldcs(p + s * ld_ws);
    }
    red[ts][tc] = acc;
    __syncthreads();
    if (ts != 0 || c >= nc) return;
    acc = (red[0][tc] + red[1][tc]) + (red[2][tc] + red[3][tc]);
    if (diag) {
        const int64_t sr = c / ST_ROWS, i = c % ST_ROWS;
        acc += rowpart[(sr * 2) * ST_ROWS + i] + rowpart[(sr * 2 + 1) * ST_ROWS + i];
    }
    if (x_shift) {
        acc *= alpha;
        if (shift != 0.0) acc = fma(shift, x_shift[c], acc);
    }
    out_c[c] = acc;
}

// ---- host side ---------------------------------------------------------------------------------------
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode_tiled() {
    static encode_tiled_fn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (encode_tiled_fn)p;
    });
    return fn;
}

static int encode_map(CUtensorMap* out, const double* base, int64_t pitch, int64_t rows) {
    encode_tiled_fn enc = get_encode_tiled();
    if (!enc) { set_error("symv_tma: cuTensorMapEncodeTiled is not available from this driver"); return MLFFPC_ERR_CUDA; }
    const cuuint64_t gdim[2] = {(cuuint64_t)pitch, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)pitch * 8};
    const cuuint32_t box[2] = {ST_COLS, ST_ROWS};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("symv_tma: cuTensorMapEncodeTiled failed (%d) for pitch=%lld rows=%lld", (int)r, (long long)pitch,
                  (long long)rows);
        return MLFFPC_ERR_CUDA;
    }
    return MLFFPC_OK;
}

// Device-resident tensor maps, cached per (base, pitch, rows, layout): a PCG run asks for the same ones
// thousands of times.  The cache lives for the process (a few KB per entry).
struct TmapEntry {
    const void* base;
    int64_t ld, nr;
    int packed, device;
    CUtensorMap* dev;
};
static std::vector<TmapEntry> g_tmaps;
static std::mutex g_tmap_mu;

static int get_tmaps(const double* K, int64_t ld, int64_t nr, int packed, cudaStream_t s, const CUtensorMap** out) {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    int device = 0;
    MLFFPC_CUDA(cudaGetDevice(&device));
    for (const auto& e : g_tmaps)
        if (e.base == K && e.ld == ld && e.nr == nr && e.packed == packed && e.device == device) { *out = e.dev; return MLFFPC_OK; }
    std::vector<CUtensorMap> host;
    if (!packed) {
        host.resize(1);
        MLFFPC_TRY(encode_map(&host[0], K, ld, nr));
    } else {
        const int64_t nb = (nr + ST_BAND_ROWS - 1) / ST_BAND_ROWS;
        host.resize((size_t)nb);
        for (int64_t b = 0; b < nb; ++b) {
            const int64_t rows = (nr - b * ST_BAND_ROWS < ST_BAND_ROWS) ? (nr - b * ST_BAND_ROWS) : ST_BAND_ROWS;
            MLFFPC_TRY(encode_map(&host[(size_t)b], K + st_band_off(b), st_band_pitch(b), rows));
        }
    }
    TmapEntry e;
    e.base = K; e.ld = ld; e.nr = nr; e.packed = packed; e.device = device;
    MLFFPC_CUDA(cudaMalloc(&e.dev, host.size() * sizeof(CUtensorMap)));
    // synchronous copy: the maps must be resident before the first kernel that names them runs on any stream
    MLFFPC_CUDA(cudaMemcpy(e.dev, host.data(), host.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    (void)s;
    if (g_tmaps.size() >= 32) {
        cudaFree(g_tmaps.front().dev);
        g_tmaps.erase(g_tmaps.begin());
    }
    g_tmaps.push_back(e);
    *out = e.dev;
    return MLFFPC_OK;
}

struct RowsPlan {
    int64_t nstrips, ups, units_total, units_per_cta, ncta;
    int nslots;
};
static RowsPlan rows_plan(int64_t nr, int64_t nc, int num_sms) {
    RowsPlan p;
    p.nstrips = (nr + ST_ROWS - 1) / ST_ROWS;
    p.ups = (nc + ST_COLS - 1) / ST_COLS;
    p.units_total = p.nstrips * p.ups;
    p.units_per_cta = (p.units_total + num_sms - 1) / num_sms;
    if (p.units_per_cta < 4) p.units_per_cta = 4;  // a CTA should amortise its pipeline fill
    p.ncta = (p.units_total + p.units_per_cta - 1) / p.units_per_cta;
    p.nslots = (int)((p.ups + p.units_per_cta - 1) / p.units_per_cta) + 1;
    return p;
}
int64_t rows_tma_ws_doubles(int64_t nr, int64_t nc, int num_sms) {
    const RowsPlan p = rows_plan(nr, nc, num_sms);
    return p.nstrips * p.nslots * ST_ROWS + 64;
}
bool rows_tma_usable(const double* A, int64_t ld, int64_t nr, int64_t nc) {
    return ((uintptr_t)A % 16 == 0) && (ld % 2 == 0) && ld < ((int64_t)1 << 31) && nr < ((int64_t)1 << 31) && nc >= 1024 &&
           get_encode_tiled() != nullptr;
}
// y[nr] = alpha * A x for row-major A[nr, nc] (pitch ld); wsd: rows_tma_ws_doubles(nr, nc, num_sms) doubles
int rows_gemv_tma(mlffpc_ctx* ctx, const double* A, int64_t nr, int64_t nc, int64_t ld, const double* x, double* y,
                  double alpha, double* wsd, cudaStream_t s) {
    RowsTmaArgs a;
    // the map spans exactly nc columns: the TMA zero-fills the ragged last unit (pitch ld in the global stride)
    {
        std::lock_guard<std::mutex> lk(g_tmap_mu);
        int device = 0;
        MLFFPC_CUDA(cudaGetDevice(&device));
        const CUtensorMap* found = nullptr;
        for (const auto& e : g_tmaps)
            if (e.base == A && e.ld == ld && e.nr == nr && e.packed == (int)(2 + (nc & 0xffffff)) && e.device == device) found = e.dev;
        if (!found) {
            encode_tiled_fn enc = get_encode_tiled();
            if (!enc) { set_error("rows_gemv_tma: cuTensorMapEncodeTiled is not available from this driver"); return MLFFPC_ERR_CUDA; }
            CUtensorMap host;
            const cuuint64_t gdim[2] = {(cuuint64_t)nc, (cuuint64_t)nr};
            const cuuint64_t gstride[1] = {(cuuint64_t)ld * 8};
            const cuuint32_t box[2] = {ST_COLS, ST_ROWS};
            const cuuint32_t estr[2] = {1, 1};
            const CUresult r = enc(&host, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(A), gdim, gstride, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { set_error("rows_gemv_tma: cuTensorMapEncodeTiled failed (%d)", (int)r); return MLFFPC_ERR_CUDA; }
            TmapEntry e;
            e.base = A; e.ld = ld; e.nr = nr; e.packed = (int)(2 + (nc & 0xffffff)); e.device = device;
            MLFFPC_CUDA(cudaMalloc(&e.dev, sizeof(CUtensorMap)));
            MLFFPC_CUDA(cudaMemcpy(e.dev, &host, sizeof(CUtensorMap), cudaMemcpyHostToDevice));
            if (g_tmaps.size() >= 32) {
                cudaFree(g_tmaps.front().dev);
                g_tmaps.erase(g_tmaps.begin());
            }
            g_tmaps.push_back(e);
            found = e.dev;
        }
        a.tmap = found;
    }
    const RowsPlan p = rows_plan(nr, nc, ctx->num_sms);
    a.nr = nr; a.nc = nc; a.nstrips = p.nstrips; a.ups = p.ups; a.units_total = p.units_total;
    a.units_per_cta = p.units_per_cta; a.nslots = p.nslots;
    a.x = x; a.rowpart = wsd;
    MLFFPC_CUDA(cudaMemsetAsync(wsd, 0, (size_t)(p.nstrips * p.nslots * ST_ROWS) * 8, s));
    if (!ctx->tma_attr_rows) {
        MLFFPC_CUDA(cudaFuncSetAttribute(rows_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ST_SMEM));
        ctx->tma_attr_rows = true;
    }
    rows_tma_kernel<<<(unsigned)p.ncta, ST_THREADS, ST_SMEM, s>>>(a);
    MLFFPC_LAUNCH_CHECK();
    rows_tma_reduce_kernel<<<(unsigned)((nr + 255) / 256), 256, 0, s>>>(wsd, p.nslots, nr, y, alpha);
    MLFFPC_LAUNCH_CHECK();
    return MLFFPC_OK;
}

int64_t symv_tma_ws_doubles(int64_t nr, int64_t nc) {
    const int64_t nstrips = (nr + ST_ROWS - 1) / ST_ROWS;
    const int64_t ld_ws = ((nc + 1) & ~(int64_t)1) + ST_COLS;  // the last unit of a strip may store past nc
    return nstrips * ld_ws + nstrips * 2 * ST_ROWS + 64;
}

bool symv_tma_usable(const double* K, int64_t ld, int64_t nr) {
    return ((uintptr_t)K % 16 == 0) && (ld % 2 == 0) && ld < ((int64_t)1 << 31) && nr < ((int64_t)1 << 31) &&
           get_encode_tiled() != nullptr;
}

// One tile pass on the TMA path.
//   columns: out_c[c] = post([diag: row sums] + column sums), rows of off-diagonal tiles: out_r[r] += row sums,
//   post(v) = alpha v + shift x_shift[c] when x_shift != NULL.  wsd: symv_tma_ws_doubles(nr, nc) doubles.
int symv_tile_tma(mlffpc_ctx* ctx, const double* K, int64_t ld, int64_t nr, int64_t nc, int diag, int packed,
                  const double* xr, const double* xc, double* wsd, double* out_c, double* out_r,
                  const double* x_shift, double alpha, double shift, cudaStream_t s) {
    SymTmaArgs a;
    MLFFPC_TRY(get_tmaps(K, ld, nr, packed, s, &a.tmaps));
    a.nr = nr; a.nc = nc; a.diag = diag; a.packed = packed;
    a.nstrips = (nr + ST_ROWS - 1) / ST_ROWS;
    a.ld_ws = ((nc + 1) & ~(int64_t)1) + ST_COLS;
    a.units_total = st_units_before(a.nstrips, diag, nc);
    const int64_t max_in_strip = st_units_in_strip(a.nstrips - 1, diag, nc);
    int64_t ncta = a.units_total / max_in_strip;  // every CTA gets >= one strip's worth: a strip spans <= 2 CTAs
    if (ncta > ctx->num_sms) ncta = ctx->num_sms;
    if (ncta < 1) ncta = 1;
    a.units_per_cta = (a.units_total + ncta - 1) / ncta;
    if (a.units_per_cta < max_in_strip) a.units_per_cta = max_in_strip;
    ncta = (a.units_total + a.units_per_cta - 1) / a.units_per_cta;
    a.xr = xr; a.xc = xc;
    a.ws = wsd;
    a.rowpart = wsd + a.nstrips * a.ld_ws;
    MLFFPC_CUDA(cudaMemsetAsync(a.rowpart, 0, (size_t)a.nstrips * 2 * ST_ROWS * 8, s));
    if (!ctx->tma_attr_symv) {  // per context (= per device): cudaFuncSetAttribute applies to the current device only
        MLFFPC_CUDA(cudaFuncSetAttribute(symv_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ST_SMEM));
        ctx->tma_attr_symv = true;
    }
    symv_tma_kernel<<<(unsigned)ncta, ST_THREADS, ST_SMEM, s>>>(a);
    MLFFPC_LAUNCH_CHECK();
    const int64_t ncb = (nc + STR_COLS - 1) / STR_COLS;
    const int64_t nrb = diag ? 0 : (nr + STR_COLS * STR_SPLIT - 1) / (STR_COLS * STR_SPLIT);
    symv_tma_reduce_kernel<<<(unsigned)(ncb + nrb), STR_COLS * STR_SPLIT, 0, s>>>(
        a.rowpart, a.ws, a.ld_ws, nr, nc, a.nstrips, diag, out_c, out_r, x_shift, alpha, shift);
    MLFFPC_LAUNCH_CHECK();
    return MLFFPC_OK;
}

// All tiles of a rank in one pass (tile 0 must be the diagonal tile; the others' rows lie inside its row block).
//   out_c of tile i receives its column sums; the diagonal tile's out_c also the row sums of every tile.
int symv_tiles_tma(mlffpc_ctx* ctx, int ntiles, const SymTileIn* in, double* wsd, cudaStream_t s) {
    MLFFPC_REQUIRE(ntiles >= 1 && ntiles <= SYM_MAX_TILES && in[0].diag, "symv_tiles: bad tile list");
    SymMultiArgs a;
    a.ntiles = ntiles;
    int64_t units = 0, max_in_strip = 1, ws_off = 0, rp_total = 0, blk = 0;
    for (int i = 0; i < ntiles; ++i) {
        SymTileDev& t = a.t[i];
        MLFFPC_TRY(get_tmaps(in[i].K, in[i].ld, in[i].nr, in[i].packed, s, &t.tmaps));
        t.nr = in[i].nr; t.nc = in[i].nc; t.diag = in[i].diag; t.packed = in[i].packed;
        t.nstrips = (t.nr + ST_ROWS - 1) / ST_ROWS;
        t.ld_ws = ((t.nc + 1) & ~(int64_t)1) + ST_COLS;
        t.unit_base = units;
        units += st_units_before(t.nstrips, t.diag, t.nc);
        const int64_t mis = st_units_in_strip(t.nstrips - 1, t.diag, t.nc);
        if (mis > max_in_strip) max_in_strip = mis;
        t.row_off = in[i].row_off;
        t.xr = in[i].xr; t.xc = in[i].xc; t.out_c = in[i].out_c;
        t.ws = wsd + ws_off;
        ws_off += t.nstrips * t.ld_ws;
        rp_total += t.nstrips * 2 * ST_ROWS;
        t.ncb = (t.nc + STR_COLS - 1) / STR_COLS;
        t.blk_base = blk;
        blk += t.ncb;
    }
    double* rp = wsd + ws_off;
    for (int i = 0; i < ntiles; ++i) { a.t[i].rowpart = rp; rp += a.t[i].nstrips * 2 * ST_ROWS; }
    for (int i = ntiles; i < SYM_MAX_TILES; ++i) a.t[i] = a.t[0];
    a.units_total = units;
    int64_t ncta = 1;
    sym_cta_split(units, max_in_strip, ctx->num_sms, &a.units_per_cta, &ncta);
    MLFFPC_CUDA(cudaMemsetAsync(wsd + ws_off, 0, (size_t)rp_total * 8, s));
    if (!ctx->tma_attr_multi) {
        MLFFPC_CUDA(cudaFuncSetAttribute(symv_tma_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ST_SMEM));
        ctx->tma_attr_multi = true;
    }
    symv_tma_multi_kernel<<<(unsigned)ncta, ST_THREADS, ST_SMEM, s>>>(a);
    MLFFPC_LAUNCH_CHECK();
    symv_tma_multi_reduce_kernel<<<(unsigned)blk, STR_COLS * STR_SPLIT, 0, s>>>(a);
    MLFFPC_LAUNCH_CHECK();
    return MLFFPC_OK;
}

}  // namespace mlffpc
