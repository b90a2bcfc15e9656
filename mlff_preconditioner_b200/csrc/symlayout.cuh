// Unit / band arithmetic shared by the symmetric operator's plan (symop.cu), its TMA kernel (symtma.cu) and
// the packed tile assembly (geometry.cu).
#pragma once
#include <stdint.h>

namespace mlffpc {

constexpr int ST_ROWS = 32;                       // rows per strip / unit
constexpr int ST_COLS = 256;                      // columns per unit
constexpr int ST_BAND_STRIPS = ST_COLS / ST_ROWS;  // strips per band
constexpr int ST_BAND_ROWS = ST_COLS;             // rows per band of the packed diagonal layout

// units of strip s: diagonal tiles own columns [0, 32 s + 32) -> s/8 + 1 units; others ceil(nc / 256)
__host__ __device__ __forceinline__ int64_t st_units_in_strip(int64_t s, int diag, int64_t nc) {
    return diag ? (s / ST_BAND_STRIPS + 1) : ((nc + ST_COLS - 1) / ST_COLS);
}
// units of strips [0, s)
__host__ __device__ __forceinline__ int64_t st_units_before(int64_t s, int diag, int64_t nc) {
    if (!diag) return s * ((nc + ST_COLS - 1) / ST_COLS);
    const int64_t q = s / ST_BAND_STRIPS, rem = s % ST_BAND_STRIPS;
    return s + (ST_BAND_STRIPS / 2) * q * (q - 1) + rem * q;
}
// packed diagonal layout: band b = rows [256 b, 256 b + 256) stores columns [0, 256 (b + 1)) with that pitch
__host__ __device__ __forceinline__ int64_t st_band_pitch(int64_t b) { return (int64_t)ST_COLS * (b + 1); }
__host__ __device__ __forceinline__ int64_t st_band_off(int64_t b) {
    return (int64_t)ST_BAND_ROWS * ST_COLS * (b * (b + 1) / 2);
}
__host__ __device__ __forceinline__ int64_t st_packed_elems(int64_t nr) {
    const int64_t nb = (nr + ST_BAND_ROWS - 1) / ST_BAND_ROWS;
    if (nb == 0) return 0;
    return st_band_off(nb - 1) + (nr - (nb - 1) * ST_BAND_ROWS) * st_band_pitch(nb - 1);
}

// ---- walking the concatenated unit lists of several tiles (the one-launch pass, symtma.cu) ------------------------
// Tile type T needs: nstrips, nc, unit_base (global index of its first unit), diag.  CTA b owns the global units
// [b upc, (b + 1) upc); upc >= the longest strip of any tile, so a strip is shared by at most two CTAs: the one that
// starts it writes row-sum slot 0, the other slot 1.
struct SymCursor {
    int ti;
    int64_t s, j, nj;  // tile, strip, unit within the strip, units of the strip
};
template <class T>
__host__ __device__ __forceinline__ void sym_cursor_init(const T* tiles, int ntiles, int64_t u0, SymCursor& c) {
    int ti = 0;
    while (ti + 1 < ntiles && tiles[ti + 1].unit_base <= u0) ++ti;
    const int64_t lu = u0 - tiles[ti].unit_base;
    int64_t lo = 0, hi = tiles[ti].nstrips;
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (st_units_before(mid, tiles[ti].diag, tiles[ti].nc) <= lu) lo = mid; else hi = mid;
    }
    c.ti = ti;
    c.s = lo;
    c.j = lu - st_units_before(lo, tiles[ti].diag, tiles[ti].nc);
    c.nj = st_units_in_strip(lo, tiles[ti].diag, tiles[ti].nc);
}
template <class T>
__host__ __device__ __forceinline__ void sym_cursor_next(const T* tiles, int ntiles, SymCursor& c) {
    if (++c.j < c.nj) return;
    c.j = 0;
    if (++c.s == tiles[c.ti].nstrips) { ++c.ti; c.s = 0; }
    c.nj = (c.ti < ntiles) ? st_units_in_strip(c.s, tiles[c.ti].diag, tiles[c.ti].nc) : 1;
}
// row-sum slot of CTA `cta` for the strip whose first unit has global index first_unit
__host__ __device__ __forceinline__ int sym_row_slot(int64_t first_unit, int64_t units_per_cta, int64_t cta) {
    return (first_unit / units_per_cta == cta) ? 0 : 1;
}
// units per CTA and CTA count for `units` units on at most max_ctas CTAs
__host__ __device__ __forceinline__ void sym_cta_split(int64_t units, int64_t max_in_strip, int64_t max_ctas,
                                                       int64_t* units_per_cta, int64_t* ncta) {
    int64_t n = units / max_in_strip;
    if (n > max_ctas) n = max_ctas;
    if (n < 1) n = 1;
    int64_t upc = (units + n - 1) / n;
    if (upc < max_in_strip) upc = max_in_strip;
    *units_per_cta = upc;
    *ncta = (units + upc - 1) / upc;
}

}  // namespace mlffpc
