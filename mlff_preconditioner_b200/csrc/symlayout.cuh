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

}  // namespace mlffpc
