// Host-only check of the unit / band arithmetic shared by the symmetric operator (csrc/symlayout.cuh):
// closed forms against brute-force enumeration.  Built and run by tests/test_native_symlayout.py (no GPU needed).
#include <cstdio>
#include <cstdlib>

#include "../../mlff_preconditioner_b200/csrc/symlayout.cuh"

using namespace mlffpc;

int main() {
    int fails = 0;
    // units of a diagonal tile: strip s owns columns [0, 32 s + 32) in chunks of 256
    for (int64_t nstrips : {1, 7, 8, 9, 64, 100, 3375}) {
        int64_t before = 0;
        for (int64_t s = 0; s < nstrips; ++s) {
            const int64_t brute = (32 * s + 32 + ST_COLS - 1) / ST_COLS;  // chunks needed to reach the diagonal block
            if (st_units_in_strip(s, 1, 0) != brute) { printf("units_in_strip(%lld)\n", (long long)s); ++fails; }
            if (st_units_before(s, 1, 0) != before) { printf("units_before(%lld)\n", (long long)s); ++fails; }
            before += brute;
        }
        if (st_units_before(nstrips, 1, 0) != before) ++fails;
    }
    // rectangular tiles
    for (int64_t nc : {1, 255, 256, 257, 13500, 54001})
        for (int64_t s : {0, 1, 5, 1000}) {
            const int64_t per = (nc + ST_COLS - 1) / ST_COLS;
            if (st_units_in_strip(s, 0, nc) != per || st_units_before(s, 0, nc) != s * per) ++fails;
        }
    // packed bands: offsets are the running sum of rows * pitch, pitches cover the diagonal block of every strip
    int64_t off = 0;
    for (int64_t b = 0; b < 500; ++b) {
        if (st_band_off(b) != off) { printf("band_off(%lld)\n", (long long)b); ++fails; }
        if (st_band_pitch(b) != ST_COLS * (b + 1)) ++fails;
        for (int64_t s = b * ST_BAND_STRIPS; s < (b + 1) * ST_BAND_STRIPS; ++s)
            if (32 * s + 32 > st_band_pitch(b)) { printf("pitch too small for strip %lld\n", (long long)s); ++fails; }
        off += (int64_t)ST_BAND_ROWS * st_band_pitch(b);
    }
    for (int64_t nr : {1, 255, 256, 257, 324, 108000, 270000}) {
        int64_t e = 0;
        for (int64_t r = 0; r < nr; ++r) e += st_band_pitch(r / ST_BAND_ROWS);
        if (st_packed_elems(nr) != e) { printf("packed_elems(%lld)\n", (long long)nr); ++fails; }
    }
    printf(fails ? "SYMLAYOUT FAIL %d\n" : "SYMLAYOUT OK\n", fails);
    return fails ? 1 : 0;
}
