// Host-only check of the unit / band arithmetic shared by the symmetric operator (csrc/symlayout.cuh):
// closed forms against brute-force enumeration.  Built and run by tests/test_native_symlayout.py (no GPU needed).
#include <cstdio>
#include <cstdlib>
#include <utility>
#include <vector>

#include "../../mlff_preconditioner_b200/csrc/symlayout.cuh"

using namespace mlffpc;


// ---- one-launch pass over several tiles: every (tile, strip, unit) is visited exactly once and in order, a strip is
// shared by at most two CTAs, and these two write different row-sum slots
struct TileH {
    int64_t nstrips, nc, unit_base;
    int diag;
};
static int check_multi(const std::vector<std::pair<int64_t, int64_t>>& shapes, int64_t max_ctas) {  // (nr, nc); tile 0 diagonal
    int fails = 0;
    std::vector<TileH> t(shapes.size());
    int64_t units = 0, max_in_strip = 1;
    for (size_t i = 0; i < shapes.size(); ++i) {
        t[i].diag = (i == 0);
        t[i].nstrips = (shapes[i].first + ST_ROWS - 1) / ST_ROWS;
        t[i].nc = shapes[i].second;
        t[i].unit_base = units;
        units += st_units_before(t[i].nstrips, t[i].diag, t[i].nc);
        const int64_t mis = st_units_in_strip(t[i].nstrips - 1, t[i].diag, t[i].nc);
        if (mis > max_in_strip) max_in_strip = mis;
    }
    int64_t upc = 0, ncta = 0;
    sym_cta_split(units, max_in_strip, max_ctas, &upc, &ncta);
    if (upc < max_in_strip || ncta > max_ctas || ncta * upc < units || (ncta - 1) * upc >= units) { printf("cta split\n"); ++fails; }
    // expected order
    int ti = 0;
    int64_t s = 0, j = 0, visited = 0;
    std::vector<std::vector<int>> writers(shapes.size());
    std::vector<std::vector<int>> slots_used(shapes.size());
    for (size_t i = 0; i < shapes.size(); ++i) { writers[i].assign((size_t)t[i].nstrips, 0); slots_used[i].assign((size_t)t[i].nstrips, 0); }
    for (int64_t b = 0; b < ncta; ++b) {
        const int64_t u0 = b * upc, u1 = (u0 + upc < units) ? (u0 + upc) : units;
        SymCursor c;
        sym_cursor_init(t.data(), (int)t.size(), u0, c);
        for (int64_t u = u0; u < u1; ++u) {
            if (c.ti != ti || c.s != s || c.j != j) { printf("order at unit %lld\n", (long long)u); return fails + 1; }
            if (c.nj != st_units_in_strip(s, t[ti].diag, t[ti].nc)) ++fails;
            if (c.j + 1 == c.nj || u + 1 == u1) {  // this CTA flushes its share of the strip's row sums
                const int64_t first_unit = t[ti].unit_base + st_units_before(s, t[ti].diag, t[ti].nc);
                const int slot = sym_row_slot(first_unit, upc, b);
                if (slots_used[ti][(size_t)s] & (1 << slot)) { printf("slot written twice: tile %d strip %lld\n", ti, (long long)s); ++fails; }
                slots_used[ti][(size_t)s] |= 1 << slot;
                ++writers[ti][(size_t)s];
            }
            ++visited;
            if (++j == st_units_in_strip(s, t[ti].diag, t[ti].nc)) { j = 0; if (++s == t[ti].nstrips) { s = 0; ++ti; } }
            sym_cursor_next(t.data(), (int)t.size(), c);
        }
    }
    if (visited != units || ti != (int)t.size()) { printf("coverage %lld of %lld\n", (long long)visited, (long long)units); ++fails; }
    for (size_t i = 0; i < shapes.size(); ++i)
        for (int64_t q = 0; q < t[i].nstrips; ++q)
            if (writers[i][(size_t)q] < 1 || writers[i][(size_t)q] > 2) { printf("strip with %d writers\n", writers[i][(size_t)q]); ++fails; }
    return fails;
}

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
    // tile lists of a rank: cfg2 on 8 / 4 / 2 ranks (rank < world / 2 and >= world / 2), cfg4 on 8, tiny and ragged ones
    typedef std::vector<std::pair<int64_t, int64_t>> Shapes;
    const Shapes lists[] = {
        {{13500, 13500}, {13500, 13500}, {13500, 13500}, {13500, 13500}, {6750, 13500}},
        {{13500, 13500}, {13500, 13500}, {13500, 13500}, {13500, 13500}, {13500, 6750}},
        {{27000, 27000}, {27000, 27000}, {13500, 27000}},
        {{54000, 54000}, {54000, 27000}},
        {{33750, 33750}, {33750, 33750}, {33750, 33750}, {33750, 33750}, {33750, 16875}},
        {{108000, 108000}},
        {{27, 27}},
        {{324, 324}, {324, 297}, {162, 324}},
        {{1000, 1000}, {1000, 37}, {33, 1000}},
    };
    for (const Shapes& l : lists)
        for (int64_t ctas : {1, 2, 3, 7, 148, 296}) fails += check_multi(l, ctas);
    printf(fails ? "SYMLAYOUT FAIL %d\n" : "SYMLAYOUT OK\n", fails);
    return fails ? 1 : 0;
}
