// 2-D block decomposition of the image over `workers` ranks.
//
// pconv_divide_rows reproduces the reference chooser's *decisions* (mpi/mpi_conv.c:363-377:
// exact divisibility, minimise rows/r + cols/c, first minimum wins) so that tile grids stay
// comparable with `mpirun -np N ./mpi_conv` (1920x2520 -> 1x1, 2x1, 2x2, 4x2).
// pconv_choose_grid extends it to sizes the reference rejects (`Cannot divide to processes`,
// mpi/mpi_conv.c:57-61) by allowing remainder tiles.
#include <cstdint>

#include "pconv_host.h"

extern "C" int pconv_divide_rows(int64_t rows, int64_t cols, int workers) {
    if (workers <= 0 || rows <= 0 || cols <= 0) return 0;
    int best = 0;
    int64_t best_cost = rows + cols + 1;
    for (int r = 1; r <= workers; ++r) {
        if (workers % r != 0) continue;
        const int c = workers / r;
        if (rows % r != 0 || cols % c != 0) continue;
        const int64_t cost = rows / r + cols / c;  // semi-perimeter of one tile
        if (cost < best_cost) {
            best_cost = cost;
            best = r;
        }
    }
    return best;
}

extern "C" int pconv_choose_grid(int64_t rows, int64_t cols, int workers, int* row_div,
                                 int* col_div, int* exact) {
    if (workers <= 0 || rows <= 0 || cols <= 0 || !row_div || !col_div) return -1;
    const int r_exact = pconv_divide_rows(rows, cols, workers);
    if (r_exact > 0) {
        *row_div = r_exact;
        *col_div = workers / r_exact;
        if (exact) *exact = 1;
        return 0;
    }
    // Remainder tiles: largest tile is ceil(rows/r) x ceil(cols/c); every rank needs >= 1 pixel.
    int best = 0;
    int64_t best_cost = -1;
    for (int r = 1; r <= workers; ++r) {
        if (workers % r != 0) continue;
        const int c = workers / r;
        if (r > rows || c > cols) continue;
        const int64_t cost = (rows + r - 1) / r + (cols + c - 1) / c;
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            best = r;
        }
    }
    if (best == 0) return -2;  // more workers than pixels along every factorisation
    *row_div = best;
    *col_div = workers / best;
    if (exact) *exact = 0;
    return 0;
}

extern "C" void pconv_split(int64_t extent, int parts, int idx, int64_t* start, int64_t* size) {
    const int64_t base = extent / parts, rem = extent % parts;
    const int64_t s = (int64_t)idx * base + (idx < rem ? idx : rem);
    if (start) *start = s;
    if (size) *size = base + (idx < rem ? 1 : 0);
}
