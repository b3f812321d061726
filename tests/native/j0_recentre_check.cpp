// Host-side restatement of the J0 phase of k_gram (frank_b200/csrc/fb_gram.cu: select / prepare / j0_gemm), compiled by
// tests/test_host_logic.py with g++ (no GPU): for random tiles of 64 baseline-sorted visibilities and random columns it
//   * picks the table row as the kernel does (middle of the tile's range of arguments) and applies its validity test,
//   * shifts the row's polynomial to the tile centre (repeated synthetic division, same FMA order) and rescales it to the
//     tile variable s = (a - a_c) j_ref,
//   * forms sqrt(w) s^d as the kernel does and accumulates the rank-8 product (degrees 4-7 first, then 0-3),
// and prints the largest deviation of G / sqrt(w) from glibc's 80-bit j0l(a j_k), per range of arguments, together with the
// share of (column, tile) pairs the validity test rejected (those take the per-visibility path, checked by j0_table_check).
#include "../../frank_b200/csrc/fb_j0_table.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>

int main(int argc, char **argv)
{
    const double x_max = argc > 1 ? atof(argv[1]) : 6400.0;
    const long n_tiles = argc > 2 ? atol(argv[2]) : 20000;
    std::vector<double> tab;
    fb_j0_build(x_max, tab);
    const int rows = fb_j0_rows_for(x_max), last_row = rows - 1;
    std::mt19937_64 rng(777);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    double worst[4] = {0, 0, 0, 0};
    long rejected = 0, total = 0;
    for (long it = 0; it < n_tiles; it++) {
        // a column and the block's reference column (the largest of the block)
        const double jk = 2.4 + U(rng) * (x_max * 0.98 - 2.4);
        const double jref = jk + U(rng) * (x_max * 0.98 - jk);
        // a tile: 64 sorted baselines around a_c; its range of arguments at this column is up to 0.1 wide
        const double xc = U(rng) < 0.3 ? U(rng) * 30.0 : U(rng) * x_max * 0.97;
        const double width = (U(rng) < 0.5 ? 0.02 : 0.1) * U(rng);
        double a[64], sw[64];
        for (int v = 0; v < 64; v++) { a[v] = (xc + (U(rng) - 0.5) * width) / jk; if (a[v] < 0) a[v] = -a[v]; sw[v] = 0.5 + 99.5 * U(rng); }
        std::sort(a, a + 64);
        const double amin = a[0], amax = a[63];
        // select (row for the middle of the range) and prepare's validity test
        const double xlo = amin * jk, xhi = amax * jk;
        int m = (int)std::nearbyint((xlo + xhi) * (0.5 * FB_J0_INVH));
        m = std::min(m, last_row);
        const double cen = (double)m * FB_J0_H;
        total++;
        if (!(std::fabs(xlo - cen) < FB_J0_ACCEPT && std::fabs(xhi - cen) < FB_J0_ACCEPT)) { rejected++; continue; }
        // prepare: shift to the tile centre, rescale
        const double ac = 0.5 * (amin + amax);
        const double e = std::fma(ac, jk, -cen);
        double c[8];
        for (int k = 0; k < 8; k++) c[k] = tab[(size_t)m * FB_J0_ROWLEN + k];
        for (int lo = 0; lo <= 6; lo++)
            for (int k = 6; k >= lo; k--) c[k] = std::fma(e, c[k + 1], c[k]);
        const double r = jk / jref, r2 = r * r, r4 = r2 * r2;
        c[1] *= r; c[2] *= r2; c[3] *= r2 * r; c[4] *= r4; c[5] *= r4 * r; c[6] *= r4 * r2; c[7] *= r4 * (r2 * r);
        for (int v = 0; v < 64; v++) {
            const double sv = (a[v] - ac) * jref, s2 = sv * sv, s4 = s2 * s2;
            double p[8];
            p[0] = sw[v]; p[1] = p[0] * sv; p[2] = p[0] * s2; p[3] = p[1] * s2;
            for (int d = 0; d < 4; d++) p[4 + d] = p[d] * s4;
            double g = 0.0;
            for (int d = 4; d < 8; d++) g = std::fma(p[d], c[d], g);
            for (int d = 0; d < 4; d++) g = std::fma(p[d], c[d], g);
            const long double x = (long double)a[v] * (long double)jk;
            const double err = std::fabs((double)((long double)g / (long double)sw[v] - j0l(x)));
            const double xd = (double)x;
            const int cls = xd < 5 ? 0 : xd < 30 ? 1 : xd < 200 ? 2 : 3;
            worst[cls] = std::max(worst[cls], err);
        }
    }
    printf("%.6e %.6e %.6e %.6e %.6f\n", worst[0], worst[1], worst[2], worst[3], (double)rejected / (double)total);
    return 0;
}
