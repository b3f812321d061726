// Host-side check of frank_b200/csrc/fb_j0_table.h (compiled by tests/test_host_logic.py with g++, no GPU):
// builds the table the kernels use, evaluates each sample with the row the kernels may pick in the WORST case
// (|t| up to the acceptance limit 1/16) and with the nearest row (gather path), and prints the largest absolute
// deviations from glibc's 80-bit j0l per argument range.
#include "../../frank_b200/csrc/fb_j0_table.h"

#include <cstdio>
#include <cstdlib>
#include <random>

static double eval_row(const std::vector<double> &tab, int m, double x)
{
    const double u = x - m * FB_J0_H;
    const double *p = &tab[(size_t)m * FB_J0_ROWLEN];
    double g = std::fma(p[7], u, p[6]);
    for (int k = 5; k >= 0; k--) g = std::fma(g, u, p[k]);
    return g;
}

int main(int argc, char **argv)
{
    const double x_max = argc > 1 ? atof(argv[1]) : 1000.0;
    const long n = argc > 2 ? atol(argv[2]) : 2000000;
    std::vector<double> tab;
    fb_j0_build(x_max, tab);
    const int rows = fb_j0_rows_for(x_max);
    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    double worst_far[4] = {0, 0, 0, 0}, worst_near[4] = {0, 0, 0, 0};
    for (long it = 0; it < n; it++) {
        const double x = U(rng) < 0.3 ? U(rng) * 30.0 : U(rng) * x_max;
        const int cls = x < 5 ? 0 : x < 30 ? 1 : x < 200 ? 2 : 3;
        const double exact = (double)j0l((long double)x);
        const long double exact_l = j0l((long double)x);
        int m = (int)llround(x * FB_J0_INVH);
        if (m > rows - 1) m = rows - 1;
        const double t = x - m * FB_J0_H;
        int m2 = t >= 0 ? m + 1 : m - 1;
        if (m2 < 0 || m2 > rows - 1) m2 = m;
        const double e_near = std::fabs((double)((long double)eval_row(tab, m, x) - exact_l));
        const double e_far = std::fabs((double)((long double)eval_row(tab, m2, x) - exact_l));
        if (e_near > worst_near[cls]) worst_near[cls] = e_near;
        if (e_far > worst_far[cls]) worst_far[cls] = e_far;
        (void)exact;
    }
    printf("%.6e %.6e %.6e %.6e %.6e %.6e %.6e %.6e\n", worst_near[0], worst_near[1], worst_near[2], worst_near[3],
           worst_far[0], worst_far[1], worst_far[2], worst_far[3]);
    return 0;
}
