// Host-side construction of the piecewise-polynomial J0 table used by the Gram and predict kernels.
// Plain C++ (no CUDA) so that tests can compile it with g++ and check it against 80-bit j0l.
//
// Row m holds the degree-7 polynomial that approximates J0(c + t), c = m / 16, for |t| <= 1/16:
//     J0(c + t) ~= p[0] + p[1] t + ... + p[7] t^7.
// The rows overlap: a row is valid up to twice its own half-spacing, which lets a whole stage of visibilities
// (sorted by baseline, so their arguments a_i j_k for one mode k are nearly equal) use ONE row per mode chosen
// from the middle of the stage, without per-visibility row selection.  Kernels accept |t| <= FB_J0_ACCEPT.
//
// Construction: Taylor coefficients of J0 about c to degree 15 in long double (power series of J0 for c <= 4,
// the three-term recurrence that follows from the Bessel equation x y'' + y' + x y = 0 above, seeded with
// glibc's 80-bit j0l / j1l), then Chebyshev interpolation of that polynomial on [-1/16, 1/16] truncated to
// degree 7 (near-minimax; truncation error <= |J0^(8)| (1/16)^8 / (8! 2^7) = 4.5e-17 |J0^(8)|, i.e. below
// 7e-18 for x > 30 where |J0^(n)| <= 0.15) and conversion back to monomials.
#pragma once
#include <cmath>
#include <vector>

constexpr int FB_J0_ROWLEN = 8;              // coefficients per row (degree 7)
constexpr double FB_J0_H = 0.0625;           // row spacing
constexpr double FB_J0_INVH = 16.0;
constexpr double FB_J0_ACCEPT = 0.0625;      // a row may be used for |t| up to this
constexpr int FB_J0_TAYLOR = 16;             // Taylor terms used for the construction

inline void fb_j0_taylor(long double c, long double *a /*[FB_J0_TAYLOR]*/)
{
    const int TD = FB_J0_TAYLOR;
    if (c <= 4.0L) {
        // J0(c + t) = sum_j (-1/4)^j (c + t)^(2j) / (j!)^2  ->  a_k = sum_j (-1/4)^j / (j!)^2 C(2j, k) c^(2j - k)
        for (int k = 0; k < TD; k++) a[k] = 0.0L;
        long double coef = 1.0L;                       // (-1/4)^j / (j!)^2
        for (int j = 0; j < 60; j++) {
            // binom(2j, k) c^(2j-k) for k = 0 .. min(2j, TD-1)
            long double term = 1.0L;                   // C(2j, k) c^(2j - k), built from k = 2j downwards
            // start at k = 2j: C = 1, c^0 = 1
            for (int k = 2 * j; k >= 0; k--) {
                if (k < TD) a[k] += coef * term;
                // C(2j, k-1) c^(2j-k+1) = C(2j, k) * k / (2j - k + 1) * c
                if (k > 0) term = term * (long double)k / (long double)(2 * j - k + 1) * c;
            }
            coef *= -0.25L / ((long double)(j + 1) * (long double)(j + 1));
        }
        return;
    }
    a[0] = j0l(c);
    a[1] = -j1l(c);
    for (int k = 0; k + 2 < TD; k++) {
        const long double prev = k == 0 ? 0.0L : a[k - 1];
        a[k + 2] = -(((long double)(k + 1) * (k + 1)) * a[k + 1] + c * a[k] + prev) / (c * (long double)((k + 2) * (k + 1)));
    }
}

inline void fb_j0_row(int m, double *out /*[FB_J0_ROWLEN]*/)
{
    const long double c = (long double)m * (long double)FB_J0_H;
    long double a[FB_J0_TAYLOR];
    fb_j0_taylor(c, a);
    const long double r = (long double)FB_J0_ACCEPT;
    const long double PI = 3.14159265358979323846264338327950288L;
    constexpr int M = 32;
    long double ch[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < M; j++) {
        const long double th = PI * ((long double)j + 0.5L) / (long double)M;
        const long double t = r * cosl(th);
        long double f = a[FB_J0_TAYLOR - 1];
        for (int k = FB_J0_TAYLOR - 2; k >= 0; k--) f = f * t + a[k];
        for (int k = 0; k < 8; k++) ch[k] += f * cosl((long double)k * th);
    }
    for (int k = 0; k < 8; k++) ch[k] *= 2.0L / (long double)M;
    ch[0] *= 0.5L;
    long double b[8];
    b[0] = ch[0] - ch[2] + ch[4] - ch[6];
    b[1] = ch[1] - 3 * ch[3] + 5 * ch[5] - 7 * ch[7];
    b[2] = 2 * ch[2] - 8 * ch[4] + 18 * ch[6];
    b[3] = 4 * ch[3] - 20 * ch[5] + 56 * ch[7];
    b[4] = 8 * ch[4] - 48 * ch[6];
    b[5] = 16 * ch[5] - 112 * ch[7];
    b[6] = 32 * ch[6];
    b[7] = 64 * ch[7];
    long double scale = 1.0L;                           // (1 / r)^k, exact powers of two
    for (int k = 0; k < 8; k++) {
        out[k] = (double)(b[k] * scale);
        scale /= r;
    }
}

inline int fb_j0_rows_for(double x_max) { return (int)std::ceil(x_max * FB_J0_INVH) + 3; }

inline void fb_j0_build(double x_max, std::vector<double> &tab)
{
    const int rows = fb_j0_rows_for(x_max);
    tab.resize((size_t)rows * FB_J0_ROWLEN);
    for (int m = 0; m < rows; m++) fb_j0_row(m, &tab[(size_t)m * FB_J0_ROWLEN]);
}
