// K3: fused design-matrix + Gram kernel.
//
// Replaces the reference's chunk loop (frank/statistical_models.py:192-214)
//     X   = H(q_chunk)                      DHT.coefficients, frank/hankel.py:187-204
//     wXT = X.T * w ;  M += wXT @ X ;  j += wXT @ V
// with one persistent kernel in which the design-matrix tile never leaves the SM:
//   * each CTA owns a block of the (N+1)x(N+1) symmetric matrix  S = G^T G,
//         G[i, k] = sqrt(w_i) J0(a_i j_k)  (k < N),   G[i, N] = sqrt(w_i) Re V_i,
//     so that M = diag(c) S[:N,:N] diag(c), j = diag(c) S[:N, N], with c_k = norm * scale_factor_k * scale;
//   * per tile of 64 visibilities all 16 warps first evaluate J0 for the block's row and column panels
//     into shared memory G[mode][vis] (FP64 piecewise Taylor table, |err| <= 0.5 ulp + 1e-17), then all
//     warps run mma.sync.m8n8k4.f64 (DMMA) over the tile with accumulators in registers;
//   * off-diagonal blocks are 19x19 tiles of 8x8 (rectangular 5x5 tiles per warp); the two diagonal
//     blocks of a panel pair are computed as skewed strips (row r, offset d -> column (r+d) mod n) so
//     that only the upper triangle is executed and every warp still owns a dense 5x5 register block;
//   * work items (block, visibility chunk) write partial blocks; a second kernel sums the chunks in a fixed
//     order (deterministic), applies c_k c_l and mirrors the triangle.
#include "fb_common.cuh"

#include <cmath>
#include <cstdlib>

namespace {

struct GramArgs {
    const double *a, *sw, *swV, *kz;
    long long n_tiles;
    const double *jk;
    const double2 *tab;
    int tab_rows;
    int N, ntypes, C;
    const FbGramType *types;
    const double *H2;
    double *partial;
    int debug_mode;   // 0 normal | 1 skip the DMMA phase | 2 skip J0 evaluation (profiling aid, FB_GRAM_DEBUG)
};

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

// J0(x), x >= 0: row m = round(4x) holds the Taylor coefficients about m/4, |t| <= 1/8.
__device__ __forceinline__ double j0_tab(double x, const double2 *__restrict__ tab, int last_row)
{
    const double MAGIC = 6755399441055744.0;   // 1.5 * 2^52: adding it rounds to the nearest integer
    double s = fma(x, 4.0, MAGIC);
    int m = min(__double2loint(s), last_row);
    double t = fma(s - MAGIC, -0.25, x);       // exact
    const double2 *row = tab + (size_t)m * (FB_J0_ROWLEN / 2);
    double2 c01 = __ldg(row), c23 = __ldg(row + 1), c45 = __ldg(row + 2), c67 = __ldg(row + 3), c89 = __ldg(row + 4);
    double y = fma(c89.y, t, c89.x);
    y = fma(y, t, c67.y);
    y = fma(y, t, c67.x);
    y = fma(y, t, c45.y);
    y = fma(y, t, c45.x);
    y = fma(y, t, c23.y);
    y = fma(y, t, c23.x);
    y = fma(y, t, c01.y);
    y = fma(y, t, c01.x);
    return y;
}

__device__ __forceinline__ int clamp5(int x) { return x < 0 ? 0 : (x > 5 ? 5 : x); }

constexpr int GT = FB_TV;              // visibilities per tile (64)
constexpr int GLD = FB_LDV;            // 68 = 4 mod 16: conflict-free m8n8k4 fragments
constexpr int GKS = GT / 4;            // k-steps per tile
constexpr int GCOLS = 2 * FB_PCOLS;    // columns held per tile (panel A | panel B)

// shared-memory carve-up (doubles unless noted)
constexpr int SM_G = 0;                                   // [GCOLS][GLD]        design-matrix tile
constexpr int SM_JK = SM_G + GCOLS * GLD;                 // [GCOLS]             j_k (>= 0), -1 data column, -2 padding
constexpr int SM_H2 = SM_JK + GCOLS;                      // [GCOLS]             debris H2_k
constexpr int SM_ROW = SM_H2 + GCOLS;                     // [GCOLS][10]         staged J0 table rows
constexpr int SM_ROWM = SM_ROW + GCOLS * FB_J0_ROWLEN;    // [GCOLS] int         staged row index
constexpr int SM_CFG = SM_ROWM + GCOLS / 2;                   // [16 warps][8] int  per-warp register-block description
constexpr int SM_DOUBLES = SM_CFG + 16 * 8 / 2;
constexpr size_t GRAM_SMEM_BYTES = sizeof(double) * SM_DOUBLES;

// rectangular register block: acc[r][c] += G_A[r0+r]^T G_B[c0+c] over one tile
template <int NR, int NC>
__device__ __forceinline__ void mma_off(double (&acc)[5][5][2], const double *__restrict__ ap, const double *__restrict__ bp)
{
#pragma unroll 2
    for (int ks = 0; ks < GKS; ks++) {
        double af[NR];
#pragma unroll
        for (int r = 0; r < NR; r++) af[r] = ap[r * 8 * GLD + ks * 4];
#pragma unroll
        for (int c = 0; c < NC; c++) {
            const double b = bp[c * 8 * GLD + ks * 4];
#pragma unroll
            for (int r = 0; r < NR; r++) dmma(acc[r][c], af[r], b);
        }
    }
}

// skewed strip of a triangle: acc[r][d] += G[r0+r]^T G[(r0+r + d0+d) mod n] over one tile;
// boff[s] = shared-memory offset of tile column (r0 + d0 + s) mod n
template <int NR, int ND>
__device__ __forceinline__ void mma_diag(double (&acc)[5][5][2], const double *__restrict__ ap, const double *__restrict__ tp,
                                         int cbase, int nmod)
{
#pragma unroll 2
    for (int ks = 0; ks < GKS; ks++) {
        double af[NR];
#pragma unroll
        for (int r = 0; r < NR; r++) af[r] = ap[r * 8 * GLD + ks * 4];
#pragma unroll
        for (int s = 0; s < NR + ND - 1; s++) {
            int ct = cbase + s;
            ct = ct >= nmod ? ct - nmod : ct;
            const double b = tp[ct * 8 * GLD + ks * 4];
#pragma unroll
            for (int r = 0; r < NR; r++) {
                const int d = s - r;
                if (d >= 0 && d < ND) dmma(acc[r][d], af[r], b);
            }
        }
    }
}

template <int NR>
__device__ __forceinline__ void mma_off_nc(double (&acc)[5][5][2], const double *ap, const double *bp, int nc)
{
    switch (nc) {
        case 5: mma_off<NR, 5>(acc, ap, bp); break;
        case 4: mma_off<NR, 4>(acc, ap, bp); break;
        case 3: mma_off<NR, 3>(acc, ap, bp); break;
        case 2: mma_off<NR, 2>(acc, ap, bp); break;
        case 1: mma_off<NR, 1>(acc, ap, bp); break;
        default: break;
    }
}
template <int NR>
__device__ __forceinline__ void mma_diag_nd(double (&acc)[5][5][2], const double *ap, const double *tp, int cbase, int nmod, int nd)
{
    switch (nd) {
        case 5: mma_diag<NR, 5>(acc, ap, tp, cbase, nmod); break;
        case 4: mma_diag<NR, 4>(acc, ap, tp, cbase, nmod); break;
        case 3: mma_diag<NR, 3>(acc, ap, tp, cbase, nmod); break;
        case 2: mma_diag<NR, 2>(acc, ap, tp, cbase, nmod); break;
        case 1: mma_diag<NR, 1>(acc, ap, tp, cbase, nmod); break;
        default: break;
    }
}

// Evaluate the staged polynomial for two visibilities of one column (shared coefficient loads).
__device__ __forceinline__ void horner2(const double2 *__restrict__ rb, double t0, double t1, double &g0, double &g1)
{
    double2 c = rb[4];
    g0 = fma(c.y, t0, c.x);           g1 = fma(c.y, t1, c.x);
    c = rb[3];
    g0 = fma(g0, t0, c.y);            g1 = fma(g1, t1, c.y);
    g0 = fma(g0, t0, c.x);            g1 = fma(g1, t1, c.x);
    c = rb[2];
    g0 = fma(g0, t0, c.y);            g1 = fma(g1, t1, c.y);
    g0 = fma(g0, t0, c.x);            g1 = fma(g1, t1, c.x);
    c = rb[1];
    g0 = fma(g0, t0, c.y);            g1 = fma(g1, t1, c.y);
    g0 = fma(g0, t0, c.x);            g1 = fma(g1, t1, c.x);
    c = rb[0];
    g0 = fma(g0, t0, c.y);            g1 = fma(g1, t1, c.y);
    g0 = fma(g0, t0, c.x);            g1 = fma(g1, t1, c.x);
}

// The persistent J0 + Gram kernel: per tile of 64 visibilities
//   (1) 304 threads stage, per column, the J0 table row the tile's first visibility needs (the visibilities
//       are sorted by baseline, so nearly always every lane needs that same row);
//   (2) all 16 warps evaluate J0 for their columns (two visibilities per lane share the coefficient loads)
//       into G[column][vis] in shared memory;
//   (3) all 16 warps run the DMMAs of their register block over the tile.
template <bool DEBRIS>
__global__ void __launch_bounds__(FB_GRAM_THREADS, 1) k_gram(GramArgs p)
{
    extern __shared__ double smem[];
    double *G = smem + SM_G;
    double *col_jk = smem + SM_JK;
    double *col_h2 = smem + SM_H2;
    double *rowbuf = smem + SM_ROW;
    int *rowm = reinterpret_cast<int *>(smem + SM_ROWM);
    int *cfg = reinterpret_cast<int *>(smem + SM_CFG);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int last_row = p.tab_rows - 1;
    const double MAGIC = 6755399441055744.0;

    for (int item = blockIdx.x; item < p.ntypes * p.C; item += gridDim.x) {
        const int type = item % p.ntypes, chunk = item / p.ntypes;
        const FbGramType ty = p.types[type];
        const int t0 = (int)((p.n_tiles * chunk) / p.C), t1 = (int)((p.n_tiles * (chunk + 1)) / p.C);
        const int ncolA = ty.a_nt * 8, ncol = (ty.a_nt + ty.b_nt) * 8;

        __syncthreads();
        for (int lc = tid; lc < ncol; lc += FB_GRAM_THREADS) {
            const bool inA = lc < ncolA;
            const int g = inA ? ty.a_t0 * 8 + lc : ty.b_t0 * 8 + (lc - ncolA);
            const int sc = inA ? lc : FB_PCOLS + (lc - ncolA);
            col_jk[sc] = g < p.N ? p.jk[g] : (g == p.N ? -1.0 : -2.0);
            if (DEBRIS) col_h2[sc] = g < p.N ? p.H2[g] : 0.0;
        }

        double acc[5][5][2];
#pragma unroll
        for (int r = 0; r < 5; r++)
#pragma unroll
            for (int c = 0; c < 5; c++) acc[r][c][0] = acc[r][c][1] = 0.0;

        // this warp's register block (all warp-uniform).  Kept in shared memory and re-read at the start of
        // every DMMA phase so that it does not occupy registers (next to 100 accumulator registers) while
        // the J0 loop runs.
        const int smsp = warp & 3, slot = warp >> 2;
        if (lane == 0) {
            int r0, c0, nr, nc, nmod = 1, base_cols = 0;
            if (ty.kind == FB_KIND_OFF) {
                r0 = 5 * slot;
                c0 = 5 * ((slot + smsp) & 3);
                nr = clamp5(ty.a_nt - r0);
                nc = clamp5(ty.b_nt - c0);
            } else {
                const int tri = slot >> 1;
                nmod = tri ? ty.b_nt : ty.a_nt;
                base_cols = tri ? FB_PCOLS : 0;
                r0 = 5 * ((smsp + slot) & 3);
                c0 = 5 * (slot & 1);                       // first skew offset d
                nr = clamp5(nmod - r0);
                nc = nmod > 0 ? clamp5(nmod / 2 + 1 - c0) : 0;
                if (nmod == 0) nmod = 1;
            }
            if (nr == 0 || nc == 0) { nr = 0; nc = 0; }
            int *cf = cfg + warp * 8;
            cf[0] = nr; cf[1] = nc; cf[2] = r0; cf[3] = c0; cf[4] = nmod;
            cf[5] = (base_cols + r0 * 8) * GLD;                                                      // A fragment base
            cf[6] = ty.kind == FB_KIND_OFF ? (FB_PCOLS + c0 * 8) * GLD : base_cols * GLD;            // B fragment base
            cf[7] = (r0 + c0) % nmod;
        }
        const int kind = ty.kind;
        const int ncolB = ncol - ncolA;
        // staging role: thread tid < ncol stages the row of shared column sc_stage
        const int sc_stage = tid < ncolA ? tid : FB_PCOLS + (tid - ncolA);

        for (int tile = t0; tile < t1; ++tile) {
            const size_t v0 = (size_t)tile * GT;
            // ---- (1) stage table rows -------------------------------------------------------------
            if (p.debug_mode != 2 || tile == t0) {
                if (tid < ncol) {
                    const double jk = col_jk[sc_stage];
                    int m = 0;
                    if (jk >= 0.0) m = min(__double2loint(fma(__dmul_rn(p.a[v0], jk), 4.0, MAGIC)), last_row);
                    rowm[sc_stage] = m;
                    const double2 *row = p.tab + (size_t)m * (FB_J0_ROWLEN / 2);
                    double2 *dst = reinterpret_cast<double2 *>(rowbuf + sc_stage * FB_J0_ROWLEN);
#pragma unroll
                    for (int k = 0; k < FB_J0_ROWLEN / 2; k++) dst[k] = __ldg(row + k);
                }
            }
            __syncthreads();          // rows staged; every warp is done with the previous tile's DMMAs
            // ---- (2) J0 evaluation ------------------------------------------------------------------
            if (p.debug_mode != 2 || tile == t0) {
                const double a0 = p.a[v0 + lane], a1 = p.a[v0 + lane + 32];
                const double w0 = p.sw[v0 + lane], w1 = p.sw[v0 + lane + 32];
                double k0 = 0.0, k1 = 0.0;
                if (DEBRIS) {
                    k0 = p.kz[v0 + lane]; k1 = p.kz[v0 + lane + 32];
                    k0 = -k0 * k0; k1 = -k1 * k1;
                }
                // panel A columns then panel B columns; the start of the second sweep is rotated by 8 warps
                // so that an odd column count per 16 warps balances out
#pragma unroll 1
                for (int half = 0; half < 2; half++) {
                    const int nh = half ? ncolB : ncolA;
                    const int base = half ? FB_PCOLS : 0;
#pragma unroll 1
                    for (int lc = half ? ((warp + 8) & 15) : warp; lc < nh; lc += FB_GRAM_THREADS / 32) {
                        const int sc = base + lc;
                        const double jk = col_jk[sc];
                        double g0, g1;
                        if (jk >= 0.0) {
                            const double x0 = __dmul_rn(a0, jk), x1 = __dmul_rn(a1, jk);
                            const double s0 = fma(x0, 4.0, MAGIC), s1 = fma(x1, 4.0, MAGIC);
                            const int mref = rowm[sc];
                            const bool same = (min(__double2loint(s0), last_row) == mref) & (min(__double2loint(s1), last_row) == mref);
                            if (__all_sync(0xffffffffu, same)) {
                                horner2(reinterpret_cast<const double2 *>(rowbuf + sc * FB_J0_ROWLEN),
                                        fma(s0 - MAGIC, -0.25, x0), fma(s1 - MAGIC, -0.25, x1), g0, g1);
                            } else {
                                g0 = j0_tab(x0, p.tab, last_row);
                                g1 = j0_tab(x1, p.tab, last_row);
                            }
                            if (DEBRIS) {
                                const double h2 = col_h2[sc];
                                g0 *= exp(k0 * h2);
                                g1 *= exp(k1 * h2);
                            }
                            g0 *= w0;
                            g1 *= w1;
                        } else if (jk == -1.0) {
                            g0 = p.swV[v0 + lane];
                            g1 = p.swV[v0 + lane + 32];
                        } else {
                            g0 = 0.0;
                            g1 = 0.0;
                        }
                        G[sc * GLD + lane] = g0;
                        G[sc * GLD + lane + 32] = g1;
                    }
                }
            }
            __syncthreads();
            // ---- (3) DMMA over the tile -----------------------------------------------------------
            if (p.debug_mode != 1) {
                const int *cf = cfg + warp * 8;
                const int nr = cf[0], nc = cf[1], nmod = cf[4], cbase = cf[7];
                const int frag = (lane >> 2) * GLD + (lane & 3);
                const double *ap = G + cf[5] + frag;
                const double *bp = G + cf[6] + frag;
                if (nr > 0) {
                    if (kind == FB_KIND_OFF) {
                        switch (nr) {
                            case 5: mma_off_nc<5>(acc, ap, bp, nc); break;
                            case 4: mma_off_nc<4>(acc, ap, bp, nc); break;
                            case 3: mma_off_nc<3>(acc, ap, bp, nc); break;
                            case 2: mma_off_nc<2>(acc, ap, bp, nc); break;
                            default: mma_off_nc<1>(acc, ap, bp, nc); break;
                        }
                    } else {
                        switch (nr) {
                            case 5: mma_diag_nd<5>(acc, ap, bp, cbase, nmod, nc); break;
                            case 4: mma_diag_nd<4>(acc, ap, bp, cbase, nmod, nc); break;
                            case 3: mma_diag_nd<3>(acc, ap, bp, cbase, nmod, nc); break;
                            case 2: mma_diag_nd<2>(acc, ap, bp, cbase, nmod, nc); break;
                            default: mma_diag_nd<1>(acc, ap, bp, cbase, nmod, nc); break;
                        }
                    }
                }
            }
        }

        // ---------------- write the partial block ------------------------------------------------
        double *out = p.partial + (size_t)item * FB_PSZ;
        const int nr = cfg[warp * 8 + 0], nc = cfg[warp * 8 + 1], r0 = cfg[warp * 8 + 2], c0 = cfg[warp * 8 + 3];
        if (kind == FB_KIND_OFF) {
#pragma unroll
            for (int r = 0; r < 5; r++)
#pragma unroll
                for (int c = 0; c < 5; c++)
                    if (r < nr && c < nc)
                        *reinterpret_cast<double2 *>(out + ((size_t)((r0 + r) * FB_PT + (c0 + c)) * 64 + lane * 2)) =
                            make_double2(acc[r][c][0], acc[r][c][1]);
        } else {
            const int tri = slot >> 1;
#pragma unroll
            for (int r = 0; r < 5; r++)
#pragma unroll
                for (int d = 0; d < 5; d++)
                    if (r < nr && d < nc)
                        *reinterpret_cast<double2 *>(out + ((size_t)(tri * FB_PT * FB_DH + (r0 + r) * FB_DH + (c0 + d)) * 64 + lane * 2)) =
                            make_double2(acc[r][d][0], acc[r][d][1]);
        }
    }
}

// Sum the chunk partials in a fixed order, scale, mirror.  One 64-thread block per upper tile pair.
__global__ void __launch_bounds__(64)
k_gram_finalize(int N, int NT, int P, int ntypes, int C, long long n_tiles, const int *__restrict__ tile_panel,
                const int *__restrict__ panel_t0, const int *__restrict__ panel_nt, const int *__restrict__ pair_code,
                const double *__restrict__ partial, const double *__restrict__ ck, double scale,
                double *__restrict__ M, double *__restrict__ jvec)
{
    // decode the upper-triangular tile pair (tr <= tc) from blockIdx.x
    int rem = blockIdx.x, tr = 0;
    while (rem >= NT - tr) { rem -= NT - tr; tr++; }
    const int tc = tr + rem;
    const int i = threadIdx.x >> 3, jx = threadIdx.x & 7;
    const int row = tr * 8 + i, col = tc * 8 + jx;
    if (tr == tc && i > jx) return;
    const int pa = tile_panel[tr], pb = tile_panel[tc];
    int type, idx;
    if (pa < pb) {
        type = pair_code[pa * P + pb];
        const int lr = tr - panel_t0[pa], lc = tc - panel_t0[pb];
        idx = (lr * FB_PT + lc) * 64 + (i * 4 + (jx >> 1)) * 2 + (jx & 1);
    } else {
        const int code = pair_code[pa * P + pa];
        type = code >> 1;
        const int tri = code & 1, n = panel_nt[pa], D = n / 2 + 1;
        const int lr = tr - panel_t0[pa], lc = tc - panel_t0[pa], d = lc - lr;
        if (d < D)
            idx = (tri * FB_PT * FB_DH + lr * FB_DH + d) * 64 + (i * 4 + (jx >> 1)) * 2 + (jx & 1);
        else   // stored as the transposed tile (row tile lc, offset n - d)
            idx = (tri * FB_PT * FB_DH + lc * FB_DH + (n - d)) * 64 + (jx * 4 + (i >> 1)) * 2 + (i & 1);
    }
    double s = 0.0;
    for (int c = 0; c < C; c++) {
        const long long t0 = (n_tiles * c) / C, t1 = (n_tiles * (c + 1)) / C;
        if (t1 > t0) s += partial[(size_t)(c * ntypes + type) * FB_PSZ + idx];
    }
    if (col < N) {           // row <= col < N
        const double val = ((ck[row] * scale) * (ck[col] * scale)) * s;
        M[(size_t)row * N + col] = val;
        M[(size_t)col * N + row] = val;
    } else if (col == N && row < N) {
        jvec[row] = (ck[row] * scale) * s;
    }
}

__global__ void k_j0_debug(int64_t n, const double *__restrict__ x, double *__restrict__ out, const double2 *tab, int rows)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = j0_tab(x[i], tab, rows - 1);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
int fb_build_j0_table(fb_ctx *ctx, double x_max)
{
    // Rows centred on m/4; Taylor coefficients from the Bessel ODE  x y'' + y' + x y = 0:
    //   a_{k+2} = -[(k+1)^2 a_{k+1} + c a_k + a_{k-1}] / (c (k+2)(k+1)),  a_0 = J0(c), a_1 = -J1(c),
    // seeded with glibc's 80-bit j0l / j1l and run in long double, then rounded to double.
    const int rows = (int)std::ceil(x_max * 4.0) + 3;
    std::vector<double> tab((size_t)rows * FB_J0_ROWLEN);
    for (int m = 0; m < rows; m++) {
        long double a[FB_J0_ROWLEN + 2];
        if (m == 0) {
            // J0(t) = sum_k (-1/4)^k t^(2k) / (k!)^2
            long double term = 1.0L;
            for (int k = 0; k < FB_J0_ROWLEN; k++) a[k] = 0.0L;
            for (int k = 0; 2 * k < FB_J0_ROWLEN; k++) {
                a[2 * k] = term;
                term *= -0.25L / ((long double)(k + 1) * (long double)(k + 1));
            }
        } else {
            const long double c = 0.25L * (long double)m;
            a[0] = j0l(c);
            a[1] = -j1l(c);
            for (int k = 0; k + 2 < FB_J0_ROWLEN; k++) {
                const long double prev = k == 0 ? 0.0L : a[k - 1];
                a[k + 2] = -(((long double)(k + 1) * (k + 1)) * a[k + 1] + c * a[k] + prev) /
                           (c * (long double)((k + 2) * (k + 1)));
            }
        }
        for (int k = 0; k < FB_J0_ROWLEN; k++) tab[(size_t)m * FB_J0_ROWLEN + k] = (double)a[k];
    }
    if (ctx->d_tab) FB_CUDA(cudaFree(ctx->d_tab));
    ctx->d_tab = nullptr;
    FB_CUDA(cudaMalloc(&ctx->d_tab, tab.size() * sizeof(double)));
    FB_CUDA(cudaMemcpy(ctx->d_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
    ctx->tab_rows = rows;
    return 0;
}

int fb_launch_gram(fb_ctx *ctx, int64_t n, int vis_model, double model_scale, double *dev_M, double *dev_j)
{
    const long long n_tiles = (n + FB_TV - 1) / FB_TV;
    // chunks per type: fill the machine, at least one tile per chunk where possible
    int C = (2 * ctx->num_sms) / ctx->ntypes;
    if (ctx->ntypes * C > ctx->num_sms && (ctx->num_sms / ctx->ntypes) >= 1) C = ctx->num_sms / ctx->ntypes;
    if (C < 1) C = 1;
    if ((long long)C > n_tiles) C = (int)(n_tiles > 0 ? n_tiles : 1);
    const size_t need = (size_t)ctx->ntypes * C * FB_PSZ;
    if (need > ctx->partial_cap) {
        if (ctx->d_partial) FB_CUDA(cudaFree(ctx->d_partial));
        ctx->d_partial = nullptr;
        FB_CUDA(cudaMalloc(&ctx->d_partial, need * sizeof(double)));
        ctx->partial_cap = need;
    }
    GramArgs args;
    args.a = ctx->d_a; args.sw = ctx->d_sw; args.swV = ctx->d_swV; args.kz = ctx->d_kz;
    args.n_tiles = n_tiles;
    args.jk = ctx->d_jk; args.tab = ctx->d_tab; args.tab_rows = ctx->tab_rows;
    args.N = ctx->N; args.ntypes = ctx->ntypes; args.C = C;
    args.types = ctx->d_types; args.H2 = ctx->d_H2; args.partial = ctx->d_partial;
    {
        const char *dbg = getenv("FB_GRAM_DEBUG");
        args.debug_mode = dbg ? atoi(dbg) : 0;
    }

    const size_t smem = GRAM_SMEM_BYTES;
    int grid = ctx->ntypes * C;
    if (grid > ctx->num_sms) grid = ctx->num_sms;
    if (n_tiles > 0) {
        if (vis_model == FB_MODEL_DEBRIS) {
            FB_CUDA(cudaFuncSetAttribute(k_gram<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_gram<true><<<grid, FB_GRAM_THREADS, smem, ctx->stream>>>(args);
        } else {
            FB_CUDA(cudaFuncSetAttribute(k_gram<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_gram<false><<<grid, FB_GRAM_THREADS, smem, ctx->stream>>>(args);
        }
        FB_CUDA(cudaGetLastError());
    }
    FB_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));
    const int npairs = ctx->NT * (ctx->NT + 1) / 2;
    k_gram_finalize<<<npairs, 64, 0, ctx->stream>>>(ctx->N, ctx->NT, ctx->P, ctx->ntypes, C, n_tiles, ctx->d_tile_panel,
                                                    ctx->d_panel_t0, ctx->d_panel_nt, ctx->d_pair_code, ctx->d_partial,
                                                    ctx->d_ck, model_scale, dev_M, dev_j);
    FB_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int fb_debug_j0(fb_ctx *ctx, int64_t n, const double *host_x, double *host_out)
{
    if (!ctx || !ctx->d_tab) return -1;
    double *dx = nullptr, *dout = nullptr;
    FB_CUDA(cudaSetDevice(ctx->device));
    FB_CUDA(cudaMalloc(&dx, sizeof(double) * n));
    FB_CUDA(cudaMalloc(&dout, sizeof(double) * n));
    FB_CUDA(cudaMemcpy(dx, host_x, sizeof(double) * n, cudaMemcpyHostToDevice));
    k_j0_debug<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, dx, dout, ctx->d_tab, ctx->tab_rows);
    FB_CUDA(cudaGetLastError());
    FB_CUDA(cudaStreamSynchronize(ctx->stream));
    FB_CUDA(cudaMemcpy(host_out, dout, sizeof(double) * n, cudaMemcpyDeviceToHost));
    cudaFree(dx);
    cudaFree(dout);
    return 0;
}
