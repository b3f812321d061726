// K3: fused design-matrix + Gram kernel.
//
// Replaces the reference's chunk loop (frank/statistical_models.py:192-214)
//     X   = H(q_chunk)                      DHT.coefficients, frank/hankel.py:187-204
//     wXT = X.T * w ;  M += wXT @ X ;  j += wXT @ V
// with one persistent kernel in which the design-matrix tile never leaves the SM:
//   * the symmetric (N+1)x(N+1) matrix  S = G^T G,
//         G[i, k] = sqrt(w_i) J0(a_i j_k)  (k < N),   G[i, N] = sqrt(w_i) Re V_i,
//     gives M = diag(c) S[:N,:N] diag(c), j = diag(c) S[:N, N], with c_k = norm * scale_factor_k * scale;
//   * a CTA owns one block of S (work-item types OFF / DIAG, fb_common.cuh) for a chunk of the visibilities, with
//     its accumulators in registers (<= 16 m8n8 tiles = 64 registers per thread, 16 warps);
//   * per tile of 64 visibilities, two phases separated by barriers (the FP64 pipe serves DMMA and DFMA alike and
//     starves a warp that issues DFMAs while others stream DMMAs -- profiles/r01_gram_ncu_summary.txt -- so the
//     phases are not overlapped):
//       (1) all warps evaluate J0 for the block's columns into shared memory G[mode][vis]: a lane owns one column
//           (its polynomial row in registers), a warp a range of visibilities, four Horner chains per lane.  The
//           visibilities are sorted by baseline, so ONE row of the J0 table per (column, tile) serves all of them
//           (rows overlap, fb_j0_table.h); rows are staged by cp.async one tile ahead; visibilities outside a row's
//           validity window are redone through a per-visibility gather (rare);
//       (2) all warps run mma.sync.m8n8k4.f64 (DMMA) over the tile;
//   * diagonal blocks are executed as skewed strips (row r, offset d -> column (r+d) mod n) so that only the
//     upper triangle is computed while every warp still owns a dense register block;
//   * work items (block, visibility chunk) write partial blocks; a second kernel sums the chunks in a fixed
//     order (deterministic), applies c_k c_l and mirrors the triangle.
#include "fb_common.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <queue>

namespace {

struct GramArgs {
    const double *a, *sw, *swV, *kz;
    const double2 *arange;    // per tile: (min a, max a)
    const int *seg;           // channel segments of the lane (fb_sort.cu): start[] | pad[]
    int chan;                 // channel this launch accumulates
    const int *status;        // status bits of the call: a flagged call skips the work
    const double *jk;
    const double2 *tab;
    int tab_rows;
    int N;
    const FbGramType *types;
    const int *cta_off;   // [grid + 1] item range of each CTA
    const int *items;     // [3 * n_items] (type, chunk, partial slot)
    const int *type_tab;  // [2 * ntypes] (chunks, first slot)
    const double *H2;
    double *partial;
    long long *prof;      // optional (FB_GRAM_PROF=1) [grid][2]: clocks thread 0 spent in the J0 / DMMA phases; results unaffected
};

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

constexpr double J0_MAGIC = 6755399441055744.0;     // 1.5 * 2^52: adding it rounds to the nearest integer

// J0(x), x >= 0, from the row nearest to x (|t| <= 1/32).  Generic (gather) path.
__device__ __forceinline__ double j0_tab(double x, const double2 *__restrict__ tab, int last_row)
{
    double s = fma(x, FB_J0_INVH, J0_MAGIC);
    int m = min(__double2loint(s), last_row);
    double t = fma((double)m, -FB_J0_H, x);       // exact
    const double2 *row = tab + (size_t)m * (FB_J0_ROWLEN / 2);
    double2 c01 = __ldg(row), c23 = __ldg(row + 1), c45 = __ldg(row + 2), c67 = __ldg(row + 3);
    double y = fma(c67.y, t, c67.x);
    y = fma(y, t, c45.y);
    y = fma(y, t, c45.x);
    y = fma(y, t, c23.y);
    y = fma(y, t, c23.x);
    y = fma(y, t, c01.y);
    y = fma(y, t, c01.x);
    return y;
}

// shared-state-space accessors (32-bit addresses: no generic-window arithmetic inside the hot loops)
__device__ __forceinline__ double lds_f64(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ double2 lds_v2f64(uint32_t addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f64(uint32_t addr, double v)
{
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void sts_v2f64(uint32_t addr, double x, double y)
{
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(x), "d"(y) : "memory");
}
constexpr int GS = FB_TV;              // visibilities per tile (64)
constexpr int GLD = FB_LDV;            // 68 = 4 mod 16: conflict-free m8n8k4 fragments
constexpr int GKS = GS / 4;            // k-steps per tile
constexpr int GCOLS = FB_GCOLS;        // columns held per stage (rows' columns | columns' columns)
constexpr int NW = FB_GRAM_THREADS / 32;
constexpr int ROWB = FB_J0_ROWLEN * 8; // bytes per table row (64)
static_assert(FB_GRAM_THREADS == 2 * FB_GCOLS && FB_J0_ROWLEN == 8, "staging assigns thread 256 + c to column c and four threads to a row");

// shared-memory carve-up (bytes)
constexpr int SMB_G = 0;                                   // [GCOLS][GLD] doubles      design-matrix tile
constexpr int SMB_ROW = SMB_G + GCOLS * GLD * 8;           // [2][4][GCOLS] double2     staged J0 table rows (coefficient pair major)
constexpr int SMB_CEN = SMB_ROW + 2 * GCOLS * ROWB;        // [2][GCOLS] double2        (centre of the staged row, +j_k row serves the tile | -j_k gather | -0 special column)
constexpr int SMB_JK = SMB_CEN + 2 * GCOLS * 16;           // [GCOLS] ints (+ pad)      table row chosen for the tile being staged
constexpr int SMB_H2 = SMB_JK + GCOLS * 8;                 // [GCOLS] doubles           debris H2_k
constexpr int SMB_VIS = SMB_H2 + GCOLS * 8;                // [2][GS][4] doubles        (a, sqrt w, kz, sqrt w Re V) per visibility of a tile
constexpr int SMB_AR = SMB_VIS + 2 * GS * 32;               // [2] double2               (min a, max a) of the tiles being staged
constexpr int SMB_SW = SMB_AR + 32;                        // [2][GS] doubles           sqrt w, compact (conflict-free re-reads of the half-warp sweep)
constexpr int GRAM_SMEM_BYTES = SMB_SW + 2 * GS * 8;

__host__ __device__ __forceinline__ int split4_size(int n, int i) { return n / 4 + (i < n % 4 ? 1 : 0); }
__host__ __device__ __forceinline__ int split4_start(int n, int i) { return i * (n / 4) + (i < n % 4 ? i : n % 4); }

// ---------------------------------------------------------------------------------------------
// J0 evaluation
// ---------------------------------------------------------------------------------------------
// exp(x) for x <= 0, branch free: x = k ln2 + r, |r| <= ln2 / 2, degree-13 Taylor polynomial, 2^k through the exponent
// field; results below the normal range flush to zero (the reference multiplies by np.exp, which denormalises
// there: differences < 2.3e-308 in a factor of the design matrix).  Error <= 1 ulp.
__device__ __forceinline__ double exp_neg(double x)
{
    const double t = fma(x, 1.4426950408889634, J0_MAGIC);
    const int k = __double2loint(t);
    const double kf = t - J0_MAGIC;
    double r = fma(kf, -6.93147180369123816490e-01, x);
    r = fma(kf, -1.90821492927058770002e-10, r);
    double q = 1.6059043836821613e-10;                      // 1/13!
    q = fma(q, r, 2.08767569878681e-09);
    q = fma(q, r, 2.505210838544172e-08);
    q = fma(q, r, 2.755731922398589e-07);
    q = fma(q, r, 2.7557319223985893e-06);
    q = fma(q, r, 2.48015873015873e-05);
    q = fma(q, r, 1.984126984126984e-04);
    q = fma(q, r, 1.388888888888889e-03);
    q = fma(q, r, 8.333333333333333e-03);
    q = fma(q, r, 4.1666666666666664e-02);
    q = fma(q, r, 1.6666666666666666e-01);
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    q = fma(q, r, 1.0);
    const int hi = __double2hiint(q) + (k << 20);
    const double y = __hiloint2double(hi, __double2loint(q));
    return k > -1021 ? y : 0.0;
}

// J0 phase of one warp for one tile: columns sc = warp, warp + 16, ... of the block; a lane holds two visibilities
// (lane, lane + 32), so the column's polynomial row is fetched once (broadcast loads from the staged rows) for two
// Horner chains.  The thread that staged the row has verified that it serves the tile's whole range of arguments
// (validity flag); a column that needs more than one row (sparse data, or very large j_k) is redone through the
// per-visibility gather path AFTER the sweep, so that the sweep itself carries no branch (the reconvergence bookkeeping of
// a rare in-loop branch cost 3 % of the kernel).  The data column and the zero padding are not written here.
template <bool DEBRIS>
__device__ __forceinline__ void j0_columns(const GramArgs &p, const uint32_t sbase, const int buf, const int ncol,
                                           const int warp, const int lane, const int last_row)
{
    const uint32_t s_vis = sbase + SMB_VIS + buf * (GS * 32);
    const double2 aw0 = lds_v2f64(s_vis + lane * 32), aw1 = lds_v2f64(s_vis + (lane + 32) * 32);
    double k0 = 0.0, k1 = 0.0;
    if (DEBRIS) {
        k0 = lds_f64(s_vis + lane * 32 + 16); k1 = lds_f64(s_vis + (lane + 32) * 32 + 16);
        k0 = -k0 * k0; k1 = -k1 * k1;                     // -kz^2
    }
    uint32_t a_row = sbase + SMB_ROW + buf * (GCOLS * ROWB) + warp * 16;
    uint32_t a_cen = sbase + SMB_CEN + (buf * GCOLS + warp) * 16;
    uint32_t a_g = sbase + SMB_G + (warp * GLD + lane) * 8;
    bool any_gather = false;
#pragma unroll 1
    for (int sc = warp; sc < ncol; sc += NW) {
        const double2 cv = lds_v2f64(a_cen);                         // (centre, +j_k | -j_k: gather | -0: no store)
        const double2 c67 = lds_v2f64(a_row + 3 * GCOLS * 16), c45 = lds_v2f64(a_row + 2 * GCOLS * 16),
                      c23 = lds_v2f64(a_row + GCOLS * 16), c01 = lds_v2f64(a_row);
        const double jk = fabs(cv.y);
        const double x0 = __dmul_rn(aw0.x, jk), x1 = __dmul_rn(aw1.x, jk);      // a * j_k as the reference rounds it
        const double u0 = __dsub_rn(x0, cv.x), u1 = __dsub_rn(x1, cv.x);        // exact
        double g0 = fma(c67.y, u0, c67.x), g1 = fma(c67.y, u1, c67.x);
        g0 = fma(g0, u0, c45.y); g1 = fma(g1, u1, c45.y);
        g0 = fma(g0, u0, c45.x); g1 = fma(g1, u1, c45.x);
        g0 = fma(g0, u0, c23.y); g1 = fma(g1, u1, c23.y);
        g0 = fma(g0, u0, c23.x); g1 = fma(g1, u1, c23.x);
        g0 = fma(g0, u0, c01.y); g1 = fma(g1, u1, c01.y);
        g0 = fma(g0, u0, c01.x); g1 = fma(g1, u1, c01.x);
        // the hot loop carries no branch: columns whose tile range needs more than one row are redone after the loop
        const bool plain = __double2hiint(cv.y) >= 0;
        any_gather |= !plain && cv.y != 0.0;
        if (DEBRIS) {
            const double h2 = lds_f64(sbase + SMB_H2 + sc * 8);
            g0 *= exp_neg(k0 * h2);
            g1 *= exp_neg(k1 * h2);
        }
        if (plain) {                                                 // the data column and the padding are written elsewhere
            sts_f64(a_g, g0 * aw0.y);
            sts_f64(a_g + 32 * 8, g1 * aw1.y);
        }
        a_row += NW * 16; a_cen += NW * 16; a_g += NW * GLD * 8;
    }
    if (any_gather) {                                                // warp-uniform, rare: per-visibility rows for the columns that need them
        for (int sc = warp; sc < ncol; sc += NW) {
            const double2 cv = lds_v2f64(sbase + SMB_CEN + (buf * GCOLS + sc) * 16);
            if (__double2hiint(cv.y) >= 0 || cv.y == 0.0) continue;
            const double jk = fabs(cv.y);
            double g0 = j0_tab(__dmul_rn(aw0.x, jk), p.tab, last_row), g1 = j0_tab(__dmul_rn(aw1.x, jk), p.tab, last_row);
            if (DEBRIS) {
                const double h2 = lds_f64(sbase + SMB_H2 + sc * 8);
                g0 *= exp_neg(k0 * h2);
                g1 *= exp_neg(k1 * h2);
            }
            const uint32_t ag = sbase + SMB_G + (sc * GLD + lane) * 8;
            sts_f64(ag, g0 * aw0.y);
            sts_f64(ag + 32 * 8, g1 * aw1.y);
        }
    }
}

// Variant of the J0 phase with one column per HALF-warp and four visibilities per lane (lane16, lane16 + 16, + 32, + 48):
// the five 16-byte loads of a column's row (centre / j_k and four coefficient pairs) are issued once for TWO columns --
// the two half-warps of a load instruction read two different rows -- so a column costs half the shared-memory
// wavefronts of its coefficient fetch (the sweep of j0_columns is bound by them: 15 wavefronts against 10 clocks of
// FP64-pipe time per column), at the price of two more Horner chains' worth of registers.
template <bool DEBRIS>
__device__ __forceinline__ void j0_columns4(const GramArgs &p, const uint32_t sbase, const int buf, const int ncol,
                                            const int warp, const int lane, const int last_row)
{
    const int l16 = lane & 15, half = lane >> 4;
    const uint32_t s_vis = sbase + SMB_VIS + buf * (GS * 32);
    const uint32_t s_sw = sbase + SMB_SW + buf * (GS * 8);
    double av[4], kk[4];             // a of this lane's four visibilities (sqrt w is re-read at the store: registers)
#pragma unroll
    for (int m = 0; m < 4; m++) {
        av[m] = lds_f64(s_vis + (l16 + 16 * m) * 32);
        kk[m] = 0.0;
        if (DEBRIS) { const double k = lds_f64(s_vis + (l16 + 16 * m) * 32 + 16); kk[m] = -k * k; }
    }
    const int c0 = 2 * warp + half;
    uint32_t a_row = sbase + SMB_ROW + buf * (GCOLS * ROWB) + c0 * 16;
    uint32_t a_cen = sbase + SMB_CEN + (buf * GCOLS + c0) * 16;
    uint32_t a_g = sbase + SMB_G + (c0 * GLD + l16) * 8;
#pragma unroll 1
    for (int sc = c0; sc < ncol; sc += 2 * NW) {
        const double2 cv = lds_v2f64(a_cen);                             // (centre, +j_k | -j_k: gather | -0: no store)
        const double2 c67 = lds_v2f64(a_row + 3 * GCOLS * 16), c45 = lds_v2f64(a_row + 2 * GCOLS * 16),
                      c23 = lds_v2f64(a_row + GCOLS * 16), c01 = lds_v2f64(a_row);
        const double jk = fabs(cv.y);
        const bool plain = __double2hiint(cv.y) >= 0;                    // one staged row serves the tile
        const bool gather = !plain && cv.y != 0.0;                       // half-warp-uniform, rare
        double h2 = 0.0;
        if (DEBRIS) h2 = lds_f64(sbase + SMB_H2 + sc * 8);
        // two passes of two chains: the row stays in registers, only two evaluations are live at a time (four live chains
        // next to the 60 accumulator registers made ptxas spill inside the DMMA loops)
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const double x0 = __dmul_rn(av[2 * h], jk), x1 = __dmul_rn(av[2 * h + 1], jk);       // a * j_k as the reference rounds it
            const double u0 = __dsub_rn(x0, cv.x), u1 = __dsub_rn(x1, cv.x);                     // exact
            double g0 = fma(c67.y, u0, c67.x), g1 = fma(c67.y, u1, c67.x);
            g0 = fma(g0, u0, c45.y); g1 = fma(g1, u1, c45.y);
            g0 = fma(g0, u0, c45.x); g1 = fma(g1, u1, c45.x);
            g0 = fma(g0, u0, c23.y); g1 = fma(g1, u1, c23.y);
            g0 = fma(g0, u0, c23.x); g1 = fma(g1, u1, c23.x);
            g0 = fma(g0, u0, c01.y); g1 = fma(g1, u1, c01.y);
            g0 = fma(g0, u0, c01.x); g1 = fma(g1, u1, c01.x);
            if (gather) {
                g0 = j0_tab(x0, p.tab, last_row);
                g1 = j0_tab(x1, p.tab, last_row);
            }
            if (DEBRIS) {
                g0 *= exp_neg(kk[2 * h] * h2);
                g1 *= exp_neg(kk[2 * h + 1] * h2);
            }
            if (plain || gather) {                                       // the data column and the padding are written elsewhere
                sts_f64(a_g + 32 * h * 8, g0 * lds_f64(s_sw + (l16 + 32 * h) * 8));
                sts_f64(a_g + (32 * h + 16) * 8, g1 * lds_f64(s_sw + (l16 + 32 * h + 16) * 8));
            }
        }
        a_row += 2 * NW * 16; a_cen += 2 * NW * 16; a_g += 2 * NW * GLD * 8;
    }
}

// ---------------------------------------------------------------------------------------------
// One work item for one warp: per stage, the DMMAs of stage s, then the warp's J0 segments of stage s + 1
// ---------------------------------------------------------------------------------------------
struct ItemCtx {
    uint32_t sbase;
    const double *G;       // generic pointer to the G stages
    int r0, c0, ncolA, nmod, ncol;
    long long q0;          // first tile of the item; its tiles are q0, q0 + qs, q0 + 2 qs, ...
    int qs;
    int nst;
    int ld;
    double *out;
    double my_jk;          // column code (j_k | -1 data column | -2 padding) of column tid - 256: staging duty of this thread
    int dcol;              // local index of the data column in this block, -1 if absent
};

// DMMAs of one tile for one warp.  G: generic pointer to the tile; fr[p]: this lane's fragment offset inside a
// tile row of parity p (the swizzle depends on the parity of the local tile index); ta / tb: first local tile index of
// the A rows / B columns of this warp's register block.
template <int KIND, int NR, int NC>
__device__ __forceinline__ void stage_dmma(double (&acc)[NR * NC > 0 ? NR * NC : 1][2], const double *__restrict__ G,
                                           const int fr0, const int fr1, const int ta, const int tb, const int cbase,
                                           const int nmod)
{
    if (NR * NC == 0) return;
    const double *A0 = G + ta * 8 * GLD + ((ta & 1) ? fr1 : fr0);          // tile rows ta, ta + 2, ...
    const double *A1 = G + ta * 8 * GLD + ((ta & 1) ? fr0 : fr1);          // tile rows ta + 1, ta + 3, ...
    const double *B0 = G + tb * 8 * GLD + ((tb & 1) ? fr1 : fr0);
    const double *B1 = G + tb * 8 * GLD + ((tb & 1) ? fr0 : fr1);
#pragma unroll 2
    for (int ks = 0; ks < GKS; ks++) {
        double af[NR > 0 ? NR : 1];
#pragma unroll
        for (int r = 0; r < NR; r++) af[r] = ((r & 1) ? A1 : A0)[r * 8 * GLD + ks * 4];
        if (KIND == FB_KIND_OFF) {
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const double bf = ((c & 1) ? B1 : B0)[c * 8 * GLD + ks * 4];
#pragma unroll
                for (int r = 0; r < NR; r++) dmma(acc[r * NC + c], af[r], bf);
            }
        } else {
#pragma unroll
            for (int s = 0; s < NR + NC - 1; s++) {
                int ct = cbase + s;
                ct = ct >= nmod ? ct - nmod : ct;
                const double bf = G[ct * 8 * GLD + ((ct & 1) ? fr1 : fr0) + ks * 4];
#pragma unroll
                for (int r = 0; r < NR; r++) {
                    const int dd = s - r;
                    if (dd >= 0 && dd < NC) dmma(acc[r * NC + dd], af[r], bf);
                }
            }
        }
    }
}

template <bool DEBRIS, int KIND, int NR, int NC>
__device__ __forceinline__ void run_item(const GramArgs &p, const ItemCtx &it, const int tid, const int lane)
{
    double acc[NR * NC > 0 ? NR * NC : 1][2];
#pragma unroll
    for (int i = 0; i < (NR * NC > 0 ? NR * NC : 1); i++) acc[i][0] = acc[i][1] = 0.0;
    const int last_row = p.tab_rows - 1;
    const int fr0 = (lane >> 2) * GLD + (lane & 3), fr1 = fr0;     // fragment offset of this lane inside a tile row
    const int ta = it.r0, tb = KIND == FB_KIND_OFF ? it.ncolA / 8 + it.c0 : 0;
    const int cbase = KIND == FB_KIND_OFF ? 0 : (it.r0 + it.c0) % it.nmod;
    const uint32_t sbase = it.sbase;

    // Staging of everything the J0 phase of a tile needs, in two parts.
    // stage_math(ar, buf) -- threads 256.. (the warps with one column less in the J0 phase), one column each: the table
    //   row that serves arguments [ar.x j_k, ar.y j_k], its centre, and whether that one row covers the whole tile.
    // stage_copy(buf, q, q2) -- all threads, cp.async only (no registers held, nothing for the FP64 pipe, so it can fly
    //   during the DMMAs): four lanes copy the four 16-byte pieces of a row, so that a warp-wide copy touches 8 rows,
    //   not 32; the per-visibility scalars (a, sqrt w, kz, sqrt w Re V) of tile q; the range (min a, max a) of tile q2.
    auto stage_math = [&](const double2 ar, const int buf) {
        const int col = tid - 256;
        if (col >= 0 && col < it.ncol) {
            int m = 0;
            double cen = 0.0, jks = -0.0;
            if (it.my_jk >= 0.0) {
                const double xlo = __dmul_rn(ar.x, it.my_jk), xhi = __dmul_rn(ar.y, it.my_jk);
                m = min(__double2loint(fma(xlo + xhi, 0.5 * FB_J0_INVH, J0_MAGIC)), last_row);
                cen = (double)m * FB_J0_H;
                jks = (fabs(xlo - cen) < FB_J0_ACCEPT && fabs(xhi - cen) < FB_J0_ACCEPT) ? it.my_jk : -it.my_jk;
            }
            sts_v2f64(sbase + SMB_CEN + (buf * GCOLS + col) * 16, cen, jks);
            asm volatile("st.shared.s32 [%0], %1;" ::"r"(sbase + SMB_JK + col * 4), "r"(m) : "memory");
        }
    };
    auto stage_copy = [&](const int buf, const long long q, const long long q2) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int col = (tid >> 2) + 128 * h, piece = tid & 3;
            if (col < it.ncol) {
                int m;
                asm volatile("ld.shared.s32 %0, [%1];" : "=r"(m) : "r"(sbase + SMB_JK + col * 4));
                const double2 *src = p.tab + (size_t)m * (FB_J0_ROWLEN / 2) + piece;
                const uint32_t dst = sbase + SMB_ROW + buf * (GCOLS * ROWB) + piece * (GCOLS * 16) + col * 16;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            }
        }
        if (tid < GS) {
            const size_t v = (size_t)q * GS + tid;
            const uint32_t d = sbase + SMB_VIS + buf * (GS * 32) + tid * 32;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(p.a + v) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d + 8), "l"(p.sw + v) : "memory");
#ifdef FB_J0_HALFWARP
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sbase + SMB_SW + (buf * GS + tid) * 8), "l"(p.sw + v) : "memory");
#endif
            if (DEBRIS) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d + 16), "l"(p.kz + v) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d + 24), "l"(p.swV + v) : "memory");
        }
        if (tid == GS && q2 >= 0)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sbase + SMB_AR + (buf ^ 1) * 16), "l"(p.arange + q2) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // J0 phase of this warp
    auto produce = [&](const int buf) {
        if (it.dcol >= 0 && tid >= FB_GRAM_THREADS - GS) {       // data column: G[N][v] = sqrt(w_v) Re V_v
            const int vv = tid - (FB_GRAM_THREADS - GS);
            sts_f64(sbase + SMB_G + (it.dcol * GLD + vv) * 8, lds_f64(sbase + SMB_VIS + buf * (GS * 32) + vv * 32 + 24));
        }
#ifdef FB_J0_HALFWARP
        j0_columns4<DEBRIS>(p, sbase, buf, it.ncol, tid >> 5, lane, last_row);
#else
        j0_columns<DEBRIS>(p, sbase, buf, it.ncol, tid >> 5, lane, last_row);
#endif
    };

    const long long q0 = it.q0, qs = it.qs;
    const int nst = it.nst;
    // ---- prologue: rows and scalars of the first tile, range of the second --------------------------------------
    stage_math(p.arange[q0], 0);
    __syncthreads();
    stage_copy(0, q0, nst > 1 ? q0 + qs : -1);
    // ---- main loop: J0 phase, DMMA phase ------------------------------------------------------------------------
    long long t_j0 = 0, t_mma = 0, t_last = clock64();
    for (int s = 0; s < nst; s++) {
        const int b = s & 1;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                     // tile s staged; DMMAs of tile s - 1 done
        if (p.prof && tid == 0) { const long long t = clock64(); t_mma += t - t_last; t_last = t; }
        produce(b);
        if (s + 1 < nst) stage_math(lds_v2f64(sbase + SMB_AR + (b ^ 1) * 16), b ^ 1);
        __syncthreads();
        if (p.prof && tid == 0) { const long long t = clock64(); t_j0 += t - t_last; t_last = t; }
        if (s + 1 < nst) stage_copy(b ^ 1, q0 + (s + 1) * qs, s + 2 < nst ? q0 + (s + 2) * qs : -1);     // flies during the DMMAs
        stage_dmma<KIND, NR, NC>(acc, it.G, fr0, fr1, ta, tb, cbase, it.nmod);
    }
    if (p.prof && tid == 0) {
        t_mma += clock64() - t_last;
        atomicAdd((unsigned long long *)&p.prof[2 * blockIdx.x], (unsigned long long)t_j0);
        atomicAdd((unsigned long long *)&p.prof[2 * blockIdx.x + 1], (unsigned long long)t_mma);

    }
    // ---- write the partial block ------------------------------------------------------------------------------------
#pragma unroll
    for (int r = 0; r < NR; r++)
#pragma unroll
        for (int c = 0; c < NC; c++)
            *reinterpret_cast<double2 *>(it.out + ((size_t)((it.r0 + r) * it.ld + (it.c0 + c)) * 64 + lane * 2)) =
                make_double2(acc[r * NC + c][0], acc[r * NC + c][1]);
}

// compile-time dispatch over the register-block shape (NR * NC <= FB_ACC, NR, NC <= 5)
#define FB_SHAPE_CASE(K, NR, NC) \
    case NR * 8 + NC: run_item<DEBRIS, K, NR, NC>(p, it, tid, lane); break;
#define FB_DISPATCH_SHAPE(K)                                                                                   \
    switch (nr * 8 + nc) {                                                                                     \
        FB_SHAPE_CASE(K, 1, 1) FB_SHAPE_CASE(K, 1, 2) FB_SHAPE_CASE(K, 1, 3) FB_SHAPE_CASE(K, 1, 4) FB_SHAPE_CASE(K, 1, 5) \
        FB_SHAPE_CASE(K, 2, 1) FB_SHAPE_CASE(K, 2, 2) FB_SHAPE_CASE(K, 2, 3) FB_SHAPE_CASE(K, 2, 4) FB_SHAPE_CASE(K, 2, 5) \
        FB_SHAPE_CASE(K, 3, 1) FB_SHAPE_CASE(K, 3, 2) FB_SHAPE_CASE(K, 3, 3) FB_SHAPE_CASE(K, 3, 4) FB_SHAPE_CASE(K, 3, 5) \
        FB_SHAPE_CASE(K, 4, 1) FB_SHAPE_CASE(K, 4, 2) FB_SHAPE_CASE(K, 4, 3) FB_SHAPE_CASE(K, 4, 4)                      \
        FB_SHAPE_CASE(K, 5, 1) FB_SHAPE_CASE(K, 5, 2) FB_SHAPE_CASE(K, 5, 3)                                            \
        FB_SHAPE_CASE(K, 0, 0)                                                                                 \
    }

template <bool DEBRIS>
__global__ void __launch_bounds__(FB_GRAM_THREADS, 1) k_gram(const GramArgs p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (*p.status) return;
    // this channel's run of tiles in the lane's sorted arrays (device-resident: no host planning per call)
    const int *pad = p.seg + FB_MAX_CHAN + 1;
    const long long tile0 = pad[p.chan] / FB_TV, n_tiles = (pad[p.chan + 1] - pad[p.chan]) / FB_TV;
    const int k_end = p.cta_off[blockIdx.x + 1];
    for (int k = p.cta_off[blockIdx.x]; k < k_end; ++k) {
        const int type = p.items[3 * k], chunk = p.items[3 * k + 1], slot_out = p.items[3 * k + 2];
        const int Ct = p.type_tab[2 * type];
        const FbGramType ty = p.types[type];
        // Chunk c of a type takes the tiles c, c + Ct, c + 2 Ct, ... of the channel's run: every item sees the same mix of
        // dense (long-baseline) and sparse (short-baseline: per-visibility gather path) tiles, so the items of a launch cost
        // the same whatever the baseline distribution -- contiguous ranges left the items holding a type's first tiles up to
        // twice as slow as the rest, a tail the host's cost model cannot see.
        if (chunk >= n_tiles) continue;
        ItemCtx it;
        it.sbase = sbase;
        it.G = reinterpret_cast<const double *>(smem_raw + SMB_G);
        it.q0 = tile0 + chunk;
        it.qs = Ct;
        it.nst = (int)((n_tiles - chunk + Ct - 1) / Ct);
        it.ncolA = ty.a_nt * 8;
        it.ncol = (ty.a_nt + ty.b_nt) * 8;
        it.ld = ty.ld;
        it.out = p.partial + (size_t)slot_out * FB_PSZ;

        // column tables of this block
        __syncthreads();                      // every warp is done with the previous item
        it.my_jk = -2.0;
        {
            const int col = tid < GCOLS ? tid : tid - GCOLS;          // both halves of the CTA look at column code `col`
            double v = -2.0, h = 0.0;
            if (col < it.ncol) {
                const int g = col < it.ncolA ? ty.a_t0 * 8 + col : ty.b_t0 * 8 + (col - it.ncolA);
                if (g < p.N) { v = p.jk[g]; if (DEBRIS) h = p.H2[g]; }
                else if (g == p.N) v = -1.0;
            }
            if (tid >= GCOLS) {
                it.my_jk = v;                                          // staging duty (stage_math)
            } else {
                if (DEBRIS) sts_f64(sbase + SMB_H2 + col * 8, h);
                if (v == -2.0 && col < it.ncol) {                      // zero padding columns stay zero for the whole item
                    for (int x = 0; x < GS; x++) sts_f64(sbase + SMB_G + (col * GLD + x) * 8, 0.0);
                }
            }
        }
        {
            const int gA = p.N - ty.a_t0 * 8, gB = p.N - ty.b_t0 * 8;
            it.dcol = (gA >= 0 && gA < it.ncolA) ? gA : ((gB >= 0 && gB < it.ncol - it.ncolA) ? it.ncolA + gB : -1);
        }

        // this warp's register block (warp-uniform).  Latin-square assignment of (row group, column group) to
        // (slot, sub-partition) balances the DMMA count of the four SM sub-partitions.
        const int smsp = warp & 3, slot = warp >> 2;
        int nr, nc;
        if (ty.kind == FB_KIND_OFF) {
            it.nmod = 1;
            it.r0 = split4_start(ty.a_nt, slot);
            nr = split4_size(ty.a_nt, slot);
            it.c0 = split4_start(ty.b_nt, (slot + smsp) & 3);
            nc = split4_size(ty.b_nt, (slot + smsp) & 3);
        } else {
            it.nmod = ty.a_nt;
            const int D = it.nmod / 2 + 1;
            it.r0 = split4_start(it.nmod, slot);
            nr = split4_size(it.nmod, slot);
            it.c0 = split4_start(D, (slot + smsp) & 3);       // first skew offset d
            nc = split4_size(D, (slot + smsp) & 3);
        }
        if (nr == 0 || nc == 0) { nr = 0; nc = 0; }
        if (ty.kind == FB_KIND_OFF) {
            FB_DISPATCH_SHAPE(FB_KIND_OFF)
        } else {
            FB_DISPATCH_SHAPE(FB_KIND_DIAG)
        }
    }
}

// Sum the partial blocks of one Gram launch in a fixed order into the unscaled Gram S of the call.  One 64-thread block
// per upper tile pair (blockIdx.x) and channel (blockIdx.y); S[chan][pair][i * 8 + jx].  first != 0: S = sum, else S += sum
// (the chunks of the host entry point's pipeline arrive in stream order, so the result is deterministic).
__global__ void __launch_bounds__(64)
k_gram_accumulate(int NT, int P, int npairs, int n_items, const int *__restrict__ seg, const int *__restrict__ tile_panel,
                  const int *__restrict__ panel_t0, const int *__restrict__ panel_nt, const int *__restrict__ pair_code,
                  const FbGramType *__restrict__ types, const int *__restrict__ type_tab, const double *__restrict__ partial,
                  double *__restrict__ S, int first, const int *__restrict__ status)
{
    if (*status) return;
    // decode the upper-triangular tile pair (tr <= tc) from blockIdx.x
    int rem = blockIdx.x, tr = 0;
    while (rem >= NT - tr) { rem -= NT - tr; tr++; }
    const int tc = tr + rem;
    const int chan = blockIdx.y;
    const int i = threadIdx.x >> 3, jx = threadIdx.x & 7;
    if (tr == tc && i > jx) return;
    const int pa = tile_panel[tr], pb = tile_panel[tc];
    const int e_direct = (i * 4 + (jx >> 1)) * 2 + (jx & 1);        // C-fragment slot of element (i, jx)
    const int e_transp = (jx * 4 + (i >> 1)) * 2 + (i & 1);         // ... of element (jx, i)
    int type, idx;
    const int *pc = pair_code + (pa * P + pb) * 3;
    if (pa < pb) {
        const int lrp = tr - panel_t0[pa], h0 = pc[2];
        const int half = lrp >= h0 ? 1 : 0;
        type = pc[half];
        idx = ((lrp - half * h0) * types[type].ld + (tc - panel_t0[pb])) * 64 + e_direct;
    } else {
        type = pc[0];
        const int n = panel_nt[pa], D = n / 2 + 1;
        const int lr = tr - panel_t0[pa], lc = tc - panel_t0[pa], d = lc - lr;
        if (d < D)
            idx = (lr * D + d) * 64 + e_direct;
        else   // stored as the transposed tile (row tile lc, offset n - d)
            idx = (lc * D + (n - d)) * 64 + e_transp;
    }
    const int *pad = seg + FB_MAX_CHAN + 1;
    const long long n_tiles = (pad[chan + 1] - pad[chan]) / FB_TV;
    const double *part = partial + (size_t)chan * n_items * FB_PSZ;
    double s = 0.0;
    const int Ct = type_tab[2 * type], first_slot = type_tab[2 * type + 1];
    for (int c = 0; c < Ct && c < n_tiles; c++) s += part[(size_t)(first_slot + c) * FB_PSZ + idx];      // chunk c is empty when c >= n_tiles
    double *out = S + ((size_t)chan * npairs + blockIdx.x) * 64 + threadIdx.x;
    *out = first ? s : *out + s;
}

// M = diag(c) S[:N, :N] diag(c) (mirrored), j = diag(c) S[:N, N].  Grid as k_gram_accumulate.
__global__ void __launch_bounds__(64)
k_gram_scale(int N, int NT, int npairs, const double *__restrict__ S, const double *__restrict__ ck, double scale,
             double *__restrict__ M, double *__restrict__ jvec, const int *__restrict__ status)
{
    if (*status) return;
    int rem = blockIdx.x, tr = 0;
    while (rem >= NT - tr) { rem -= NT - tr; tr++; }
    const int tc = tr + rem;
    const int chan = blockIdx.y;
    const int i = threadIdx.x >> 3, jx = threadIdx.x & 7;
    const int row = tr * 8 + i, col = tc * 8 + jx;
    if (tr == tc && i > jx) return;
    const double s = S[((size_t)chan * npairs + blockIdx.x) * 64 + threadIdx.x];
    double *Mc = M + (size_t)chan * N * N;
    if (col < N) {           // row <= col < N
        const double val = ((ck[row] * scale) * (ck[col] * scale)) * s;
        Mc[(size_t)row * N + col] = val;
        Mc[(size_t)col * N + row] = val;
    } else if (col == N && row < N) {
        jvec[(size_t)chan * N + row] = (ck[row] * scale) * s;
    }
}

// K8: predicted visibilities V_i = sum_k H_ik I_k, H_ik = c_k J0(a_i j_k) scale_ik   (statistical_models.py:279-329).
// One warp per visibility, lanes stride over the modes; fixed shuffle tree.  Arguments beyond the J0 table raise `flag`
// (flag[0] = 1, flag[1] = bits of the largest argument met): the host grows the table and repeats the call.
__device__ __forceinline__ double predict_row(double a, double kk, int N, const double *__restrict__ jk, const double *__restrict__ ck,
                                              const double *__restrict__ Ik, const double *__restrict__ H2, double scale,
                                              const double2 *__restrict__ tab, int rows, int lane, unsigned long long *flag)
{
    const double x_table = (double)(rows - 3) * FB_J0_H;
    double acc = 0.0;
    for (int k = lane; k < N; k += 32) {
        const double x = __dmul_rn(a, jk[k]);
        if (x > x_table) { flag[0] = 1ull; atomicMax(flag + 1, (unsigned long long)__double_as_longlong(x)); }
        double h = ck[k] * j0_tab(x, tab, rows - 1);
        h *= H2 ? exp(kk * H2[k]) : scale;
        acc = fma(h, Ik[k], acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    return acc;                      // valid in lane 0
}

__global__ void __launch_bounds__(256)
k_predict(int64_t n, int N, const double *__restrict__ q, const double *__restrict__ kz, double invQmax,
          const double *__restrict__ jk, const double *__restrict__ ck, const double *__restrict__ Ik,
          const double *__restrict__ H2, double scale, const double2 *__restrict__ tab, int rows, double *__restrict__ V,
          unsigned long long *__restrict__ flag)
{
    const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const double a = __dmul_rn(q[i], invQmax);
    double kk = 0.0;
    if (H2) { kk = kz[i]; kk = -kk * kk; }
    const double acc = predict_row(a, kk, N, jk, ck, Ik, H2, scale, tab, rows, lane, flag);
    if (lane == 0) V[i] = acc;
}

// FrankRadialFit.predict (radial_fitters.py:56-98) in one pass over sky-plane baselines: deproject (geometry.py:111-131),
// q = hypot, V = H(q) I, then undo_correction (geometry.py:239-265): re-project the deprojected baseline and rotate the
// phase, V_sky = V (cos phi + i sin phi) -- every step a single correctly rounded operation in the reference's order.
__global__ void __launch_bounds__(256)
k_predict_sky(int64_t n, int N, const double *__restrict__ u, const double *__restrict__ v, fb_geometry g, double invQmax,
              const double *__restrict__ jk, const double *__restrict__ ck, const double *__restrict__ Ik,
              const double *__restrict__ H2, double scale, const double2 *__restrict__ tab, int rows, double2 *__restrict__ Vsky,
              unsigned long long *__restrict__ flag)
{
    const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const double ui = u[i], vi = v[i];
    double up = __dsub_rn(__dmul_rn(ui, g.cos_pa), __dmul_rn(vi, g.sin_pa));
    const double vp = __dadd_rn(__dmul_rn(ui, g.sin_pa), __dmul_rn(vi, g.cos_pa));
    const double kz = __dmul_rn(up, g.sin_inc);
    up = __dmul_rn(up, g.cos_inc);
    const double a = __dmul_rn(hypot_glibc(up, vp), invQmax);
    const double acc = predict_row(a, -kz * kz, N, jk, ck, Ik, H2, scale, tab, rows, lane, flag);
    if (lane == 0) {
        // reproject: u'' = u' / cos(inc); rotate by -PA (sin(PA) * -1), geometry.py:115-127
        const double ud = __ddiv_rn(up, g.cos_inc), nst = __dmul_rn(g.sin_pa, -1.0);
        const double ur = __dsub_rn(__dmul_rn(ud, g.cos_pa), __dmul_rn(vp, nst));
        const double vr = __dadd_rn(__dmul_rn(ud, nst), __dmul_rn(vp, g.cos_pa));
        const double phi = __dadd_rn(__dmul_rn(ur, g.a_ra), __dmul_rn(vr, g.a_dec));
        double s, c;
        sincos(phi, &s, &c);
        Vsky[i] = make_double2(acc * c, acc * s);
    }
}

// far = 0: the row nearest to x (|t| <= 1/32, the gather path); far = 1: the neighbouring row on the other side
// (1/32 <= |t| <= 1/16), i.e. the worst case the producers accept
__global__ void k_j0_debug(int64_t n, const double *__restrict__ x, double *__restrict__ out, const double2 *tab, int rows, int far)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!far) { out[i] = j0_tab(x[i], tab, rows - 1); return; }
    const double xv = x[i];
    int m = min(__double2loint(fma(xv, FB_J0_INVH, J0_MAGIC)), rows - 1);
    double t = fma((double)m, -FB_J0_H, xv);
    int m2 = t >= 0.0 ? m + 1 : m - 1;
    if (m2 < 0 || m2 > rows - 1) m2 = m;
    t = fma((double)m2, -FB_J0_H, xv);
    const double2 *row = tab + (size_t)m2 * (FB_J0_ROWLEN / 2);
    const double2 c01 = row[0], c23 = row[1], c45 = row[2], c67 = row[3];
    double y = fma(c67.y, t, c67.x);
    y = fma(y, t, c45.y);
    y = fma(y, t, c45.x);
    y = fma(y, t, c23.y);
    y = fma(y, t, c23.x);
    y = fma(y, t, c01.y);
    out[i] = fma(y, t, c01.x);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
int fb_build_j0_table(fb_ctx *ctx, double x_max)
{
    std::vector<double> tab;
    fb_j0_build(x_max, tab);
    FB_CUDA(cudaDeviceSynchronize());                     // a kernel of an earlier call may still read the old table
    if (ctx->d_tab) FB_CUDA(cudaFree(ctx->d_tab));
    ctx->d_tab = nullptr;
    FB_CUDA(cudaMalloc(&ctx->d_tab, tab.size() * sizeof(double)));
    FB_CUDA(cudaMemcpy(ctx->d_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
    ctx->tab_rows = fb_j0_rows_for(x_max);
    return 0;
}

namespace {

int cdiv4(int n) { return (n + 3) / 4; }
bool off_fits(int a, int b) { return cdiv4(a) * cdiv4(b) <= FB_ACC && cdiv4(a) <= 5 && cdiv4(b) <= 5 && (a + b) * 8 <= FB_GCOLS; }
bool diag_fits(int n) { return cdiv4(n) * cdiv4(n / 2 + 1) <= FB_ACC && cdiv4(n) <= 5 && n * 8 <= FB_GCOLS; }

// Cost of a block per tile of FB_TV visibilities, in FP64-pipe clocks of the busiest SM sub-partition:
// a DMMA holds the pipe for 16 clocks, a J0 evaluation is ~10 FP64 instructions of 2 clocks for 32 lanes.
double block_cost(const FbGramType &ty)
{
    const int nrows = ty.a_nt, ncols = ty.kind == FB_KIND_OFF ? ty.b_nt : ty.a_nt / 2 + 1;
    int worst = 0;
    for (int s = 0; s < 4; s++) {
        int t = 0;
        for (int slot = 0; slot < 4; slot++) t += split4_size(nrows, slot) * split4_size(ncols, (slot + s) & 3);
        worst = std::max(worst, t);
    }
    const int cols = 8 * (ty.kind == FB_KIND_OFF ? ty.a_nt + ty.b_nt : ty.a_nt);
    static const double j0w = getenv("FB_J0_COST") ? atof(getenv("FB_J0_COST")) : 12.0;
    return 16.0 * (FB_TV / 4) * worst + j0w * cols;
}

struct GramPlan {
    int P = 0;
    std::vector<int> panel_t0, panel_nt;
    std::vector<FbGramType> types;
    std::vector<int> pair_code;
    double cost = 0.0;
    bool ok = false;
};

GramPlan make_plan(int NT, int P)
{
    GramPlan pl;
    pl.P = P;
    pl.panel_t0.resize(P);
    pl.panel_nt.resize(P);
    const int base = NT / P, rem = NT % P;
    int t = 0;
    for (int p = 0; p < P; p++) {
        pl.panel_t0[p] = t;
        pl.panel_nt[p] = base + (p < rem ? 1 : 0);
        t += pl.panel_nt[p];
    }
    pl.pair_code.assign((size_t)P * P * 3, -1);
    for (int pa = 0; pa < P; pa++)
        for (int pb = pa + 1; pb < P; pb++) {
            int *pc = &pl.pair_code[((size_t)pa * P + pb) * 3];
            const int na = pl.panel_nt[pa], nb = pl.panel_nt[pb];
            if (off_fits(na, nb)) {
                pc[0] = (int)pl.types.size(); pc[1] = -1; pc[2] = na;
                pl.types.push_back({FB_KIND_OFF, pl.panel_t0[pa], na, pl.panel_t0[pb], nb, nb});
            } else {
                const int h0 = (na + 1) / 2;
                if (!off_fits(h0, nb) || na - h0 < 1) return pl;
                pc[0] = (int)pl.types.size(); pc[2] = h0;
                pl.types.push_back({FB_KIND_OFF, pl.panel_t0[pa], h0, pl.panel_t0[pb], nb, nb});
                pc[1] = (int)pl.types.size();
                pl.types.push_back({FB_KIND_OFF, pl.panel_t0[pa] + h0, na - h0, pl.panel_t0[pb], nb, nb});
            }
        }
    for (int p = 0; p < P; p++) {
        if (!diag_fits(pl.panel_nt[p])) return pl;
        pl.pair_code[((size_t)p * P + p) * 3] = (int)pl.types.size();
        pl.types.push_back({FB_KIND_DIAG, pl.panel_t0[p], pl.panel_nt[p], 0, 0, pl.panel_nt[p] / 2 + 1});
    }
    for (const FbGramType &ty : pl.types) pl.cost += block_cost(ty);
    pl.ok = true;
    return pl;
}

}  // namespace

static int build_work_table(fb_ctx *ctx);

// Choose the panel decomposition of the NT x NT tile grid (cheapest of a few panel counts) and upload it.
int fb_build_gram_plan(fb_ctx *ctx)
{
    const int NT = ctx->NT;
    GramPlan best;
    const int Pmin = (NT + FB_PT - 1) / FB_PT;
    for (int P = Pmin; P <= std::min(NT, Pmin + 3); P++) {
        GramPlan pl = make_plan(NT, P);
        if (pl.ok && (!best.ok || pl.cost < best.cost)) best = pl;
    }
    if (!best.ok) FB_FAIL(-16, "fb_dht_setup: no feasible block decomposition");
    const int P = best.P;
    ctx->P = P;
    if (getenv("FB_GRAM_PROF")) {
        fprintf(stderr, "[fb_gram plan] NT=%d P=%d types=%d cost=%.0f :", NT, P, (int)best.types.size(), best.cost);
        for (const FbGramType &ty : best.types) fprintf(stderr, " %s%dx%d", ty.kind == FB_KIND_OFF ? "O" : "D", ty.a_nt, ty.kind == FB_KIND_OFF ? ty.b_nt : ty.a_nt / 2 + 1);
        fprintf(stderr, "\n");
    }
    ctx->h_types = best.types;
    ctx->ntypes = (int)best.types.size();
    std::vector<int> tile_panel(NT);
    for (int p = 0; p < P; p++)
        for (int i = 0; i < best.panel_nt[p]; i++) tile_panel[best.panel_t0[p] + i] = p;
    for (void **p : {(void **)&ctx->d_types, (void **)&ctx->d_tile_panel, (void **)&ctx->d_panel_t0,
                     (void **)&ctx->d_panel_nt, (void **)&ctx->d_pair_code}) {
        if (*p) FB_CUDA(cudaFree(*p));
        *p = nullptr;
    }
    FB_CUDA(cudaMalloc(&ctx->d_types, sizeof(FbGramType) * ctx->ntypes));
    FB_CUDA(cudaMalloc(&ctx->d_tile_panel, sizeof(int) * NT));
    FB_CUDA(cudaMalloc(&ctx->d_panel_t0, sizeof(int) * P));
    FB_CUDA(cudaMalloc(&ctx->d_panel_nt, sizeof(int) * P));
    FB_CUDA(cudaMalloc(&ctx->d_pair_code, sizeof(int) * P * P * 3));
    FB_CUDA(cudaMemcpy(ctx->d_types, ctx->h_types.data(), sizeof(FbGramType) * ctx->ntypes, cudaMemcpyHostToDevice));
    FB_CUDA(cudaMemcpy(ctx->d_tile_panel, tile_panel.data(), sizeof(int) * NT, cudaMemcpyHostToDevice));
    FB_CUDA(cudaMemcpy(ctx->d_panel_t0, best.panel_t0.data(), sizeof(int) * P, cudaMemcpyHostToDevice));
    FB_CUDA(cudaMemcpy(ctx->d_panel_nt, best.panel_nt.data(), sizeof(int) * P, cudaMemcpyHostToDevice));
    FB_CUDA(cudaMemcpy(ctx->d_pair_code, best.pair_code.data(), sizeof(int) * P * P * 3, cudaMemcpyHostToDevice));
    return build_work_table(ctx);
}

// Work items = (type, chunk of the visibilities).  Chunks per type proportional to the type's cost; the items are dealt
// to the CTAs longest-first (deterministic).  The table depends on the block decomposition and the grid only (the
// kernel derives a chunk's tile range from the channel's tile count, which it reads from device memory), so it is built
// once per fb_dht_setup and stays on the device.
static int build_work_table(fb_ctx *ctx)
{
    const int ntypes = ctx->ntypes;
    const int grid = ctx->num_sms;
    std::vector<double> cost(ntypes);
    double total = 0.0;
    for (int t = 0; t < ntypes; t++) { cost[t] = block_cost(ctx->h_types[t]); total += cost[t]; }
    // number of items: a multiple of the grid (6 per SM, more when there are many types), split over the types by
    // largest remainders so that the total is exact
    long long target = std::max<long long>(6LL * grid, 4LL * ntypes);
    target = (target + grid - 1) / grid * grid;
    std::vector<int> C(ntypes, 1);
    {
        std::vector<std::pair<double, int>> rem(ntypes);
        long long sum = 0;
        for (int t = 0; t < ntypes; t++) {
            const double x = (double)target * cost[t] / total;
            long long c = std::max<long long>(1, (long long)std::floor(x));
            C[t] = (int)c;
            sum += c;
            rem[t] = {x - std::floor(x), t};
        }
        std::sort(rem.begin(), rem.end(), [](const std::pair<double, int> &x, const std::pair<double, int> &y) {
            return x.first != y.first ? x.first > y.first : x.second < y.second;
        });
        for (int i = 0; sum < target && i < ntypes; i++, sum++) C[rem[i].second]++;
    }
    std::vector<int> type_tab(2 * ntypes);
    int n_items = 0;
    for (int t = 0; t < ntypes; t++) { type_tab[2 * t] = C[t]; type_tab[2 * t + 1] = n_items; n_items += C[t]; }
    struct Item { double cost; int type, chunk; };
    std::vector<Item> order;
    order.reserve(n_items);
    for (int t = 0; t < ntypes; t++)
        for (int c = 0; c < C[t]; c++) order.push_back({cost[t] / (double)C[t], t, c});
    std::stable_sort(order.begin(), order.end(), [](const Item &x, const Item &y) { return x.cost > y.cost; });
    std::vector<std::vector<int>> per_cta(grid);
    {
        typedef std::pair<double, int> Load;     // (load, cta), smallest load first, ties by CTA index
        std::priority_queue<Load, std::vector<Load>, std::greater<Load>> heap;
        for (int b = 0; b < grid; b++) heap.push({0.0, b});
        for (int i = 0; i < n_items; i++) {
            Load l = heap.top();
            heap.pop();
            per_cta[l.second].push_back(i);
            heap.push({l.first + order[i].cost, l.second});
        }
    }
    // work buffer: [grid + 1] CTA offsets | [3 * n_items] items | [2 * ntypes] type table
    std::vector<int> work(grid + 1 + 3 * (size_t)n_items + 2 * ntypes);
    {
        int pos = 0;
        for (int b = 0; b < grid; b++) {
            work[b] = pos;
            for (int i : per_cta[b]) {
                const Item &it = order[i];
                int *w = &work[grid + 1 + 3 * (size_t)pos];
                w[0] = it.type; w[1] = it.chunk; w[2] = type_tab[2 * it.type + 1] + it.chunk;
                pos++;
            }
        }
        work[grid] = pos;
        std::copy(type_tab.begin(), type_tab.end(), work.begin() + grid + 1 + 3 * (size_t)n_items);
    }
    if (ctx->d_work) FB_CUDA(cudaFree(ctx->d_work));
    ctx->d_work = nullptr;
    FB_CUDA(cudaMalloc(&ctx->d_work, sizeof(int) * work.size()));
    FB_CUDA(cudaMemcpy(ctx->d_work, work.data(), sizeof(int) * work.size(), cudaMemcpyHostToDevice));
    ctx->n_items = n_items;
    return 0;
}

// Enqueue k_gram for channel `chan` of the lane's sorted visibilities; partial blocks go to the lane's set `chan`.
int fb_enqueue_gram(fb_ctx *ctx, FbLane &ln, int chan, int vis_model)
{
    const int grid = ctx->num_sms;
    const int n_items = ctx->n_items;
    GramArgs args;
    args.a = ln.d_a; args.sw = ln.d_sw; args.swV = ln.d_swV; args.kz = ln.d_kz; args.arange = (const double2 *)ln.d_amid;
    args.seg = ln.d_seg; args.chan = chan; args.status = ctx->d_status;
    args.jk = ctx->d_jk; args.tab = ctx->d_tab; args.tab_rows = ctx->tab_rows;
    args.N = ctx->N;
    args.types = ctx->d_types;
    args.cta_off = ctx->d_work; args.items = ctx->d_work + grid + 1; args.type_tab = ctx->d_work + grid + 1 + 3 * (size_t)n_items;
    args.H2 = ctx->d_H2; args.partial = ln.d_partial + (size_t)chan * n_items * FB_PSZ;
    args.prof = nullptr;
    static const bool prof = getenv("FB_GRAM_PROF") != nullptr;
    if (prof) {
        FB_CUDA(cudaMalloc(&args.prof, sizeof(long long) * 2 * grid));
        FB_CUDA(cudaMemsetAsync(args.prof, 0, sizeof(long long) * 2 * grid, ln.stream));
    }
    if (vis_model == FB_MODEL_DEBRIS) {
        static bool attr = false;
        if (!attr) { FB_CUDA(cudaFuncSetAttribute(k_gram<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GRAM_SMEM_BYTES)); attr = true; }
        k_gram<true><<<grid, FB_GRAM_THREADS, GRAM_SMEM_BYTES, ln.stream>>>(args);
    } else {
        static bool attr = false;
        if (!attr) { FB_CUDA(cudaFuncSetAttribute(k_gram<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GRAM_SMEM_BYTES)); attr = true; }
        k_gram<false><<<grid, FB_GRAM_THREADS, GRAM_SMEM_BYTES, ln.stream>>>(args);
    }
    FB_CUDA(cudaGetLastError());
    if (args.prof) {
        std::vector<long long> h(2 * grid);
        FB_CUDA(cudaStreamSynchronize(ln.stream));
        FB_CUDA(cudaMemcpy(h.data(), args.prof, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost));
        long long j0_sum = 0, mma_sum = 0, j0_max = 0, mma_max = 0, tot_max = 0, tot_min = -1;
        for (int b = 0; b < grid; b++) {
            j0_sum += h[2 * b]; mma_sum += h[2 * b + 1];
            j0_max = std::max(j0_max, h[2 * b]); mma_max = std::max(mma_max, h[2 * b + 1]);
            tot_max = std::max(tot_max, h[2 * b] + h[2 * b + 1]);
            tot_min = tot_min < 0 ? h[2 * b] + h[2 * b + 1] : std::min(tot_min, h[2 * b] + h[2 * b + 1]);
        }
        fprintf(stderr, "[fb_gram prof] chan=%d  per-CTA clocks: J0 avg %.0f max %lld | DMMA avg %.0f max %lld | total min %lld max %lld\n",
                chan, (double)j0_sum / grid, j0_max, (double)mma_sum / grid, mma_max, tot_min, tot_max);
        cudaFree(args.prof);
    }
    return 0;
}

// Fold the lane's partial blocks (all channels) into the call's unscaled Gram S.
int fb_enqueue_accumulate(fb_ctx *ctx, FbLane &ln, int nchan, int first)
{
    const int npairs = ctx->NT * (ctx->NT + 1) / 2;
    const int grid = ctx->num_sms;
    k_gram_accumulate<<<dim3(npairs, nchan), 64, 0, ln.stream>>>(ctx->NT, ctx->P, npairs, ctx->n_items, ln.d_seg, ctx->d_tile_panel,
                                                              ctx->d_panel_t0, ctx->d_panel_nt, ctx->d_pair_code, ctx->d_types,
                                                              ctx->d_work + grid + 1 + 3 * (size_t)ctx->n_items, ln.d_partial,
                                                              ctx->d_S, first, ctx->d_status);
    FB_CUDA(cudaGetLastError());
    return 0;
}

// S -> M, j (scaled, mirrored) for every channel.
int fb_enqueue_scale(fb_ctx *ctx, cudaStream_t st, int nchan, double model_scale, double *dev_M, double *dev_j)
{
    const int npairs = ctx->NT * (ctx->NT + 1) / 2;
    k_gram_scale<<<dim3(npairs, nchan), 64, 0, st>>>(ctx->N, ctx->NT, npairs, ctx->d_S, ctx->d_ck, model_scale, dev_M, dev_j,
                                                   ctx->d_status);
    FB_CUDA(cudaGetLastError());
    return 0;
}

static int debug_j0_impl(fb_ctx *ctx, int64_t n, const double *host_x, double *host_out, int far)
{
    if (!ctx || !ctx->d_tab) return -1;
    double *dx = nullptr, *dout = nullptr;
    FB_CUDA(cudaSetDevice(ctx->device));
    FB_CUDA(cudaMalloc(&dx, sizeof(double) * n));
    FB_CUDA(cudaMalloc(&dout, sizeof(double) * n));
    FB_CUDA(cudaMemcpy(dx, host_x, sizeof(double) * n, cudaMemcpyHostToDevice));
    k_j0_debug<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, dx, dout, ctx->d_tab, ctx->tab_rows, far);
    FB_CUDA(cudaGetLastError());
    FB_CUDA(cudaStreamSynchronize(ctx->stream));
    FB_CUDA(cudaMemcpy(host_out, dout, sizeof(double) * n, cudaMemcpyDeviceToHost));
    cudaFree(dx);
    cudaFree(dout);
    return 0;
}

extern "C" int fb_debug_j0(fb_ctx *ctx, int64_t n, const double *host_x, double *host_out)
{
    return debug_j0_impl(ctx, n, host_x, host_out, 0);
}

extern "C" int fb_debug_j0_far(fb_ctx *ctx, int64_t n, const double *host_x, double *host_out)
{
    return debug_j0_impl(ctx, n, host_x, host_out, 1);
}

// shared driver of the prediction entry points: upload I (and H2), run `launch`, grow the J0 table and repeat if the
// kernel met arguments beyond it
template <typename Launch>
static int predict_driver(fb_ctx *ctx, const double *host_I, int vis_model, const double *host_H2, Launch launch)
{
    const int N = ctx->N;
    const bool debris = vis_model == FB_MODEL_DEBRIS;
    if (ctx->predI_cap < N) {
        FB_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->d_predI) FB_CUDA(cudaFree(ctx->d_predI));
        ctx->d_predI = nullptr;
        FB_CUDA(cudaMalloc(&ctx->d_predI, sizeof(double) * (N + 2)));
        ctx->predI_cap = N;
    }
    unsigned long long *d_flag = (unsigned long long *)(ctx->d_predI + ctx->predI_cap);
    if (debris) {
        std::vector<double> h2(ctx->NC, 0.0);
        for (int k = 0; k < N; k++) h2[k] = host_H2[k];
        if (h2 != ctx->h_H2) {
            FB_CUDA(cudaDeviceSynchronize());
            FB_CUDA(cudaMemcpy(ctx->d_H2, h2.data(), sizeof(double) * ctx->NC, cudaMemcpyHostToDevice));
            ctx->h_H2 = h2;
        }
    }
    FB_CUDA(cudaMemcpyAsync(ctx->d_predI, host_I, sizeof(double) * N, cudaMemcpyHostToDevice, ctx->stream));
    for (int attempt = 0; attempt < 2; attempt++) {
        FB_CUDA(cudaMemsetAsync(d_flag, 0, 16, ctx->stream));
        launch(ctx->d_predI, debris ? ctx->d_H2 : nullptr, d_flag);
        FB_CUDA(cudaGetLastError());
        unsigned long long fl[2] = {0, 0};
        FB_CUDA(cudaMemcpyAsync(fl, d_flag, 16, cudaMemcpyDeviceToHost, ctx->stream));
        FB_CUDA(cudaStreamSynchronize(ctx->stream));
        if (!fl[0]) return 0;
        double xneed;
        memcpy(&xneed, &fl[1], 8);
        int rc = fb_build_j0_table(ctx, xneed * 1.05);
        if (rc) return rc;
    }
    FB_FAIL(-17, "fb_predict_visibilities: J0 table could not be grown to cover the data");
}

extern "C" int fb_predict_visibilities_dev(fb_ctx *ctx, int64_t n, const double *dev_q, const double *dev_kz, const double *host_I,
                                           int vis_model, double model_scale, const double *host_H2, double *dev_V)
{
    if (!ctx) return -1;
    if (ctx->N == 0) FB_FAIL(-11, "fb_predict_visibilities: fb_dht_setup has not been called");
    if (n < 0 || !dev_q || !host_I || !dev_V) FB_FAIL(-12, "fb_predict_visibilities: bad arguments");
    if (vis_model == FB_MODEL_DEBRIS && (!host_H2 || !dev_kz)) FB_FAIL(-15, "fb_predict_visibilities: debris model needs kz and H2");
    if (n == 0) return 0;
    FB_CUDA(cudaSetDevice(ctx->device));
    return predict_driver(ctx, host_I, vis_model, host_H2, [&](const double *d_I, const double *d_H2, unsigned long long *d_flag) {
        k_predict<<<(unsigned)((n + 7) / 8), 256, 0, ctx->stream>>>(n, ctx->N, dev_q, dev_kz, ctx->invQmax, ctx->d_jk, ctx->d_ck, d_I, d_H2,
                                                                  model_scale, ctx->d_tab, ctx->tab_rows, dev_V, d_flag);
    });
}

extern "C" int fb_predict_sky_dev(fb_ctx *ctx, int64_t n, const double *dev_u, const double *dev_v, const fb_geometry *geom,
                                  const double *host_I, int vis_model, double model_scale, const double *host_H2, double *dev_V_reim)
{
    if (!ctx) return -1;
    if (ctx->N == 0) FB_FAIL(-11, "fb_predict_sky: fb_dht_setup has not been called");
    if (n < 0 || !dev_u || !dev_v || !geom || !host_I || !dev_V_reim) FB_FAIL(-12, "fb_predict_sky: bad arguments");
    if (vis_model == FB_MODEL_DEBRIS && !host_H2) FB_FAIL(-15, "fb_predict_sky: debris model needs H2");
    if (n == 0) return 0;
    FB_CUDA(cudaSetDevice(ctx->device));
    const fb_geometry g = *geom;
    return predict_driver(ctx, host_I, vis_model, host_H2, [&](const double *d_I, const double *d_H2, unsigned long long *d_flag) {
        k_predict_sky<<<(unsigned)((n + 7) / 8), 256, 0, ctx->stream>>>(n, ctx->N, dev_u, dev_v, g, ctx->invQmax, ctx->d_jk, ctx->d_ck, d_I, d_H2,
                                                                      model_scale, ctx->d_tab, ctx->tab_rows, (double2 *)dev_V_reim, d_flag);
    });
}

// host arrays: staged through the lane-0 input buffer
extern "C" int fb_predict_visibilities(fb_ctx *ctx, int64_t n, const double *host_q, const double *host_kz, const double *host_I,
                                       int vis_model, double model_scale, const double *host_H2, double *host_V)
{
    if (!ctx) return -1;
    if (ctx->N == 0) FB_FAIL(-11, "fb_predict_visibilities: fb_dht_setup has not been called");
    if (n < 0 || !host_q || !host_I || !host_V) FB_FAIL(-12, "fb_predict_visibilities: bad arguments");
    if (vis_model == FB_MODEL_DEBRIS && (!host_H2 || !host_kz)) FB_FAIL(-15, "fb_predict_visibilities: debris model needs kz and H2");
    if (n == 0) return 0;
    FB_CUDA(cudaSetDevice(ctx->device));
    FbLane &ln = ctx->lane[0];
    const int64_t need = 3 * n + 8;
    if (need > ln.in_cap) {
        FB_CUDA(cudaDeviceSynchronize());
        if (ln.d_in) FB_CUDA(cudaFree(ln.d_in));
        ln.d_in = nullptr;
        FB_CUDA(cudaMalloc(&ln.d_in, sizeof(double) * need));
        ln.in_cap = need;
    }
    double *d_q = ln.d_in, *d_kz = d_q + n, *d_V = d_kz + n;
    const bool debris = vis_model == FB_MODEL_DEBRIS;
    FB_CUDA(cudaMemcpyAsync(d_q, host_q, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    if (debris) FB_CUDA(cudaMemcpyAsync(d_kz, host_kz, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    int rc = fb_predict_visibilities_dev(ctx, n, d_q, debris ? d_kz : nullptr, host_I, vis_model, model_scale, host_H2, d_V);
    if (rc) return rc;
    FB_CUDA(cudaMemcpyAsync(host_V, d_V, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    FB_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
