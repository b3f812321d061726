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
//       (1) the design-matrix tile G[mode][vis] of the block's columns is formed in shared memory -- also on the tensor
//           pipe: the visibilities are sorted by baseline, so ONE row of the J0 table (a degree-7 polynomial; rows
//           overlap, fb_j0_table.h) serves a (column, tile) pair, and re-centred on the tile it makes G a rank-8
//           product of per-visibility powers and per-column coefficients (j0_gemm below).  Rows are staged by cp.async
//           two tiles ahead and re-centred one tile ahead; columns whose tile range needs more than one row are redone
//           per visibility from the table (rare);
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
    long long *prof;      // builds with -DFB_GRAM_CLOCKS, FB_GRAM_PROF=1: [grid][2] clocks thread 0 spent in the J0 / DMMA phases
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
constexpr int SMB_ROW = SMB_G + GCOLS * GLD * 8;           // [4][GCOLS] double2        staged J0 table rows of the NEXT tile (coefficient pair major)
constexpr int SMB_CM = SMB_ROW + GCOLS * ROWB;             // [2][2][GCOLS][4] doubles  per-column polynomials in the tile variable: (buffer, degrees 0-3 | 4-7, column, degree)
constexpr int SMB_P = SMB_CM + 2 * 2 * GCOLS * 32;         // [2][2][GS][4] doubles     sqrt(w) * powers of the tile variable: (buffer, degrees 0-3 | 4-7, visibility, degree)
constexpr int SMB_JK = SMB_P + 2 * 2 * GS * 32;            // [GCOLS] ints (+ pad)      table row chosen for the tile being staged
constexpr int SMB_H2 = SMB_JK + GCOLS * 8;                 // [GCOLS] doubles           debris H2_k
constexpr int SMB_VIS = SMB_H2 + GCOLS * 8;                // [2][GS][4] doubles        (a, sqrt w, kz, sqrt w Re V) per visibility of a tile
constexpr int SMB_AR = SMB_VIS + 2 * GS * 32;              // [4] double2               ring of (min a, max a) of the coming tiles
constexpr int SMB_COL = SMB_AR + 4 * 16;                   // [GCOLS] double2           (column code j_k | -1 data column | -2 padding, j_k / j_ref)
constexpr int SMB_JREF = SMB_COL + GCOLS * 16;             // double (+ pad)            largest j_k of the block: the tile variable is s = (a - a_c) j_ref
constexpr int SMB_LIST = SMB_JREF + 16;                    // [2][GCOLS] shorts         columns the polynomial does not serve (and the data column), per buffer
constexpr int SMB_NLIST = SMB_LIST + 2 * GCOLS * 2;        // [2] ints (+ pad)          their number
constexpr int GRAM_SMEM_BYTES = SMB_NLIST + 16;
static_assert(GRAM_SMEM_BYTES <= 227 * 1024, "shared memory of k_gram");

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

// J0 phase of one warp for one tile, on the tensor pipe.  For the visibilities of a tile (sorted: a narrow range of
// baselines a around the tile centre a_c) and a column k whose table row is valid over the tile,
//     J0(a j_k) = sum_d c_d (a j_k - cen)^d = sum_d b_dk s^d,     s = (a - a_c) j_ref,
// with b_dk the row's polynomial shifted to the tile centre and rescaled (prepare() in run_item, one thread per column), so
// that the design-matrix tile is a rank-8 product
//     G[v][k] = sqrt(w_v) J0(a_v j_k) = sum_d P[v][d] B[d][k],    P[v][d] = sqrt(w_v) s_v^d,
// i.e. two m8n8k4 DMMAs per 8 x 8 block of G instead of 8 x 8 Horner chains fed by broadcast loads of the row: the sweep
// with DFMAs was bound by those loads (15 shared-memory wavefronts per column against 10 clocks of FP64-pipe time) and ran
// the FP64 pipe at 50 %.  A warp owns a quarter of the tile's visibilities (2 groups of 8: its A fragments, the powers,
// stay in registers) and every fourth mode tile.
template <bool DEBRIS>
__device__ __forceinline__ void j0_gemm(const GramArgs &p, const uint32_t sbase, const int buf, const int ntile,
                                        const int warp, const int lane, const int last_row)
{
    const int qv = (warp & 3) * 16;                                       // this warp's quarter of the tile: visibilities qv .. qv + 15
    const int lr = lane >> 2, lk = lane & 3;
    const uint32_t s_vis = sbase + SMB_VIS + buf * (GS * 32);
    const uint32_t s_p = sbase + SMB_P + buf * (2 * GS * 32) + ((qv + lr) * 4 + lk) * 8;
    const double pa00 = lds_f64(s_p), pa01 = lds_f64(s_p + GS * 32);                      // group 0: degrees 0-3 | 4-7
    const double pa10 = lds_f64(s_p + 8 * 32), pa11 = lds_f64(s_p + GS * 32 + 8 * 32);    // group 1
    double kq0 = 0.0, kq1 = 0.0;
    if (DEBRIS) {
        kq0 = lds_f64(s_vis + (qv + lr) * 32 + 16); kq1 = lds_f64(s_vis + (qv + 8 + lr) * 32 + 16);
        kq0 = -kq0 * kq0; kq1 = -kq1 * kq1;                               // -kz^2
    }
    const uint32_t s_cm = sbase + SMB_CM + buf * (2 * GCOLS * 32) + (lr * 4 + lk) * 8;
    // The accumulator fragment holds (visibility lr, modes 2 lk, 2 lk + 1).  Rows of G are 8 banks apart (GLD = 4 mod 16
    // doubles, what the DMMA phase's fragment loads want), mode pairs 16: stored as they come, lanes lk = 0, 2 (and 1, 3)
    // would hit the same banks.  So a store instruction takes the even mode from lanes lk < 2 and the odd one from lanes
    // lk >= 2, and a second one the rest: conflict free.
    const bool lo = lk < 2;
    const uint32_t s_g0 = sbase + SMB_G + ((2 * lk + (lo ? 0 : 1)) * GLD + qv + lr) * 8;      // first store: even mode | odd mode
    const uint32_t s_g1 = sbase + SMB_G + ((2 * lk + (lo ? 1 : 0)) * GLD + qv + lr) * 8;      // second store: the other one

    // Columns to redo per visibility from the table (the tile's range of arguments needs more than one row: sparse data,
    // very large j_k) and the data column: items (column, half of the tile) dealt to the warps, a lane per visibility.
    // Done first: the latency of their table loads hides behind the DMMAs of the other warps (done last, even on the warps
    // that idle while the others prepare the next tile, they cost 0.6 % more).
    int nl;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(nl) : "r"(sbase + SMB_NLIST + buf * 4));
    for (int item = warp; item < 2 * nl; item += NW) {
        unsigned short c16;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(c16) : "r"(sbase + SMB_LIST + (buf * GCOLS + (item >> 1)) * 2));
        const int c = c16, v = (item & 1) * 32 + lane;
        const double jk = lds_f64(sbase + SMB_COL + c * 16);
        const double2 aw = lds_v2f64(s_vis + v * 32);
        double g;
        if (jk >= 0.0) {
            g = j0_tab(__dmul_rn(aw.x, jk), p.tab, last_row) * aw.y;
            if (DEBRIS) { const double kz = lds_f64(s_vis + v * 32 + 16); g *= exp_neg(-kz * kz * lds_f64(sbase + SMB_H2 + c * 8)); }
        } else {
            g = lds_f64(s_vis + v * 32 + 24);                              // data column: G[N][v] = sqrt(w_v) Re V_v
        }
        sts_f64(sbase + SMB_G + (c * GLD + v) * 8, g);
    }
    // A warp takes every fourth mode tile of its quarter.  The next tile's operands are fetched right behind the DMMAs of the
    // current one, before its results are stored, so a warp waits for the latency of its own DMMAs only.  Columns on the list
    // (below) carry NaN coefficients (prepare): their results are NaN for every visibility and are not stored.
    int t = warp >> 2;
    if (t < ntile) {
        double bN0 = lds_f64(s_cm + t * (8 * 32)), bN1 = lds_f64(s_cm + GCOLS * 32 + t * (8 * 32));
#pragma unroll 1
        for (;;) {
            const int tn = t + NW / 4;
            double d0[2] = {0.0, 0.0}, d1[2] = {0.0, 0.0};
            dmma(d0, pa01, bN1); dmma(d1, pa11, bN1);
            dmma(d0, pa00, bN0); dmma(d1, pa10, bN0);
            if (tn < ntile) { bN0 = lds_f64(s_cm + tn * (8 * 32)); bN1 = lds_f64(s_cm + GCOLS * 32 + tn * (8 * 32)); }
            if (DEBRIS) {
                const double2 h2 = lds_v2f64(sbase + SMB_H2 + (t * 8 + 2 * lk) * 8);
                d0[0] *= exp_neg(kq0 * h2.x); d0[1] *= exp_neg(kq0 * h2.y);
                d1[0] *= exp_neg(kq1 * h2.x); d1[1] *= exp_neg(kq1 * h2.y);
            }
            const uint32_t off = t * (8 * GLD * 8);
            const double x0 = lo ? d0[0] : d0[1], y0 = lo ? d0[1] : d0[0];
            const double x1 = lo ? d1[0] : d1[1], y1 = lo ? d1[1] : d1[0];
            // (NaN test on the exponent bits: an FP64 compare would queue behind the other warps' DMMAs)
            if ((__double2hiint(x0) & 0x7ff00000) != 0x7ff00000) { sts_f64(s_g0 + off, x0); sts_f64(s_g0 + off + 64, x1); }
            if ((__double2hiint(y0) & 0x7ff00000) != 0x7ff00000) { sts_f64(s_g1 + off, y0); sts_f64(s_g1 + off + 64, y1); }
            if (tn >= ntile) break;
            t = tn;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// One work item: per tile, the J0 phase (design-matrix tile into shared memory), then the DMMA phase (Gram update)
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
};

// DMMAs of one tile for one warp.  G: generic pointer to the tile; fr: this lane's fragment offset inside a tile row;
// ta / tb: first local tile index of the A rows / B columns of this warp's register block.
template <int KIND, int NR, int NC>
__device__ __forceinline__ void stage_dmma(double (&acc)[NR * NC > 0 ? NR * NC : 1][2], const double *__restrict__ G,
                                           const int fr, const int ta, const int tb, const int cbase, const int nmod)
{
    if (NR * NC == 0) return;
    const double *A = G + ta * 8 * GLD + fr;
    const double *B = G + tb * 8 * GLD + fr;
    const double *Gf = G + fr;
#pragma unroll 2
    for (int ks = 0; ks < GKS; ks++) {
        double af[NR > 0 ? NR : 1];
#pragma unroll
        for (int r = 0; r < NR; r++) af[r] = A[r * 8 * GLD + ks * 4];
        if (KIND == FB_KIND_OFF) {
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const double bf = B[c * 8 * GLD + ks * 4];
#pragma unroll
                for (int r = 0; r < NR; r++) dmma(acc[r * NC + c], af[r], bf);
            }
        } else {
#pragma unroll
            for (int s = 0; s < NR + NC - 1; s++) {
                int ct = cbase + s;
                ct = ct >= nmod ? ct - nmod : ct;
                const double bf = Gf[ct * 8 * GLD + ks * 4];
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
    const int fr = (lane >> 2) * GLD + (lane & 3);                 // fragment offset of this lane inside a tile row
    const int ta = it.r0, tb = KIND == FB_KIND_OFF ? it.ncolA / 8 + it.c0 : 0;
    const int cbase = KIND == FB_KIND_OFF ? 0 : (it.r0 + it.c0) % it.nmod;
    const uint32_t sbase = it.sbase;
    const int col = tid - 256;                                     // staging duty: column `col` of the block

    // Everything the J0 phase of a tile needs is prepared one to two tiles ahead, in three steps.
    // select(ar)        -- thread 256 + c, J0 phase of tile s: the table row for column c over tile s + 2 (arguments
    //                      [ar.x j_k, ar.y j_k]).
    // stage_copy(...)   -- all threads, DMMA phase of tile s, cp.async only (no registers held, nothing for the FP64 pipe):
    //                      the chosen rows (four lanes copy the four 16-byte pieces of a row, so a warp-wide copy touches 8
    //                      rows, not 32), the per-visibility scalars (a, sqrt w, kz, sqrt w Re V) of tile s + 2, the
    //                      range (min a, max a) of tile s + 3.
    // prepare(ar, ...)  -- J0 phase of tile s, for tile s + 1: thread 256 + c shifts the staged row's polynomial of column
    //                      c to the tile centre (B); thread v < 64 forms the powers of visibility v's tile variable (P).
    auto select = [&](const double2 ar) {
        if (col >= 0 && col < it.ncol) {
            int m = 0;
            const double jk = lds_f64(sbase + SMB_COL + col * 16);
            if (jk >= 0.0) {
                const double xlo = __dmul_rn(ar.x, jk), xhi = __dmul_rn(ar.y, jk);
                m = min(__double2loint(fma(xlo + xhi, 0.5 * FB_J0_INVH, J0_MAGIC)), last_row);
            }
            asm volatile("st.shared.s32 [%0], %1;" ::"r"(sbase + SMB_JK + col * 4), "r"(m) : "memory");
        }
    };
    auto prepare = [&](const double2 ar, const int vbuf, const int buf) {
        const double ac = 0.5 * (ar.x + ar.y);
        if (col >= 0 && col < it.ncol) {
            double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0, c4 = 0.0, c5 = 0.0, c6 = 0.0, c7 = 0.0;
            const double2 jr = lds_v2f64(sbase + SMB_COL + col * 16);                  // (j_k | -1 data column | -2 padding, j_k / j_ref)
            bool listed = jr.x == -1.0;
            if (jr.x >= 0.0) {
                int m;
                asm volatile("ld.shared.s32 %0, [%1];" : "=r"(m) : "r"(sbase + SMB_JK + col * 4));
                const double cen = (double)m * FB_J0_H;
                const double xlo = __dmul_rn(ar.x, jr.x), xhi = __dmul_rn(ar.y, jr.x);
                const bool ok = fabs(xlo - cen) < FB_J0_ACCEPT && fabs(xhi - cen) < FB_J0_ACCEPT;
                listed = !ok;
                if (ok) {
                    const uint32_t rw = sbase + SMB_ROW + col * 16;
                    const double2 r01 = lds_v2f64(rw), r23 = lds_v2f64(rw + GCOLS * 16), r45 = lds_v2f64(rw + 2 * GCOLS * 16),
                                  r67 = lds_v2f64(rw + 3 * GCOLS * 16);
                    c0 = r01.x; c1 = r01.y; c2 = r23.x; c3 = r23.y; c4 = r45.x; c5 = r45.y; c6 = r67.x; c7 = r67.y;
                    // a j_k - cen = s r + e: shift the polynomial by e (repeated synthetic division) ...
                    const double e = fma(ac, jr.x, -cen);
#define FB_SHIFT_PASS(lo)                                          \
                    c6 = fma(e, c7, c6);                           \
                    if (lo <= 5) c5 = fma(e, c6, c5);              \
                    if (lo <= 4) c4 = fma(e, c5, c4);              \
                    if (lo <= 3) c3 = fma(e, c4, c3);              \
                    if (lo <= 2) c2 = fma(e, c3, c2);              \
                    if (lo <= 1) c1 = fma(e, c2, c1);              \
                    if (lo <= 0) c0 = fma(e, c1, c0);
                    FB_SHIFT_PASS(0) FB_SHIFT_PASS(1) FB_SHIFT_PASS(2) FB_SHIFT_PASS(3) FB_SHIFT_PASS(4) FB_SHIFT_PASS(5) FB_SHIFT_PASS(6)
#undef FB_SHIFT_PASS
                    // ... and rescale to the tile variable: b_d = c_d r^d
                    const double r = jr.y, r2 = r * r, r4 = r2 * r2;
                    c1 *= r; c2 *= r2; c3 *= r2 * r; c4 *= r4; c5 *= r4 * r; c6 *= r4 * r2; c7 *= r4 * (r2 * r);
                }
            }
            if (listed) {                                                   // not served by the tile's DMMAs: NaN marks the column (j0_gemm)
                c0 = __longlong_as_double(0x7ff8000000000000LL);
                int idx;
                asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(idx) : "r"(sbase + SMB_NLIST + buf * 4) : "memory");
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(sbase + SMB_LIST + (buf * GCOLS + idx) * 2), "h"((unsigned short)col) : "memory");
            }
            const uint32_t cm = sbase + SMB_CM + buf * (2 * GCOLS * 32) + col * 32;
            sts_v2f64(cm, c0, c1); sts_v2f64(cm + 16, c2, c3);
            sts_v2f64(cm + GCOLS * 32, c4, c5); sts_v2f64(cm + GCOLS * 32 + 16, c6, c7);
        }
        if (tid < GS) {
            const double2 aw = lds_v2f64(sbase + SMB_VIS + vbuf * (GS * 32) + tid * 32);
            const double sv = aw.y != 0.0 ? (aw.x - ac) * lds_f64(sbase + SMB_JREF) : 0.0;            // padding slots: a = 0, w = 0
            const double s2 = sv * sv, s4 = s2 * s2;
            const double p0 = aw.y, p1 = p0 * sv, p2 = p0 * s2, p3 = p1 * s2;
            const uint32_t pp = sbase + SMB_P + buf * (2 * GS * 32) + tid * 32;
            sts_v2f64(pp, p0, p1); sts_v2f64(pp + 16, p2, p3);
            sts_v2f64(pp + GS * 32, p0 * s4, p1 * s4); sts_v2f64(pp + GS * 32 + 16, p2 * s4, p3 * s4);
        }
    };
    // rows: copy the rows chosen by select(); qv >= 0: scalars of tile qv into buffer vbuf; qa >= 0: range of tile qa into ring slot
    auto stage_copy = [&](const bool rows, const long long qv, const int vbuf, const long long qa, const int slot) {
        if (rows) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int c = (tid >> 2) + 128 * h, piece = tid & 3;
                if (c < it.ncol) {
                    int m;
                    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(m) : "r"(sbase + SMB_JK + c * 4));
                    const double2 *src = p.tab + (size_t)m * (FB_J0_ROWLEN / 2) + piece;
                    const uint32_t dst = sbase + SMB_ROW + piece * (GCOLS * 16) + c * 16;
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                }
            }
        }
        if (qv >= 0 && tid < GS) {
            const size_t v = (size_t)qv * GS + tid;
            const uint32_t d = sbase + SMB_VIS + vbuf * (GS * 32) + tid * 32;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(p.a + v) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d + 8), "l"(p.sw + v) : "memory");
            if (DEBRIS) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d + 16), "l"(p.kz + v) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d + 24), "l"(p.swV + v) : "memory");
        }
        if (qa >= 0 && tid == GS)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sbase + SMB_AR + slot * 16), "l"(p.arange + qa) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto ring = [&](const int s) { return lds_v2f64(sbase + SMB_AR + (s & 3) * 16); };

    const long long q0 = it.q0, qs = it.qs;
    const int nst = it.nst;
    auto tile = [&](const int s) { return s < nst ? q0 + s * qs : -1LL; };
    // ---- prologue: tile 0 prepared, rows of tile 1 chosen and on their way -----------------------------------------
    {
        const double2 ar0 = p.arange[q0];
        select(ar0);
        if (tid == 0) asm volatile("st.shared.v2.s32 [%0], {%1, %1};" ::"r"(sbase + SMB_NLIST), "r"(0) : "memory");
        __syncthreads();
        stage_copy(true, tile(0), 0, tile(1), 1);
        stage_copy(false, tile(1), 1, tile(2), 2);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        prepare(ar0, 0, 0);
        if (nst > 1) select(ring(1));
        __syncthreads();
        stage_copy(nst > 1, -1, 0, -1, 0);
    }
    // ---- main loop: J0 phase, DMMA phase ------------------------------------------------------------------------
#ifdef FB_GRAM_CLOCKS
    long long t_j0 = 0, t_mma = 0, t_g0 = 0, t_last = clock64();
#endif
    for (int s = 0; s < nst; s++) {
        const int b = s & 1;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                     // rows of tile s + 1 and scalars of tile s + 2 staged; DMMAs of tile s - 1 done
#ifdef FB_GRAM_CLOCKS
        if (p.prof && tid == 0) { const long long t = clock64(); t_mma += t - t_last; t_last = t; }
#endif
        // the tile's DMMAs first: the per-column DFMA chains of prepare() crawl while other warps stream DMMAs (one FP64
        // pipe serves both: run first, they stretched the phase by a third), so they go last, on an idle pipe
        j0_gemm<DEBRIS>(p, sbase, b, it.ncol / 8, tid >> 5, lane, last_row);
#ifdef FB_GRAM_CLOCKS
        if (p.prof && tid == 0) t_g0 += clock64() - t_last;
#endif
        if (s + 1 < nst) {
            prepare(ring(s + 1), b ^ 1, b ^ 1);
            if (s + 2 < nst) select(ring(s + 2));
        }
        __syncthreads();
#ifdef FB_GRAM_CLOCKS
        if (p.prof && tid == 0) { const long long t = clock64(); t_j0 += t - t_last; t_last = t; }
#endif
        if (tid == 0) asm volatile("st.shared.s32 [%0], %1;" ::"r"(sbase + SMB_NLIST + b * 4), "r"(0) : "memory");   // refilled by the next phase's prepare()
        stage_copy(s + 2 < nst, tile(s + 2), b, tile(s + 3), (s + 3) & 3);     // flies during the DMMAs
        stage_dmma<KIND, NR, NC>(acc, it.G, fr, ta, tb, cbase, it.nmod);
    }
#ifdef FB_GRAM_CLOCKS
    if (p.prof && tid == 0) {
        t_mma += clock64() - t_last;
        atomicAdd((unsigned long long *)&p.prof[2 * blockIdx.x], (unsigned long long)t_j0);
        atomicAdd((unsigned long long *)&p.prof[2 * blockIdx.x + 1], (unsigned long long)t_mma);
        atomicAdd((unsigned long long *)&p.prof[2 * gridDim.x], (unsigned long long)t_g0);
    }
#endif
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
        {
            const int last = ty.kind == FB_KIND_OFF ? max(ty.a_t0 + ty.a_nt, ty.b_t0 + ty.b_nt) : ty.a_t0 + ty.a_nt;
            const double jref = p.jk[min(p.N, last * 8) - 1];
            if (tid == 0) sts_f64(sbase + SMB_JREF, jref);
            if (tid < it.ncol) {
                const int g = tid < it.ncolA ? ty.a_t0 * 8 + tid : ty.b_t0 * 8 + (tid - it.ncolA);
                double v = -2.0, h = 0.0;
                if (g < p.N) { v = p.jk[g]; if (DEBRIS) h = p.H2[g]; }
                else if (g == p.N) v = -1.0;
                sts_v2f64(sbase + SMB_COL + tid * 16, v, v >= 0.0 ? v / jref : 0.0);
                if (DEBRIS) sts_f64(sbase + SMB_H2 + tid * 8, h);
            }
        }
        __syncthreads();                      // the column tables are read by other threads (select / prepare)

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

// Cost of a block per tile of FB_TV visibilities, in FP64-pipe clocks of the busiest SM sub-partition: a DMMA holds the
// pipe for 16 clocks; a column of the design-matrix tile costs 8 (two DMMAs per 8 x 8 block, a quarter of the tile's
// visibilities per sub-partition); re-centring the polynomials, the barriers and the ramps cost about 1500 per tile
// (measured with the in-kernel clocks; the split between the block types is flat within 0.5 % around these values).
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
    static const double j0w = getenv("FB_J0_COST") ? atof(getenv("FB_J0_COST")) : 8.0;
    static const double j0f = getenv("FB_J0_FIXED") ? atof(getenv("FB_J0_FIXED")) : 1500.0;
    return 16.0 * (FB_TV / 4) * worst + j0w * cols + j0f;
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
#ifdef FB_GRAM_CLOCKS
    static const bool prof = getenv("FB_GRAM_PROF") != nullptr;
#else
    const bool prof = false;
#endif
    if (prof) {
        FB_CUDA(cudaMalloc(&args.prof, sizeof(long long) * (2 * grid + 1)));
        FB_CUDA(cudaMemsetAsync(args.prof, 0, sizeof(long long) * (2 * grid + 1), ln.stream));
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
        std::vector<long long> h(2 * grid + 1);
        FB_CUDA(cudaStreamSynchronize(ln.stream));
        FB_CUDA(cudaMemcpy(h.data(), args.prof, sizeof(long long) * (2 * grid + 1), cudaMemcpyDeviceToHost));
        long long j0_sum = 0, mma_sum = 0, j0_max = 0, mma_max = 0, tot_max = 0, tot_min = -1;
        for (int b = 0; b < grid; b++) {
            j0_sum += h[2 * b]; mma_sum += h[2 * b + 1];
            j0_max = std::max(j0_max, h[2 * b]); mma_max = std::max(mma_max, h[2 * b + 1]);
            tot_max = std::max(tot_max, h[2 * b] + h[2 * b + 1]);
            tot_min = tot_min < 0 ? h[2 * b] + h[2 * b + 1] : std::min(tot_min, h[2 * b] + h[2 * b + 1]);
        }
        fprintf(stderr, "[fb_gram prof] chan=%d  per-CTA clocks: J0 avg %.0f max %lld (warp 0's own J0 DMMAs %.0f) | DMMA avg %.0f max %lld | total min %lld max %lld\n",
                chan, (double)j0_sum / grid, j0_max, (double)h[2 * grid] / grid, (double)mma_sum / grid, mma_max, tot_min, tot_max);
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
