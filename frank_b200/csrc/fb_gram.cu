// K3: fused design-matrix + Gram kernel.
//
// Replaces the reference's chunk loop (frank/statistical_models.py:192-214)
//     X   = H(q_chunk)                      DHT.coefficients, frank/hankel.py:187-204
//     wXT = X.T * w ;  M += wXT @ X ;  j += wXT @ V
// with one persistent kernel in which the design-matrix tile never leaves the SM:
//   * the symmetric (N+1)x(N+1) matrix  S = G^T G,
//         G[i, k] = sqrt(w_i) J0(a_i j_k)  (k < N),   G[i, N] = sqrt(w_i) Re V_i,
//     gives M = diag(c) S[:N,:N] diag(c), j = diag(c) S[:N, N], with c_k = norm * scale_factor_k * scale;
//   * a CTA owns one block of S (work-item types OFF / DIAG, fb_common.cuh) with its accumulators in registers
//     (<= 15 m8n8 tiles = 60 registers per thread, 16 warps);
//   * per tile of 64 visibilities: (1) stage, per column, the row of the J0 Taylor table the tile needs (the
//     visibilities are sorted by baseline so a whole warp shares it); (2) all warps evaluate J0 for the block's
//     columns into shared memory G[mode][vis] (two visibilities per lane, coefficients broadcast from shared
//     memory; |err| <= 0.5 ulp + 1e-17); (3) all warps run mma.sync.m8n8k4.f64 (DMMA) over the tile;
//   * diagonal blocks are executed as skewed strips (row r, offset d -> column (r+d) mod n) so that only the
//     upper triangle is computed while every warp still owns a dense register block;
//   * work items (block, visibility chunk) write partial blocks; a second kernel sums the chunks in a fixed
//     order (deterministic), applies c_k c_l and mirrors the triangle.
#include "fb_common.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace {

struct GramArgs {
    const double *a, *sw, *swV, *kz;
    long long n_tiles;
    const double *jk;
    const double2 *tab;
    int tab_rows;
    int N, n_items;
    const FbGramType *types;
    const int *work;      // [3 * n_items] (type, chunk, partial slot), then [2 * ntypes] (C_t, first slot)
    int work_types_off;   // offset of the per-type table inside work
    const double *H2;
    double *partial;
    int debug_mode;       // 0 normal | 1 skip the DMMA phase | 2 skip J0 evaluation after the first tile (profiling aid)
};

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

// J0(x), x >= 0: row m = round(4x) holds the Taylor coefficients about m/4, |t| <= 1/8.  Generic (gather) path.
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
__device__ __forceinline__ int lds_s32(uint32_t addr)
{
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f64(uint32_t addr, double v)
{
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}

constexpr int GT = FB_TV;              // visibilities per tile (64)
constexpr int GLD = FB_LDV;            // 68 = 4 mod 16: conflict-free m8n8k4 fragments
constexpr int GKS = GT / 4;            // k-steps per tile
constexpr int GCOLS = FB_GCOLS;        // columns held per tile (rows' columns | columns' columns)

// shared-memory carve-up (in doubles)
constexpr int SM_G = 0;                                   // [GCOLS][GLD]        design-matrix tile
constexpr int SM_JK = SM_G + GCOLS * GLD;                 // [GCOLS]             j_k (>= 0), -1 data column, -2 padding
constexpr int SM_H2 = SM_JK + GCOLS;                      // [GCOLS]             debris H2_k
constexpr int SM_ROW = SM_H2 + GCOLS;                     // [2][GCOLS][10]      staged J0 table rows (double buffered)
constexpr int SM_ROWM = SM_ROW + 2 * GCOLS * FB_J0_ROWLEN;   // [2][GCOLS] int   staged row index
constexpr int SM_DOUBLES = SM_ROWM + GCOLS;
constexpr size_t GRAM_SMEM_BYTES = sizeof(double) * SM_DOUBLES;

__host__ __device__ __forceinline__ int split4_size(int n, int i) { return n / 4 + (i < n % 4 ? 1 : 0); }
__host__ __device__ __forceinline__ int split4_start(int n, int i) { return i * (n / 4) + (i < n % 4 ? i : n % 4); }

// rectangular register block: acc[r * NC + c] += G_A[r0+r]^T G_B[c0+c] over one tile
template <int NR, int NC>
__device__ __forceinline__ void mma_off(double (&acc)[FB_ACC][2], const double *__restrict__ ap, const double *__restrict__ bp)
{
    static_assert(NR * NC <= FB_ACC, "register block too large");
#pragma unroll 2
    for (int ks = 0; ks < GKS; ks++) {
        double af[NR];
#pragma unroll
        for (int r = 0; r < NR; r++) af[r] = ap[r * 8 * GLD + ks * 4];
#pragma unroll
        for (int c = 0; c < NC; c++) {
            const double b = bp[c * 8 * GLD + ks * 4];
#pragma unroll
            for (int r = 0; r < NR; r++) dmma(acc[r * NC + c], af[r], b);
        }
    }
}

// skewed strip of a triangle: acc[r * ND + d] += G[r0+r]^T G[(r0+r + d0+d) mod n] over one tile
template <int NR, int ND>
__device__ __forceinline__ void mma_diag(double (&acc)[FB_ACC][2], const double *__restrict__ ap, const double *__restrict__ tp,
                                         int cbase, int nmod)
{
    static_assert(NR * ND <= FB_ACC, "register block too large");
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
                if (d >= 0 && d < ND) dmma(acc[r * ND + d], af[r], b);
            }
        }
    }
}

template <int NR, int NC>
__device__ __forceinline__ void store_block(const double (&acc)[FB_ACC][2], double *__restrict__ out, int r0, int c0, int ld, int lane)
{
#pragma unroll
    for (int r = 0; r < NR; r++)
#pragma unroll
        for (int c = 0; c < NC; c++)
            *reinterpret_cast<double2 *>(out + ((size_t)((r0 + r) * ld + (c0 + c)) * 64 + lane * 2)) =
                make_double2(acc[r * NC + c][0], acc[r * NC + c][1]);
}

// compile-time dispatch over the register-block shape (NR * NC <= FB_ACC)
#define FB_DISPATCH_SHAPE(nr, nc, CALL)                                               \
    switch ((nr) * 8 + (nc)) {                                                        \
        case 1 * 8 + 1: { CALL(1, 1); } break; case 1 * 8 + 2: { CALL(1, 2); } break; \
        case 1 * 8 + 3: { CALL(1, 3); } break; case 1 * 8 + 4: { CALL(1, 4); } break; \
        case 1 * 8 + 5: { CALL(1, 5); } break; case 2 * 8 + 1: { CALL(2, 1); } break; \
        case 2 * 8 + 2: { CALL(2, 2); } break; case 2 * 8 + 3: { CALL(2, 3); } break; \
        case 2 * 8 + 4: { CALL(2, 4); } break; case 2 * 8 + 5: { CALL(2, 5); } break; \
        case 3 * 8 + 1: { CALL(3, 1); } break; case 3 * 8 + 2: { CALL(3, 2); } break; \
        case 3 * 8 + 3: { CALL(3, 3); } break; case 3 * 8 + 4: { CALL(3, 4); } break; \
        case 3 * 8 + 5: { CALL(3, 5); } break; case 4 * 8 + 1: { CALL(4, 1); } break; \
        case 4 * 8 + 2: { CALL(4, 2); } break; case 4 * 8 + 3: { CALL(4, 3); } break; \
        case 5 * 8 + 1: { CALL(5, 1); } break; case 5 * 8 + 2: { CALL(5, 2); } break; \
        case 5 * 8 + 3: { CALL(5, 3); } break;                                        \
        default: break;                                                               \
    }

// Evaluate the staged polynomial for two visibilities of one column (shared coefficient loads, all five
// coefficient pairs requested before the first use).
__device__ __forceinline__ void horner2(const double2 *__restrict__ rb, double t0, double t1, double &g0, double &g1)
{
    const double2 c4 = rb[4], c3 = rb[3], c2 = rb[2], c1 = rb[1], c0 = rb[0];
    g0 = fma(c4.y, t0, c4.x);         g1 = fma(c4.y, t1, c4.x);
    g0 = fma(g0, t0, c3.y);           g1 = fma(g1, t1, c3.y);
    g0 = fma(g0, t0, c3.x);           g1 = fma(g1, t1, c3.x);
    g0 = fma(g0, t0, c2.y);           g1 = fma(g1, t1, c2.y);
    g0 = fma(g0, t0, c2.x);           g1 = fma(g1, t1, c2.x);
    g0 = fma(g0, t0, c1.y);           g1 = fma(g1, t1, c1.y);
    g0 = fma(g0, t0, c1.x);           g1 = fma(g1, t1, c1.x);
    g0 = fma(g0, t0, c0.y);           g1 = fma(g1, t1, c0.y);
    g0 = fma(g0, t0, c0.x);           g1 = fma(g1, t1, c0.x);
}

template <bool DEBRIS>
__global__ void __launch_bounds__(FB_GRAM_THREADS, 1) k_gram(GramArgs p)
{
    extern __shared__ double smem[];
    double *G = smem + SM_G;
    double *col_jk = smem + SM_JK;
    double *col_h2 = smem + SM_H2;
    double *rowbuf = smem + SM_ROW;
    int *rowm = reinterpret_cast<int *>(smem + SM_ROWM);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int last_row = p.tab_rows - 1;
    const double MAGIC = 6755399441055744.0;
    const uint32_t s_G = (uint32_t)__cvta_generic_to_shared(G), s_jk = (uint32_t)__cvta_generic_to_shared(col_jk);
    const uint32_t s_row = (uint32_t)__cvta_generic_to_shared(rowbuf), s_rowm = (uint32_t)__cvta_generic_to_shared(rowm);

    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int type = p.work[3 * item], chunk = p.work[3 * item + 1], slot_out = p.work[3 * item + 2];
        const int Ct = p.work[p.work_types_off + 2 * type];
        const FbGramType ty = p.types[type];
        const int t0 = (int)((p.n_tiles * chunk) / Ct), t1 = (int)((p.n_tiles * (chunk + 1)) / Ct);
        const int ncolA = ty.a_nt * 8, ncol = (ty.a_nt + ty.b_nt) * 8;

        __syncthreads();
        for (int lc = tid; lc < ncol; lc += FB_GRAM_THREADS) {
            const int g = lc < ncolA ? ty.a_t0 * 8 + lc : ty.b_t0 * 8 + (lc - ncolA);
            col_jk[lc] = g < p.N ? p.jk[g] : (g == p.N ? -1.0 : -2.0);
            if (DEBRIS) col_h2[lc] = g < p.N ? p.H2[g] : 0.0;
        }

        double acc[FB_ACC][2];
#pragma unroll
        for (int i = 0; i < FB_ACC; i++) acc[i][0] = acc[i][1] = 0.0;

        // this warp's register block (warp-uniform).  Latin-square assignment of (row group, column group) to
        // (slot, sub-partition) balances the DMMA count of the four SM sub-partitions.
        const int smsp = warp & 3, slot = warp >> 2;
        const int kind = ty.kind;
        int r0, c0, nr, nc, nmod;
        if (kind == FB_KIND_OFF) {
            nmod = 1;
            r0 = split4_start(ty.a_nt, slot);
            nr = split4_size(ty.a_nt, slot);
            c0 = split4_start(ty.b_nt, (slot + smsp) & 3);
            nc = split4_size(ty.b_nt, (slot + smsp) & 3);
        } else {
            nmod = ty.a_nt;
            const int D = nmod / 2 + 1;
            r0 = split4_start(nmod, slot);
            nr = split4_size(nmod, slot);
            c0 = split4_start(D, (slot + smsp) & 3);       // first skew offset d
            nc = split4_size(D, (slot + smsp) & 3);
        }
        if (nr == 0 || nc == 0) { nr = 0; nc = 0; }
        const int frag = (lane >> 2) * GLD + (lane & 3);
        const double *ap = G + (r0 * 8) * GLD + frag;
        const double *bp = kind == FB_KIND_OFF ? G + (ncolA + c0 * 8) * GLD + frag : G + frag;
        const int cbase = (r0 + c0) % nmod;

        // Software pipeline over tiles: while the DMMAs of tile t run, the table rows of tile t+1 stream into the
        // other half of rowbuf (cp.async) and its per-visibility scalars into registers.
        // stage_rows(tile, buf): thread tid < ncol copies the row its column needs at the tile's first visibility.
        auto stage_rows = [&](double aref, int buf) {
            if (tid < ncol) {
                const double jk = col_jk[tid];
                int m = 0;
                if (jk >= 0.0) m = min(__double2loint(fma(__dmul_rn(aref, jk), 4.0, MAGIC)), last_row);
                rowm[buf * GCOLS + tid] = m;
                const double2 *row = p.tab + (size_t)m * (FB_J0_ROWLEN / 2);
                const uint32_t dst = s_row + (buf * GCOLS + tid) * (FB_J0_ROWLEN * 8);
#pragma unroll
                for (int k = 0; k < FB_J0_ROWLEN / 2; k++)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + 16 * k), "l"(row + k) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        double na0 = 0.0, na1 = 0.0, nw0 = 0.0, nw1 = 0.0, nk0 = 0.0, nk1 = 0.0, aref_next = 0.0;
        if (t0 < t1) {
            const size_t v0 = (size_t)t0 * GT;
            stage_rows(p.a[v0], 0);
            na0 = p.a[v0 + lane]; na1 = p.a[v0 + lane + 32];
            nw0 = p.sw[v0 + lane]; nw1 = p.sw[v0 + lane + 32];
            if (DEBRIS) { nk0 = p.kz[v0 + lane]; nk1 = p.kz[v0 + lane + 32]; }
            if (t0 + 1 < t1) aref_next = p.a[v0 + GT];
        }

        for (int tile = t0; tile < t1; ++tile) {
            const size_t v0 = (size_t)tile * GT;
            const int buf = (tile - t0) & 1;
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();          // rows of this tile staged; every warp is done with the previous tile's DMMAs
            // ---- (2) J0 evaluation ------------------------------------------------------------------
            if (p.debug_mode != 2 || tile == t0) {
                const double a0 = na0, a1 = na1, w0 = nw0, w1 = nw1;
                double k0 = 0.0, k1 = 0.0;
                if (DEBRIS) { k0 = -nk0 * nk0; k1 = -nk1 * nk1; }
                // One column per iteration, two visibilities per lane.  The column's j_k and staged row
                // index are fetched one iteration ahead; the polynomial is evaluated from the staged row
                // unconditionally and redone through the gather path only if some lane needs another row,
                // so neither the loads nor the vote sit on the dependency chain.
                uint32_t a_jk = s_jk + warp * 8, a_rm = s_rowm + (buf * GCOLS + warp) * 4;
                uint32_t a_rb = s_row + (buf * GCOLS + warp) * (FB_J0_ROWLEN * 8);
                uint32_t a_g = s_G + (warp * GLD + lane) * 8;
                double jk_n = 0.0;
                int mref_n = 0;
                if (warp < ncol) { jk_n = lds_f64(a_jk); mref_n = lds_s32(a_rm); }
#pragma unroll 1
                for (int sc = warp; sc < ncol; sc += FB_GRAM_THREADS / 32) {
                    const double jk = jk_n;
                    const int mref = mref_n;
                    if (sc + FB_GRAM_THREADS / 32 < ncol) {
                        jk_n = lds_f64(a_jk + (FB_GRAM_THREADS / 32) * 8);
                        mref_n = lds_s32(a_rm + (FB_GRAM_THREADS / 32) * 4);
                    }
                    const double2 c4 = lds_v2f64(a_rb + 64), c3 = lds_v2f64(a_rb + 48), c2 = lds_v2f64(a_rb + 32),
                                  c1 = lds_v2f64(a_rb + 16), c0 = lds_v2f64(a_rb);
                    const double x0 = __dmul_rn(a0, jk), x1 = __dmul_rn(a1, jk);
                    const double s0 = fma(x0, 4.0, MAGIC), s1 = fma(x1, 4.0, MAGIC);
                    const double u0 = fma(s0 - MAGIC, -0.25, x0), u1 = fma(s1 - MAGIC, -0.25, x1);
                    double g0 = fma(c4.y, u0, c4.x), g1 = fma(c4.y, u1, c4.x);
                    g0 = fma(g0, u0, c3.y);           g1 = fma(g1, u1, c3.y);
                    g0 = fma(g0, u0, c3.x);           g1 = fma(g1, u1, c3.x);
                    g0 = fma(g0, u0, c2.y);           g1 = fma(g1, u1, c2.y);
                    g0 = fma(g0, u0, c2.x);           g1 = fma(g1, u1, c2.x);
                    g0 = fma(g0, u0, c1.y);           g1 = fma(g1, u1, c1.y);
                    g0 = fma(g0, u0, c1.x);           g1 = fma(g1, u1, c1.x);
                    g0 = fma(g0, u0, c0.y);           g1 = fma(g1, u1, c0.y);
                    g0 = fma(g0, u0, c0.x);           g1 = fma(g1, u1, c0.x);
                    const bool same = (jk < 0.0) | ((__double2loint(s0) == mref) & (__double2loint(s1) == mref));
                    if (!__all_sync(0xffffffffu, same)) {          // rare: some lane straddles a table row
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
                    if (jk < 0.0) {                                 // data column / zero padding (warp-uniform)
                        g0 = jk == -1.0 ? p.swV[v0 + lane] : 0.0;
                        g1 = jk == -1.0 ? p.swV[v0 + lane + 32] : 0.0;
                    }
                    sts_f64(a_g, g0);
                    sts_f64(a_g + 32 * 8, g1);
                    a_jk += (FB_GRAM_THREADS / 32) * 8;
                    a_rm += (FB_GRAM_THREADS / 32) * 4;
                    a_rb += (FB_GRAM_THREADS / 32) * (FB_J0_ROWLEN * 8);
                    a_g += (FB_GRAM_THREADS / 32) * GLD * 8;
                }
            }
            // ---- prefetch for tile + 1 (in flight during the DMMA phase) -------------------------------
            if (tile + 1 < t1 && p.debug_mode != 2) {
                stage_rows(aref_next, buf ^ 1);
                na0 = p.a[v0 + GT + lane]; na1 = p.a[v0 + GT + lane + 32];
                nw0 = p.sw[v0 + GT + lane]; nw1 = p.sw[v0 + GT + lane + 32];
                if (DEBRIS) { nk0 = p.kz[v0 + GT + lane]; nk1 = p.kz[v0 + GT + lane + 32]; }
                if (tile + 2 < t1) aref_next = p.a[v0 + 2 * GT];
            }
            __syncthreads();
            // ---- (3) DMMA over the tile -----------------------------------------------------------
            if (p.debug_mode != 1 && nr > 0) {
                if (kind == FB_KIND_OFF) {
#define FB_CALL_OFF(NR, NC) mma_off<NR, NC>(acc, ap, bp)
                    FB_DISPATCH_SHAPE(nr, nc, FB_CALL_OFF)
                } else {
#define FB_CALL_DIAG(NR, ND) mma_diag<NR, ND>(acc, ap, bp, cbase, nmod)
                    FB_DISPATCH_SHAPE(nr, nc, FB_CALL_DIAG)
                }
            }
        }

        // ---------------- write the partial block ------------------------------------------------
        double *out = p.partial + (size_t)slot_out * FB_PSZ;
        const int ld = kind == FB_KIND_OFF ? FB_PT : FB_DH;
#define FB_CALL_STORE(NR, NC) store_block<NR, NC>(acc, out, r0, c0, ld, lane)
        FB_DISPATCH_SHAPE(nr, nc, FB_CALL_STORE)
    }
}

// Sum the chunk partials in a fixed order, scale, mirror.  One 64-thread block per upper tile pair.
__global__ void __launch_bounds__(64)
k_gram_finalize(int N, int NT, int P, long long n_tiles, const int *__restrict__ tile_panel,
                const int *__restrict__ panel_t0, const int *__restrict__ panel_nt, const int *__restrict__ pair_code,
                const int *__restrict__ type_tab, const double *__restrict__ partial, const double *__restrict__ ck,
                double scale, double *__restrict__ M, double *__restrict__ jvec)
{
    // decode the upper-triangular tile pair (tr <= tc) from blockIdx.x
    int rem = blockIdx.x, tr = 0;
    while (rem >= NT - tr) { rem -= NT - tr; tr++; }
    const int tc = tr + rem;
    const int i = threadIdx.x >> 3, jx = threadIdx.x & 7;
    const int row = tr * 8 + i, col = tc * 8 + jx;
    if (tr == tc && i > jx) return;
    const int pa = tile_panel[tr], pb = tile_panel[tc];
    const int e_direct = (i * 4 + (jx >> 1)) * 2 + (jx & 1);        // C-fragment slot of element (i, jx)
    const int e_transp = (jx * 4 + (i >> 1)) * 2 + (i & 1);         // ... of element (jx, i)
    int type, idx;
    if (pa < pb) {
        const int lrp = tr - panel_t0[pa], h0 = (panel_nt[pa] + 1) / 2;
        const int half = lrp >= h0 ? 1 : 0;
        type = pair_code[(pa * P + pb) * 2 + half];
        idx = ((lrp - half * h0) * FB_PT + (tc - panel_t0[pb])) * 64 + e_direct;
    } else {
        type = pair_code[(pa * P + pa) * 2];
        const int n = panel_nt[pa], D = n / 2 + 1;
        const int lr = tr - panel_t0[pa], lc = tc - panel_t0[pa], d = lc - lr;
        if (d < D)
            idx = (lr * FB_DH + d) * 64 + e_direct;
        else   // stored as the transposed tile (row tile lc, offset n - d)
            idx = (lc * FB_DH + (n - d)) * 64 + e_transp;
    }
    const int Ct = type_tab[2 * type], first = type_tab[2 * type + 1];
    double s = 0.0;
    for (int c = 0; c < Ct; c++) {
        const long long t0 = (n_tiles * c) / Ct, t1 = (n_tiles * (c + 1)) / Ct;
        if (t1 > t0) s += partial[(size_t)(first + c) * FB_PSZ + idx];
    }
    if (col < N) {           // row <= col < N
        const double val = ((ck[row] * scale) * (ck[col] * scale)) * s;
        M[(size_t)row * N + col] = val;
        M[(size_t)col * N + row] = val;
    } else if (col == N && row < N) {
        jvec[row] = (ck[row] * scale) * s;
    }
}

// K8: predicted visibilities V_i = sum_k H_ik I_k, H_ik = c_k J0(a_i j_k) scale_ik   (statistical_models.py:279-329).
// One warp per visibility, lanes stride over the modes; fixed shuffle tree.
__global__ void __launch_bounds__(256)
k_predict(int64_t n, int N, const double *__restrict__ q, const double *__restrict__ kz, double invQmax,
          const double *__restrict__ jk, const double *__restrict__ ck, const double *__restrict__ Ik,
          const double *__restrict__ H2, double scale, const double2 *__restrict__ tab, int rows, double *__restrict__ V)
{
    const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const double a = __dmul_rn(q[i], invQmax);
    double kk = 0.0;
    if (H2) { kk = kz[i]; kk = -kk * kk; }
    double acc = 0.0;
    for (int k = lane; k < N; k += 32) {
        double h = ck[k] * j0_tab(__dmul_rn(a, jk[k]), tab, rows - 1);
        h *= H2 ? exp(kk * H2[k]) : scale;
        acc = fma(h, Ik[k], acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if (lane == 0) V[i] = acc;
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
    const int ntypes = ctx->ntypes;
    // Chunks per type, proportional to the type's cost per visibility on the FP64 pipe
    // (64 FMA-lanes per accumulator tile, ~14 per J0 evaluation) so that all SMs finish together.
    std::vector<double> cost(ntypes);
    double total = 0.0;
    for (int t = 0; t < ntypes; t++) {
        const FbGramType &ty = ctx->h_types[t];
        double tiles, cols;
        if (ty.kind == FB_KIND_OFF) { tiles = (double)ty.a_nt * ty.b_nt; cols = 8.0 * (ty.a_nt + ty.b_nt); }
        else { tiles = (double)ty.a_nt * (ty.a_nt / 2 + 1); cols = 8.0 * ty.a_nt; }
        cost[t] = 64.0 * tiles + 14.0 * cols;
        total += cost[t];
    }
    std::vector<int> C(ntypes, 1);
    if (ntypes < ctx->num_sms) {
        int used = 0;
        for (int t = 0; t < ntypes; t++) {
            C[t] = std::max(1, (int)std::floor(ctx->num_sms * cost[t] / total));
            used += C[t];
        }
        // hand the left-over SMs to the types with the largest cost per chunk
        while (used < ctx->num_sms) {
            int best = 0;
            for (int t = 1; t < ntypes; t++)
                if (cost[t] / C[t] > cost[best] / C[best]) best = t;
            C[best]++;
            used++;
        }
    }
    for (int t = 0; t < ntypes; t++)
        if ((long long)C[t] > n_tiles) C[t] = (int)std::max<long long>(1, n_tiles);
    std::vector<int> work;
    std::vector<int> type_tab(2 * ntypes);
    int n_items = 0;
    for (int t = 0; t < ntypes; t++) { type_tab[2 * t] = C[t]; type_tab[2 * t + 1] = n_items; n_items += C[t]; }
    work.resize(3 * (size_t)n_items);
    {
        // costliest chunks first; types interleaved
        std::vector<std::pair<double, std::pair<int, int>>> order;
        for (int t = 0; t < ntypes; t++)
            for (int c = 0; c < C[t]; c++) order.push_back({-cost[t] / C[t], {c, t}});
        std::stable_sort(order.begin(), order.end(), [](const auto &x, const auto &y) { return x.first < y.first; });
        for (int i = 0; i < n_items; i++) {
            const int t = order[i].second.second, c = order[i].second.first;
            work[3 * i] = t; work[3 * i + 1] = c; work[3 * i + 2] = type_tab[2 * t + 1] + c;
        }
    }
    const int types_off = 3 * n_items;
    work.insert(work.end(), type_tab.begin(), type_tab.end());
    if ((int)work.size() > ctx->work_cap) {
        if (ctx->d_work) FB_CUDA(cudaFree(ctx->d_work));
        ctx->d_work = nullptr;
        FB_CUDA(cudaMalloc(&ctx->d_work, sizeof(int) * (work.size() + 1024)));
        ctx->work_cap = (int)work.size() + 1024;
    }
    FB_CUDA(cudaMemcpyAsync(ctx->d_work, work.data(), sizeof(int) * work.size(), cudaMemcpyHostToDevice, ctx->stream));
    FB_CUDA(cudaStreamSynchronize(ctx->stream));     // `work` is a stack object
    const size_t need = (size_t)n_items * FB_PSZ;
    if (need > ctx->partial_cap) {
        if (ctx->d_partial) FB_CUDA(cudaFree(ctx->d_partial));
        ctx->d_partial = nullptr;
        FB_CUDA(cudaMalloc(&ctx->d_partial, need * sizeof(double)));
        ctx->partial_cap = need;
    }
    FB_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));

    GramArgs args;
    args.a = ctx->d_a; args.sw = ctx->d_sw; args.swV = ctx->d_swV; args.kz = ctx->d_kz;
    args.n_tiles = n_tiles;
    args.jk = ctx->d_jk; args.tab = ctx->d_tab; args.tab_rows = ctx->tab_rows;
    args.N = ctx->N; args.n_items = n_items;
    args.types = ctx->d_types; args.work = ctx->d_work; args.work_types_off = types_off;
    args.H2 = ctx->d_H2; args.partial = ctx->d_partial;
    {
        const char *dbg = getenv("FB_GRAM_DEBUG");
        args.debug_mode = dbg ? atoi(dbg) : 0;
    }
    int grid = std::min(n_items, ctx->num_sms);
    if (n_tiles > 0) {
        if (vis_model == FB_MODEL_DEBRIS) {
            FB_CUDA(cudaFuncSetAttribute(k_gram<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GRAM_SMEM_BYTES));
            k_gram<true><<<grid, FB_GRAM_THREADS, GRAM_SMEM_BYTES, ctx->stream>>>(args);
        } else {
            FB_CUDA(cudaFuncSetAttribute(k_gram<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GRAM_SMEM_BYTES));
            k_gram<false><<<grid, FB_GRAM_THREADS, GRAM_SMEM_BYTES, ctx->stream>>>(args);
        }
        FB_CUDA(cudaGetLastError());
    }
    FB_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));
    const int npairs = ctx->NT * (ctx->NT + 1) / 2;
    k_gram_finalize<<<npairs, 64, 0, ctx->stream>>>(ctx->N, ctx->NT, ctx->P, n_tiles, ctx->d_tile_panel, ctx->d_panel_t0,
                                                    ctx->d_panel_nt, ctx->d_pair_code, ctx->d_work + types_off,
                                                    ctx->d_partial, ctx->d_ck, model_scale, dev_M, dev_j);
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

extern "C" int fb_predict_visibilities(fb_ctx *ctx, int64_t n, const double *host_q, const double *host_kz, const double *host_I,
                                       int vis_model, double model_scale, const double *host_H2, double *host_V)
{
    if (!ctx) return -1;
    if (ctx->N == 0) FB_FAIL(-11, "fb_predict_visibilities: fb_dht_setup has not been called");
    if (n < 0 || !host_q || !host_I || !host_V) FB_FAIL(-12, "fb_predict_visibilities: bad arguments");
    if (vis_model == FB_MODEL_DEBRIS && (!host_H2 || !host_kz)) FB_FAIL(-15, "fb_predict_visibilities: debris model needs kz and H2");
    if (n == 0) return 0;
    FB_CUDA(cudaSetDevice(ctx->device));
    const int N = ctx->N;
    double qmax = 0.0;
    for (int64_t i = 0; i < n; i++) qmax = host_q[i] > qmax ? host_q[i] : qmax;
    const double xneed = qmax * ctx->invQmax * ctx->h_jk[N - 1];
    if (xneed * 4.0 + 2.0 > (double)ctx->tab_rows) {
        int rc = fb_build_j0_table(ctx, xneed * 1.05);
        if (rc) return rc;
    }
    const bool debris = vis_model == FB_MODEL_DEBRIS;
    double *d_q = nullptr, *d_kz = nullptr, *d_I = nullptr, *d_V = nullptr;
    FB_CUDA(cudaMalloc(&d_q, sizeof(double) * n));
    FB_CUDA(cudaMalloc(&d_V, sizeof(double) * n));
    FB_CUDA(cudaMalloc(&d_I, sizeof(double) * N));
    FB_CUDA(cudaMemcpyAsync(d_q, host_q, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    FB_CUDA(cudaMemcpyAsync(d_I, host_I, sizeof(double) * N, cudaMemcpyHostToDevice, ctx->stream));
    if (debris) {
        FB_CUDA(cudaMalloc(&d_kz, sizeof(double) * n));
        FB_CUDA(cudaMemcpyAsync(d_kz, host_kz, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
        std::vector<double> h2(ctx->NC, 0.0);
        for (int k = 0; k < N; k++) h2[k] = host_H2[k];
        FB_CUDA(cudaMemcpyAsync(ctx->d_H2, h2.data(), sizeof(double) * ctx->NC, cudaMemcpyHostToDevice, ctx->stream));
        FB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    k_predict<<<(unsigned)((n + 7) / 8), 256, 0, ctx->stream>>>(n, N, d_q, d_kz, ctx->invQmax, ctx->d_jk, ctx->d_ck, d_I,
                                                              debris ? ctx->d_H2 : nullptr, model_scale, ctx->d_tab, ctx->tab_rows, d_V);
    FB_CUDA(cudaGetLastError());
    FB_CUDA(cudaMemcpyAsync(host_V, d_V, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    FB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_q); cudaFree(d_V); cudaFree(d_I);
    if (d_kz) cudaFree(d_kz);
    return 0;
}
