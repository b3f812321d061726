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

template <bool DEBRIS>
__global__ void __launch_bounds__(FB_GRAM_THREADS, 1) k_gram(GramArgs p)
{
    extern __shared__ double smem[];
    double *G = smem;                                  // [2*PCOLS][LDV]
    double *col_jk = G + 2 * FB_PCOLS * FB_LDV;        // [2*PCOLS]  j_k, or -1 (data column), -2 (padding)
    double *col_h2 = col_jk + 2 * FB_PCOLS;            // [2*PCOLS]  debris H2_k

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int smsp = warp & 3, slot = warp >> 2;
    const int last_row = p.tab_rows - 1;
    const int frag = (lane >> 2) * FB_LDV + (lane & 3);   // fragment offset inside an 8-mode x 4-vis brick

    for (int item = blockIdx.x; item < p.ntypes * p.C; item += gridDim.x) {
        const int type = item % p.ntypes, chunk = item / p.ntypes;
        const FbGramType ty = p.types[type];
        const long long t0 = (p.n_tiles * chunk) / p.C, t1 = (p.n_tiles * (chunk + 1)) / p.C;
        const int ncolA = ty.a_nt * 8, ncol = (ty.a_nt + ty.b_nt) * 8;

        __syncthreads();
        for (int lc = tid; lc < ncol; lc += FB_GRAM_THREADS) {
            int g = lc < ncolA ? ty.a_t0 * 8 + lc : ty.b_t0 * 8 + (lc - ncolA);
            col_jk[lc] = g < p.N ? p.jk[g] : (g == p.N ? -1.0 : -2.0);
            if (DEBRIS) col_h2[lc] = g < p.N ? p.H2[g] : 0.0;
        }

        double acc[5][5][2];
#pragma unroll
        for (int r = 0; r < 5; r++)
#pragma unroll
            for (int c = 0; c < 5; c++) acc[r][c][0] = acc[r][c][1] = 0.0;

        // warp's register block
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
        __syncthreads();

        for (long long tile = t0; tile < t1; ++tile) {
            // ---------------- phase 1: design-matrix tile into shared memory ----------------------
            const long long v0 = tile * FB_TV;
            const double a0 = p.a[v0 + lane], a1 = p.a[v0 + lane + 32];
            const double s0 = p.sw[v0 + lane], s1 = p.sw[v0 + lane + 32];
            double k0 = 0.0, k1 = 0.0;
            if (DEBRIS) {
                k0 = p.kz[v0 + lane]; k1 = p.kz[v0 + lane + 32];
                k0 = -k0 * k0; k1 = -k1 * k1;
            }
            for (int lc = warp; lc < ncol; lc += FB_GRAM_THREADS / 32) {
                const double jk = col_jk[lc];
                const int sc = lc < ncolA ? lc : FB_PCOLS + (lc - ncolA);
                double g0, g1;
                if (jk >= 0.0) {
                    g0 = j0_tab(__dmul_rn(a0, jk), p.tab, last_row);
                    g1 = j0_tab(__dmul_rn(a1, jk), p.tab, last_row);
                    if (DEBRIS) {
                        const double h2 = col_h2[lc];
                        g0 *= exp(k0 * h2);
                        g1 *= exp(k1 * h2);
                    }
                    g0 *= s0;
                    g1 *= s1;
                } else if (jk == -1.0) {
                    g0 = p.swV[v0 + lane];
                    g1 = p.swV[v0 + lane + 32];
                } else {
                    g0 = 0.0;
                    g1 = 0.0;
                }
                G[sc * FB_LDV + lane] = g0;
                G[sc * FB_LDV + lane + 32] = g1;
            }
            __syncthreads();

            // ---------------- phase 2: DMMA over the tile -----------------------------------------
            if (ty.kind == FB_KIND_OFF) {
                const double *ap = G + (r0 * 8) * FB_LDV + frag;
                const double *bp = G + (FB_PCOLS + c0 * 8) * FB_LDV + frag;
                if (nr == 5 && nc == 5) {
#pragma unroll 2
                    for (int ks = 0; ks < FB_TV / 4; ks++) {
                        double af[5];
#pragma unroll
                        for (int r = 0; r < 5; r++) af[r] = ap[r * 8 * FB_LDV + ks * 4];
#pragma unroll
                        for (int c = 0; c < 5; c++) {
                            const double b = bp[c * 8 * FB_LDV + ks * 4];
#pragma unroll
                            for (int r = 0; r < 5; r++) dmma(acc[r][c], af[r], b);
                        }
                    }
                } else if (nr > 0 && nc > 0) {
                    for (int ks = 0; ks < FB_TV / 4; ks++) {
                        double af[5];
#pragma unroll
                        for (int r = 0; r < 5; r++) af[r] = r < nr ? ap[r * 8 * FB_LDV + ks * 4] : 0.0;
#pragma unroll
                        for (int c = 0; c < 5; c++) {
                            if (c < nc) {
                                const double b = bp[c * 8 * FB_LDV + ks * 4];
#pragma unroll
                                for (int r = 0; r < 5; r++)
                                    if (r < nr) dmma(acc[r][c], af[r], b);
                            }
                        }
                    }
                }
            } else if (nr > 0 && nc > 0) {
                // skewed strip of a triangle: acc[r][d] += G_{r0+r}^T G_{(r0+r + c0+d) mod n}
                const double *ap = G + (base_cols + r0 * 8) * FB_LDV + frag;
                const double *tp = G + base_cols * FB_LDV + frag;
                int boff[9];
#pragma unroll
                for (int s = 0; s < 9; s++) boff[s] = ((r0 + c0 + s) % nmod) * 8 * FB_LDV;
                const bool full = (nr == 5 && nc == 5);
#pragma unroll 1
                for (int ks = 0; ks < FB_TV / 4; ks++) {
                    double af[5];
#pragma unroll
                    for (int r = 0; r < 5; r++) af[r] = r < nr ? ap[r * 8 * FB_LDV + ks * 4] : 0.0;
#pragma unroll
                    for (int s = 0; s < 9; s++) {
                        if (full || s < nr + nc - 1) {
                            const double b = tp[boff[s] + ks * 4];
#pragma unroll
                            for (int r = 0; r < 5; r++) {
                                const int d = s - r;
                                if (d >= 0 && d < 5) {
                                    if (full || (r < nr && d < nc)) dmma(acc[r][d], af[r], b);
                                }
                            }
                        }
                    }
                }
            }
            __syncthreads();
        }

        // ---------------- write the partial block ------------------------------------------------
        double *out = p.partial + (size_t)item * FB_PSZ;
        if (ty.kind == FB_KIND_OFF) {
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

    const size_t smem = sizeof(double) * (2 * FB_PCOLS * FB_LDV + 4 * FB_PCOLS);
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
