// K1: geometry pre-pass.  One coalesced pass over (u, v, V, w):
//   phase-centre shift     frank/geometry.py:69-75   (apply_phase_shift, inverse=True)
//   deprojection           frank/geometry.py:111-131 (deproject)
//   q = hypot(u', v')      frank/statistical_models.py:166
//   Re(V'), w broadcast    frank/statistical_models.py:172-173
//   H0 terms, min/max q    frank/statistical_models.py:218, 512-535
// and emits what the Gram kernel consumes: a = q * (1/Qmax) (hankel.py:189,202: `k * q`),
// sqrt(w), sqrt(w) * Re(V'), kz.
//
// Bit-level contract: q (and therefore the J0 argument a * j_k) must equal NumPy's float64 result, because
// J0 turns a 1-ulp change of an argument near 900 into a 3e-15 change of the value.  So every operation on
// the (u, v) -> q path is a single correctly rounded IEEE operation in NumPy's order (no FMA contraction),
// and hypot() follows glibc's non-FMA kernel (sysdeps/ieee754/dbl-64/e_hypot.c, glibc >= 2.35), which is what
// np.hypot calls; tests/test_prep_bits.py checks bit equality.
#include "fb_common.cuh"

#include <algorithm>

namespace {

constexpr int PREP_THREADS = 256;
constexpr int PREP_ITEMS = 8;     // visibilities per thread -> 2048 per block

// deterministic block reduction (fixed shuffle tree, fixed warp order)
template <typename Op>
__device__ __forceinline__ double block_reduce(double v, double *scratch, Op op, double ident)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_down_sync(0xffffffffu, v, o));
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < (blockDim.x >> 5) ? scratch[lane] : ident;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_down_sync(0xffffffffu, v, o));
    }
    return v;   // valid in thread 0
}

struct OpAdd { __device__ double operator()(double a, double b) const { return a + b; } };
struct OpMin { __device__ double operator()(double a, double b) const { return fmin(a, b); } };
struct OpMax { __device__ double operator()(double a, double b) const { return fmax(a, b); } };

__global__ void __launch_bounds__(PREP_THREADS)
k_prep(int64_t n, const double *__restrict__ u, const double *__restrict__ v,
       const double2 *__restrict__ V, const double *__restrict__ w, int w_stride, fb_geometry g, double invQmax,
       double4 *__restrict__ out_rec, double *__restrict__ red)
{
    __shared__ double scratch[PREP_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * (PREP_THREADS * PREP_ITEMS);
    double h0 = 0.0, qmin = INFINITY, qmax = -INFINITY;
    const double two_pi = 6.283185307179586;   // float64(2*np.pi)
#pragma unroll
    for (int it = 0; it < PREP_ITEMS; it++) {
        int64_t i = base + (int64_t)it * PREP_THREADS + threadIdx.x;
        if (i < n) {
            double ui = u[i], vi = v[i];
            double2 Vi = V[i];
            double wi = w[i * w_stride];
            // phi = u*dRA + v*dDec                                       geometry.py:72
            double phi = __dadd_rn(__dmul_rn(ui, g.a_ra), __dmul_rn(vi, g.a_dec));
            double s, c;
            sincos(phi, &s, &c);
            // Re[ V / (cos phi + i sin phi) ] with NumPy's complex128 division (Smith's algorithm)
            double vre;
            if (fabs(c) >= fabs(s)) {
                double rat = __ddiv_rn(s, c);
                double scl = __ddiv_rn(1.0, __dadd_rn(c, __dmul_rn(s, rat)));
                vre = __dmul_rn(__dadd_rn(Vi.x, __dmul_rn(Vi.y, rat)), scl);
            } else {
                double rat = __ddiv_rn(c, s);
                double scl = __ddiv_rn(1.0, __dadd_rn(s, __dmul_rn(c, rat)));
                vre = __dmul_rn(__dadd_rn(__dmul_rn(Vi.x, rat), Vi.y), scl);
            }
            // deprojection                                                  geometry.py:122-131
            double up = __dsub_rn(__dmul_rn(ui, g.cos_pa), __dmul_rn(vi, g.sin_pa));
            double vp = __dadd_rn(__dmul_rn(ui, g.sin_pa), __dmul_rn(vi, g.cos_pa));
            double kz = __dmul_rn(up, g.sin_inc);
            up = __dmul_rn(up, g.cos_inc);
            double q = hypot_glibc(up, vp);
            double sw = __dsqrt_rn(wi);
            out_rec[i] = make_double4(__dmul_rn(q, invQmax), sw, __dmul_rn(sw, vre), kz);
            // H0 term                                                       statistical_models.py:218
            h0 += log(__ddiv_rn(wi, two_pi)) - __dmul_rn(__dmul_rn(vre, wi), vre);
            qmin = fmin(qmin, q);
            qmax = fmax(qmax, q);
        }
    }
    double r0 = block_reduce(h0, scratch, OpAdd(), 0.0);
    double r1 = block_reduce(qmin, scratch, OpMin(), INFINITY);
    double r2 = block_reduce(qmax, scratch, OpMax(), -INFINITY);
    if (threadIdx.x == 0) {
        red[3 * (int64_t)blockIdx.x + 0] = r0;
        red[3 * (int64_t)blockIdx.x + 1] = r1;
        red[3 * (int64_t)blockIdx.x + 2] = r2;
    }
}

// SourceGeometry.apply_correction as a stand-alone pass (geometry.py:202-236) for callers that want the corrected
// arrays themselves (uv binning, plotting of deprojected visibilities): the same single correctly rounded operations
// as k_prep, full complex quotient V / (cos phi + i sin phi) by NumPy's Smith division, optional q = hypot(u', v')
// with the glibc kernel (bit-equal to np.hypot).  32 B read, up to 48 B written per visibility, coalesced.
__global__ void __launch_bounds__(256)
k_apply_correction(int64_t n, const double *__restrict__ u, const double *__restrict__ v, const double2 *__restrict__ V,
                   fb_geometry g, double *__restrict__ up_out, double *__restrict__ vp_out, double *__restrict__ wp_out,
                   double2 *__restrict__ Vp_out, double *__restrict__ q_out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double ui = u[i], vi = v[i];
        if (V && Vp_out) {
            const double2 Vi = V[i];
            const double phi = __dadd_rn(__dmul_rn(ui, g.a_ra), __dmul_rn(vi, g.a_dec));       // geometry.py:72
            double s, c;
            sincos(phi, &s, &c);
            double re, im;
            if (fabs(c) >= fabs(s)) {
                const double rat = __ddiv_rn(s, c);
                const double scl = __ddiv_rn(1.0, __dadd_rn(c, __dmul_rn(s, rat)));
                re = __dmul_rn(__dadd_rn(Vi.x, __dmul_rn(Vi.y, rat)), scl);
                im = __dmul_rn(__dsub_rn(Vi.y, __dmul_rn(Vi.x, rat)), scl);
            } else {
                const double rat = __ddiv_rn(c, s);
                const double scl = __ddiv_rn(1.0, __dadd_rn(s, __dmul_rn(c, rat)));
                re = __dmul_rn(__dadd_rn(__dmul_rn(Vi.x, rat), Vi.y), scl);
                im = __dmul_rn(__dsub_rn(__dmul_rn(Vi.y, rat), Vi.x), scl);
            }
            Vp_out[i] = make_double2(re, im);
        }
        double up = __dsub_rn(__dmul_rn(ui, g.cos_pa), __dmul_rn(vi, g.sin_pa));               // geometry.py:122-131
        const double vp = __dadd_rn(__dmul_rn(ui, g.sin_pa), __dmul_rn(vi, g.cos_pa));
        const double wp = __dmul_rn(up, g.sin_inc);
        up = __dmul_rn(up, g.cos_inc);
        if (up_out) up_out[i] = up;
        if (vp_out) vp_out[i] = vp;
        if (wp_out) wp_out[i] = wp;
        if (q_out) q_out[i] = hypot_glibc(up, vp);
    }
}

// single block: fixed-order reduction of the per-block partials of one chunk; out = {0.5 * sum, qmin, qmax, 0}.
// Also raises the call's status bits on the device, so that no host read is needed in the middle of a call
// (statistical_models.py:526: data beyond the last collocation point; J0 table shorter than a_max * j_{N-1}).
__global__ void __launch_bounds__(1024)
k_prep_reduce(int nblocks, const double *__restrict__ red, double *__restrict__ out, int check_qbounds, double q_last,
              double x_per_q, double x_table, int *__restrict__ status)
{
    __shared__ double scratch[32];
    double h0 = 0.0, qmin = INFINITY, qmax = -INFINITY;
    // contiguous slab per thread keeps the order fixed and the partial sums of similar magnitude
    int per = (nblocks + blockDim.x - 1) / blockDim.x;
    int b0 = threadIdx.x * per, b1 = min(nblocks, b0 + per);
    for (int b = b0; b < b1; b++) {
        h0 += red[3 * b];
        qmin = fmin(qmin, red[3 * b + 1]);
        qmax = fmax(qmax, red[3 * b + 2]);
    }
    double r0 = block_reduce(h0, scratch, OpAdd(), 0.0);
    double r1 = block_reduce(qmin, scratch, OpMin(), INFINITY);
    double r2 = block_reduce(qmax, scratch, OpMax(), -INFINITY);
    if (threadIdx.x == 0) {
        out[0] = 0.5 * r0; out[1] = r1; out[2] = r2; out[3] = 0.0;
        int st = 0;
        if (check_qbounds && q_last < r2) st |= FB_ST_QRANGE;
        else if (r2 * x_per_q > x_table) st |= FB_ST_TABLE;
        if (st) atomicOr(status, st);
    }
}

// End of a call: combine the chunks' reductions in chunk order -> result = {H0, min q, max q, status}.
__global__ void k_map_result(int nchunks, const double *__restrict__ chunkred, const int *__restrict__ status,
                             double *__restrict__ result, double *__restrict__ dev_H0)
{
    double h0 = 0.0, qmin = INFINITY, qmax = -INFINITY;
    for (int c = 0; c < nchunks; c++) {
        h0 += chunkred[4 * c];
        qmin = fmin(qmin, chunkred[4 * c + 1]);
        qmax = fmax(qmax, chunkred[4 * c + 2]);
    }
    result[0] = h0; result[1] = qmin; result[2] = qmax; result[3] = (double)*status;
    if (dev_H0 && !(*status)) dev_H0[0] = h0;
}

}  // namespace

// Grow the per-visibility workspaces of a lane (sorted SoA arrays, records, sort items, permutation, tile ranges) to
// hold n visibilities of nchan channels (every channel's segment is padded to a whole tile).  Never called while the
// lane has work in flight that uses the old buffers: the entry points reserve for the largest chunk up front.
int fb_reserve_lane(fb_ctx *ctx, FbLane &ln, int64_t n, int nchan)
{
    const int64_t n_pad = (n + FB_TV - 1) / FB_TV * FB_TV + (int64_t)FB_TV * nchan;
    if (n_pad > ln.cap) {
        FB_CUDA(cudaStreamSynchronize(ln.stream));
        const int64_t cap = n_pad + n_pad / 8 + 4096;
        for (double **p : {&ln.d_a, &ln.d_sw, &ln.d_swV, &ln.d_kz}) {
            if (*p) FB_CUDA(cudaFree(*p));
            *p = nullptr;
            FB_CUDA(cudaMalloc(p, sizeof(double) * cap));
        }
        if (ln.d_rec) FB_CUDA(cudaFree(ln.d_rec));
        if (ln.d_items) FB_CUDA(cudaFree(ln.d_items));
        if (ln.d_perm) FB_CUDA(cudaFree(ln.d_perm));
        if (ln.d_amid) FB_CUDA(cudaFree(ln.d_amid));
        ln.d_rec = nullptr; ln.d_items = nullptr; ln.d_perm = nullptr; ln.d_amid = nullptr;
        FB_CUDA(cudaMalloc(&ln.d_amid, sizeof(double) * 2 * (cap / FB_TV + 1)));
        FB_CUDA(cudaMalloc(&ln.d_rec, sizeof(double) * 4 * cap));
        FB_CUDA(cudaMalloc(&ln.d_items, sizeof(uint64_t) * 2 * cap));
        FB_CUDA(cudaMalloc(&ln.d_perm, sizeof(uint32_t) * cap));
        ln.cap = cap;
    }
    const int per_block = PREP_THREADS * PREP_ITEMS;
    const int nblocks = (int)std::max<int64_t>(1, (n + per_block - 1) / per_block);
    if (nblocks > ln.red_cap) {
        FB_CUDA(cudaStreamSynchronize(ln.stream));
        if (ln.d_red) FB_CUDA(cudaFree(ln.d_red));
        ln.d_red = nullptr;
        const int cap = nblocks + nblocks / 4 + 16;
        FB_CUDA(cudaMalloc(&ln.d_red, sizeof(double) * (3 * (size_t)cap + 8)));
        ln.red_cap = cap;
    }
    if (!ln.d_seg) FB_CUDA(cudaMalloc(&ln.d_seg, sizeof(int) * 2 * (FB_MAX_CHAN + 1)));
    return fb_reserve_sort(ctx, ln, n);
}

extern "C" int fb_apply_correction_dev(fb_ctx *ctx, int64_t n, const double *dev_u, const double *dev_v, const double *dev_V_reim,
                                       const fb_geometry *geom, double *dev_up, double *dev_vp, double *dev_wp, double *dev_Vp_reim,
                                       double *dev_q)
{
    if (!ctx) return -1;
    if (n < 0 || !dev_u || !dev_v || !geom) FB_FAIL(-12, "fb_apply_correction_dev: bad arguments");
    if (n == 0) return 0;
    FB_CUDA(cudaSetDevice(ctx->device));
    const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->num_sms * 32);
    k_apply_correction<<<grid, 256, 0, ctx->stream>>>(n, dev_u, dev_v, (const double2 *)dev_V_reim, *geom, dev_up, dev_vp, dev_wp,
                                                      (double2 *)dev_Vp_reim, dev_q);
    FB_CUDA(cudaGetLastError());
    FB_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// Enqueue the geometry pre-pass of one chunk on the lane's stream: records, block partials, the chunk's reduction
// (H0 sum, min / max q) into ctx->d_chunkred[chunk] and the call's status bits.  Nothing is read back here.
int fb_enqueue_prep(fb_ctx *ctx, FbLane &ln, int chunk, int64_t n, const double *u, const double *v, const double *V,
                    const double *w, int w_stride, const int32_t *chan, const FbMapJob &job)
{
    (void)chan;
    const int per_block = PREP_THREADS * PREP_ITEMS;
    int nblocks = (int)((n + per_block - 1) / per_block);
    if (nblocks < 1) nblocks = 1;
    k_prep<<<nblocks, PREP_THREADS, 0, ln.stream>>>(n, u, v, (const double2 *)V, w, w_stride, job.geom, ctx->invQmax,
                                                    (double4 *)ln.d_rec, ln.d_red);
    FB_CUDA(cudaGetLastError());
    // largest J0 argument the table serves: rows reach (tab_rows - 3) / 16 (fb_j0_rows_for)
    const double x_table = (double)(ctx->tab_rows - 3) * FB_J0_H;
    k_prep_reduce<<<1, 1024, 0, ln.stream>>>(nblocks, ln.d_red, ctx->d_chunkred + 4 * chunk, job.check_qbounds, job.q_last,
                                             ctx->invQmax * ctx->h_jk[ctx->N - 1], x_table, ctx->d_status);
    FB_CUDA(cudaGetLastError());
    ctx->last_n = n;
    return 0;
}

int fb_enqueue_result(fb_ctx *ctx, cudaStream_t st, int nchunks, double *dev_H0)
{
    k_map_result<<<1, 1, 0, st>>>(nchunks, ctx->d_chunkred, ctx->d_status, ctx->d_result, dev_H0);
    FB_CUDA(cudaGetLastError());
    return 0;
}


// ---- small deterministic Gram of K device vectors (the normal equations of the geometry fit's Levenberg-Marquardt step) ----
// G[a][b] = sum_i x_a[i] x_b[i], a <= b < K <= 8.  Pass 1: a fixed slab per block, fixed tree inside the block; pass 2: one
// block sums the block partials in a fixed order.  Replaces the 2n x 4 Jacobian that scipy.optimize.least_squares builds and
// factorises on the host at every iteration of FitGeometryFourierBessel (frank/geometry.py:745-746).
namespace {

constexpr int CG_MAXK = 8, CG_PAIRS = CG_MAXK * (CG_MAXK + 1) / 2, CG_BLOCKS = 592;

struct ColPtrs { const double *p[CG_MAXK]; };

__global__ void __launch_bounds__(256)
k_cols_gram(int64_t n, int K, ColPtrs cols, double *__restrict__ partial)
{
    __shared__ double scratch[PREP_THREADS / 32];
    double acc[CG_PAIRS];
#pragma unroll
    for (int e = 0; e < CG_PAIRS; e++) acc[e] = 0.0;
    const int64_t per = (n + gridDim.x - 1) / gridDim.x, i0 = (int64_t)blockIdx.x * per, i1 = i0 + per < n ? i0 + per : n;
    for (int64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        double x[CG_MAXK];
#pragma unroll
        for (int a = 0; a < CG_MAXK; a++) x[a] = a < K ? cols.p[a][i] : 0.0;
        int e = 0;
#pragma unroll
        for (int a = 0; a < CG_MAXK; a++)
#pragma unroll
            for (int b = a; b < CG_MAXK; b++, e++) acc[e] = fma(x[a], x[b], acc[e]);
    }
    int e = 0;
    for (int a = 0; a < CG_MAXK; a++)
        for (int b = a; b < CG_MAXK; b++, e++) {
            if (a >= K || b >= K) continue;
            const double r = block_reduce(acc[e], scratch, OpAdd(), 0.0);
            if (threadIdx.x == 0) partial[(size_t)blockIdx.x * CG_PAIRS + e] = r;
        }
}

__global__ void __launch_bounds__(64)
k_cols_gram_final(int nblocks, int K, const double *__restrict__ partial, double *__restrict__ G)
{
    const int e = threadIdx.x;
    if (e >= CG_PAIRS) return;
    int a = 0, rem = e;
    while (rem >= CG_MAXK - a) { rem -= CG_MAXK - a; a++; }
    const int b = a + rem;
    if (a >= K || b >= K) return;
    double s = 0.0;
    for (int k = 0; k < nblocks; k++) s += partial[(size_t)k * CG_PAIRS + e];
    G[a * K + b] = s;
    G[b * K + a] = s;
}

}  // namespace

extern "C" int fb_columns_gram_dev(fb_ctx *ctx, int64_t n, int K, const double *const *dev_cols, double *host_G)
{
    if (!ctx) return -1;
    if (n < 1 || K < 1 || K > CG_MAXK || !dev_cols || !host_G) FB_FAIL(-12, "fb_columns_gram_dev: bad arguments (1 <= K <= 8)");
    FB_CUDA(cudaSetDevice(ctx->device));
    FbLane &ln = ctx->lane[0];
    const int need = CG_BLOCKS * CG_PAIRS / 3 + 64;                  // d_red holds 3 doubles per entry of red_cap
    if (need > ln.red_cap) {
        FB_CUDA(cudaStreamSynchronize(ln.stream));
        if (ln.d_red) FB_CUDA(cudaFree(ln.d_red));
        ln.d_red = nullptr;
        FB_CUDA(cudaMalloc(&ln.d_red, sizeof(double) * (3 * (size_t)need + 8)));
        ln.red_cap = need;
    }
    ColPtrs cols;
    for (int a = 0; a < CG_MAXK; a++) cols.p[a] = a < K ? dev_cols[a] : nullptr;
    double *d_G = ln.d_red + (size_t)CG_BLOCKS * CG_PAIRS;
    const int nblocks = (int)std::min<int64_t>(CG_BLOCKS, (n + 255) / 256);
    k_cols_gram<<<nblocks, 256, 0, ctx->stream>>>(n, K, cols, ln.d_red);
    k_cols_gram_final<<<1, 64, 0, ctx->stream>>>(nblocks, K, ln.d_red, d_G);
    FB_CUDA(cudaGetLastError());
    FB_CUDA(cudaMemcpyAsync(host_G, d_G, sizeof(double) * K * K, cudaMemcpyDeviceToHost, ctx->stream));
    FB_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
