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

__device__ __forceinline__ double hypot_glibc(double x, double y)
{
    double ax = fabs(x), ay = fabs(y);
    if (ax < ay) { double t = ax; ax = ay; ay = t; }
    // scaling branches of glibc (huge / tiny operands) are irrelevant for baselines in wavelengths
    // (1 .. 1e9) but kept for exactness of the common-case predicate
    if (ax >= __ddiv_rn(ay, 0x1p-54)) return __dadd_rn(ax, ay);
    double h = __dsqrt_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)));
    double t1, t2;
    if (h <= __dmul_rn(2.0, ay)) {
        double delta = __dsub_rn(h, ay);
        t1 = __dmul_rn(ax, __dsub_rn(__dmul_rn(2.0, delta), ax));
        t2 = __dmul_rn(__dsub_rn(delta, __dmul_rn(2.0, __dsub_rn(ax, ay))), delta);
    } else {
        double delta = __dsub_rn(h, ax);
        t1 = __dmul_rn(__dmul_rn(2.0, delta), __dsub_rn(ax, __dmul_rn(2.0, ay)));
        t2 = __dadd_rn(__dmul_rn(__dsub_rn(__dmul_rn(4.0, delta), ay), ay), __dmul_rn(delta, delta));
    }
    h = __dsub_rn(h, __ddiv_rn(__dadd_rn(t1, t2), __dmul_rn(2.0, h)));
    return h;
}

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

// single block: fixed-order reduction of the per-block partials; out = {0.5*sum, qmin, qmax}
__global__ void __launch_bounds__(1024) k_prep_reduce(int nblocks, const double *__restrict__ red, double *__restrict__ out)
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
    if (threadIdx.x == 0) { out[0] = 0.5 * r0; out[1] = r1; out[2] = r2; }
}

}  // namespace

// Grow the per-visibility workspaces (sorted SoA arrays, records, sort items, permutation, tile ranges) to hold
// n_pad visibilities.  Callers that run a mapping call in parts reserve the largest part up front, so that no
// buffer is reallocated while an earlier part's kernels are in flight.
int fb_reserve_prep(fb_ctx *ctx, int64_t n_pad)
{
    if (n_pad > ctx->cap) {
        int64_t cap = n_pad + n_pad / 8 + 4096;
        for (double **p : {&ctx->d_a, &ctx->d_sw, &ctx->d_swV, &ctx->d_kz}) {
            if (*p) FB_CUDA(cudaFree(*p));
            *p = nullptr;
            FB_CUDA(cudaMalloc(p, sizeof(double) * cap));
        }
        if (ctx->d_rec) FB_CUDA(cudaFree(ctx->d_rec));
        if (ctx->d_items) FB_CUDA(cudaFree(ctx->d_items));
        if (ctx->d_perm) FB_CUDA(cudaFree(ctx->d_perm));
        if (ctx->d_amid) FB_CUDA(cudaFree(ctx->d_amid));
        ctx->d_rec = nullptr; ctx->d_items = nullptr; ctx->d_perm = nullptr; ctx->d_amid = nullptr;
        FB_CUDA(cudaMalloc(&ctx->d_amid, sizeof(double) * 2 * (cap / FB_TV + 1)));
        FB_CUDA(cudaMalloc(&ctx->d_rec, sizeof(double) * 4 * cap));
        FB_CUDA(cudaMalloc(&ctx->d_items, sizeof(uint64_t) * 2 * cap));
        FB_CUDA(cudaMalloc(&ctx->d_perm, sizeof(uint32_t) * cap));
        ctx->cap = cap;
    }
    const int per_block = PREP_THREADS * PREP_ITEMS;
    const int nblocks = (int)std::max<int64_t>(1, (n_pad + per_block - 1) / per_block);
    if (nblocks > ctx->red_cap) {
        if (ctx->d_red) FB_CUDA(cudaFree(ctx->d_red));
        ctx->d_red = nullptr;
        const int cap = nblocks + nblocks / 4 + 16;
        FB_CUDA(cudaMalloc(&ctx->d_red, sizeof(double) * (3 * (size_t)cap + 8)));
        ctx->red_cap = cap;
    }
    return fb_reserve_sort(ctx, n_pad);
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

int fb_launch_prep(fb_ctx *ctx, int64_t n, const double *u, const double *v, const double *V, const double *w,
                   int w_stride, const fb_geometry *g, double *dev_H0, double *host_qminmax, double *host_H0)
{
    const int64_t n_pad = ((n + FB_TV - 1) / FB_TV) * FB_TV;
    {
        int rc = fb_reserve_prep(ctx, n_pad);
        if (rc) return rc;
    }
    const int per_block = PREP_THREADS * PREP_ITEMS;
    int nblocks = (int)((n_pad + per_block - 1) / per_block);
    if (nblocks < 1) nblocks = 1;
    k_prep<<<nblocks, PREP_THREADS, 0, ctx->stream>>>(n, u, v, (const double2 *)V, w, w_stride, *g, ctx->invQmax,
                                                      (double4 *)ctx->d_rec, ctx->d_red);
    FB_CUDA(cudaGetLastError());
    double *fin = ctx->d_red + 3 * (size_t)ctx->red_cap;   // 3 doubles after the block partials
    k_prep_reduce<<<1, 1024, 0, ctx->stream>>>(nblocks, ctx->d_red, fin);
    FB_CUDA(cudaGetLastError());
    double h[3];
    FB_CUDA(cudaMemcpyAsync(h, fin, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    if (dev_H0) FB_CUDA(cudaMemcpyAsync(dev_H0, fin, sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    FB_CUDA(cudaStreamSynchronize(ctx->stream));
    host_qminmax[0] = h[1];
    host_qminmax[1] = h[2];
    if (host_H0) *host_H0 = h[0];
    ctx->last_n = n;
    // order the visibilities by baseline bin (stable) and lay them out for the Gram kernel
    return fb_launch_sort(ctx, n, n_pad, n > 0 ? h[2] * ctx->invQmax : 0.0);
}
