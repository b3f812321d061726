// K2: uv binning (frank/utilities.py:180-400, UVDataBinner).
//
//   index      idx = int32(floor(uv * (1/width))) with the reference's three fix-ups against
//              bins[i] = float64(i) * width                                   (utilities.py:205-213, 333-347)
//   sums       per bin  sum w,  sum w uv,  sum w Re V,  sum w Im V,  count     (utilities.py:349-361)
//   errors     per bin  sum w^2 (Re V - mu_Re)^2,  sum w^2 (Im V - mu_Im)^2    (utilities.py:236-247)
//
// The reference accumulates with np.bincount in blocks of 65536; here the visibilities are stably sorted by bin
// (fb_sort.cu) so that each bin is a contiguous segment, and every segment is reduced by one warp in a fixed
// order -> deterministic, no floating-point atomics.  Index arithmetic uses single correctly rounded
// multiplications (no FMA contraction): bin indices and counts are bit-exact with NumPy.
#include "fb_common.cuh"

namespace {

__global__ void __launch_bounds__(256)
k_uv_max(int64_t n, const double *__restrict__ uv, double *__restrict__ blockmax)
{
    __shared__ double sh[8];
    double m = -INFINITY;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = fmax(m, uv[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) m = fmax(m, sh[w]);
        blockmax[blockIdx.x] = m;
    }
}

// Pass 1 (coalesced; 32 B read, 44 B written per visibility): bin index, the sort item (key << 32 | position) and
// the packed record (uv, Re V, Im V, w) that the segmented reduction later fetches with one aligned 32-byte load
// instead of three partial-sector gathers.  Out-of-range indices get the key `nbins`: they sort behind every bin.
__global__ void __launch_bounds__(256)
k_bin_index(int64_t n, const double *__restrict__ uv, const double2 *__restrict__ Vc, const double *__restrict__ Vr,
            const double *__restrict__ w, int w_stride, double width, double norm, int nbins, int32_t *__restrict__ idx_out,
            uint64_t *__restrict__ items, double4 *__restrict__ rec)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double x = uv[i];
        int idx = (int)floor(__dmul_rn(x, norm));                                  // utilities.py:338
        if (x < __dmul_rn((double)idx, width)) idx -= 1;                           // :341  fix rounding
        if (idx == nbins) idx -= 1;                                                // :343  point on the outer boundary
        if (x >= __dmul_rn((double)(idx + 1), width) && idx + 1 != nbins) idx += 1;   // :346-347
        idx_out[i] = idx;
        const uint32_t key = (idx >= 0 && idx < nbins) ? (uint32_t)idx : (uint32_t)nbins;
        items[i] = ((uint64_t)key << 32) | (uint64_t)(uint32_t)i;
        double re, im = 0.0;
        if (Vc) { const double2 z = Vc[i]; re = z.x; im = z.y; } else re = Vr[i];
        rec[i] = make_double4(x, re, im, w[(size_t)i * w_stride]);
    }
}

// Segment starts of the sorted items: start[b] = first position whose key is >= b, for b = 0 .. nbins
// (start[nbins] = number of in-range visibilities).  Each key change fills the bins it skips, so empty bins need no
// second pass; no atomics -> the counts do not depend on the input order or on contention.
__global__ void __launch_bounds__(256)
k_bin_starts(int64_t n, const uint64_t *__restrict__ sorted, int nbins, uint32_t *__restrict__ start)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
        const int key = (int)(sorted[p] >> 32);
        const int prev = p > 0 ? (int)(sorted[p - 1] >> 32) : -1;
        for (int b = prev + 1; b <= key; b++) start[b] = (uint32_t)p;
        if (p == n - 1)
            for (int b = key + 1; b <= nbins; b++) start[b] = (uint32_t)n;
    }
}

// Segmented reduction: WPB warps per bin walk the bin's segment of the sorted items ONCE, in a fixed lane-strided
// order, gathering one 32-byte record per visibility, and combine through a fixed shuffle / shared-memory tree.
//   sums (utilities.py:349-361)      sum w uv, sum w, sum w Re V, sum w Im V
//   variance sums (:236-247)         sum w^2 (V - mu)^2 per component, mu = the weighted bin mean
// The reference needs a second pass for the variance because mu is only known after the first; here the deviations
// are accumulated about a pivot c = the bin's first visibility (in sorted order), and
//   sum w^2 (V - mu)^2 = sum w^2 (V - c)^2 - 2 (mu - c) sum w^2 (V - c) + (mu - c)^2 sum w^2
// is exact algebra; because c is a sample of the bin, |mu - c| is of the order of the bin's scatter and the
// subtraction costs a few ulps, not digits (the textbook one-pass formula is this with c = 0, which cancels badly
// for high-S/N bins).  Halves the DRAM traffic of the pass.
template <int WPB>
__global__ void __launch_bounds__(256)
k_bin_reduce(int nbins, const uint32_t *__restrict__ start, const uint64_t *__restrict__ items, const double4 *__restrict__ rec,
             long long *__restrict__ counts, double *__restrict__ sums, double *__restrict__ err)
{
    constexpr int BPB = 8 / WPB;                                    // bins per block
    constexpr int NACC = 9;
    __shared__ double part[8][NACC];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = warp % WPB, lb = warp / WPB;
    const int bin = blockIdx.x * BPB + lb;
    const bool live = bin < nbins;
    const uint32_t s0p = live ? start[bin] : 0u, cnt = live ? start[bin + 1] - s0p : 0u;
    double c_re = 0.0, c_im = 0.0;
    if (cnt > 0) {
        const double4 r0 = rec[(uint32_t)(items[s0p] & 0xffffffffull)];
        c_re = r0.y; c_im = r0.z;
    }
    double a[NACC];
#pragma unroll
    for (int x = 0; x < NACC; x++) a[x] = 0.0;
    for (uint32_t k = sub * 32 + lane; k < cnt; k += 32 * WPB) {
        const double4 r = rec[(uint32_t)(items[s0p + k] & 0xffffffffull)];
        const double wi = r.w, w2 = wi * wi, dr = r.y - c_re, di = r.z - c_im;
        a[0] += wi * r.x;
        a[1] += wi;
        a[2] += wi * r.y;
        a[3] += wi * r.z;
        a[4] += w2 * (dr * dr);
        a[5] += w2 * (di * di);
        a[6] += w2 * dr;
        a[7] += w2 * di;
        a[8] += w2;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int x = 0; x < NACC; x++) a[x] += __shfl_down_sync(0xffffffffu, a[x], o);
    if (WPB > 1) {
        if (lane == 0)
#pragma unroll
            for (int x = 0; x < NACC; x++) part[warp][x] = a[x];
        __syncthreads();
        if (sub == 0 && lane == 0) {
#pragma unroll
            for (int x = 0; x < NACC; x++) {
                double t = 0.0;
                for (int y = 0; y < WPB; y++) t += part[warp + y][x];
                a[x] = t;
            }
        }
    }
    if (sub == 0 && lane == 0 && live) {
        sums[4 * (size_t)bin] = a[0]; sums[4 * (size_t)bin + 1] = a[1]; sums[4 * (size_t)bin + 2] = a[2]; sums[4 * (size_t)bin + 3] = a[3];
        counts[bin] = (long long)cnt;
        double e0 = 0.0, e1 = 0.0;
        if (cnt > 0) {
            const double d_re = a[2] / a[1] - c_re, d_im = a[3] / a[1] - c_im;      // mu - c   (mu: utilities.py:223-224)
            e0 = fmax(0.0, a[4] - 2.0 * d_re * a[6] + d_re * d_re * a[8]);
            e1 = fmax(0.0, a[5] - 2.0 * d_im * a[7] + d_im * d_im * a[8]);
        }
        err[2 * (size_t)bin] = e0; err[2 * (size_t)bin + 1] = e1;
    }
}

}  // namespace

// Workspace of the binner beyond the sort buffers: segment starts [nbins + 1].
static int bin_reserve(fb_ctx *ctx, int64_t n, int nbins)
{
    int rc = fb_reserve_lane(ctx, ctx->lane[0], n, 1);          // sort items (2 x 8 B per visibility), records, digit histograms
    if (rc) return rc;
    if ((size_t)nbins + 1 > ctx->bin_cap) {
        if (ctx->d_binstart) FB_CUDA(cudaFree(ctx->d_binstart));
        ctx->d_binstart = nullptr;
        const size_t cap = (size_t)nbins + nbins / 4 + 1024;
        FB_CUDA(cudaMalloc(&ctx->d_binstart, sizeof(uint32_t) * cap));
        ctx->bin_cap = cap;
    }
    return 0;
}

extern "C" {

int fb_uv_bin_dev(fb_ctx *ctx, int64_t n, const double *dev_uv, const double *dev_V, int v_is_complex, const double *dev_w,
                  int w_stride, double bin_width, int nbins, int32_t *dev_idx, long long *dev_counts, double *dev_sums,
                  double *dev_err)
{
    if (!ctx) return -1;
    if (!dev_uv || !dev_V || !dev_w || !dev_idx || !dev_counts || !dev_sums || !dev_err || n < 1 || nbins < 1 || !(bin_width > 0))
        FB_FAIL(-61, "fb_uv_bin: bad arguments");
    if (n > 0xffffffffLL) FB_FAIL(-60, "fb_uv_bin: more than 2^32 visibilities per call");
    FB_CUDA(cudaSetDevice(ctx->device));
    int rc = bin_reserve(ctx, n, nbins);
    if (rc) return rc;
    const double norm = 1.0 / bin_width;                                        // utilities.py:213
    FbLane &ln = ctx->lane[0];
    uint64_t *buf0 = ln.d_items, *buf1 = ln.d_items + ln.cap;
    const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->num_sms * 32);
    const double2 *Vc = v_is_complex ? (const double2 *)dev_V : nullptr;
    const double *Vr = v_is_complex ? nullptr : dev_V;
    double4 *rec = (double4 *)ln.d_rec;
    k_bin_index<<<grid, 256, 0, ctx->stream>>>(n, dev_uv, Vc, Vr, dev_w, w_stride, bin_width, norm, nbins, dev_idx, buf0, rec);
    FB_CUDA(cudaGetLastError());
    int nbits = 8;
    while (nbits < 32 && (1LL << nbits) <= nbins) nbits += 8;                   // keys 0 .. nbins
    int st = 0;
    const uint64_t *sorted = fb_radix_sort_items(ctx, ln, n, buf0, buf1, nbits, &st);
    if (st) return st;
    k_bin_starts<<<grid, 256, 0, ctx->stream>>>(n, sorted, nbins, ctx->d_binstart);
    // warps per bin: one while there are enough bins to fill the machine, up to a whole block for coarse binnings
    const int want = ctx->num_sms * 32;
    if (nbins >= want)
        k_bin_reduce<1><<<(nbins + 7) / 8, 256, 0, ctx->stream>>>(nbins, ctx->d_binstart, sorted, rec, dev_counts, dev_sums, dev_err);
    else if (nbins * 2 >= want)
        k_bin_reduce<2><<<(nbins + 3) / 4, 256, 0, ctx->stream>>>(nbins, ctx->d_binstart, sorted, rec, dev_counts, dev_sums, dev_err);
    else if (nbins * 4 >= want)
        k_bin_reduce<4><<<(nbins + 1) / 2, 256, 0, ctx->stream>>>(nbins, ctx->d_binstart, sorted, rec, dev_counts, dev_sums, dev_err);
    else
        k_bin_reduce<8><<<nbins, 256, 0, ctx->stream>>>(nbins, ctx->d_binstart, sorted, rec, dev_counts, dev_sums, dev_err);
    FB_CUDA(cudaGetLastError());
    FB_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int fb_uv_max(fb_ctx *ctx, int64_t n, const double *host_uv, double *host_max)
{
    if (!ctx || !host_uv || !host_max || n < 1) return -1;
    FB_CUDA(cudaSetDevice(ctx->device));
    double *d_uv = nullptr, *d_bm = nullptr;
    FB_CUDA(cudaMalloc(&d_uv, sizeof(double) * n));
    FB_CUDA(cudaMalloc(&d_bm, sizeof(double) * 1024));
    FB_CUDA(cudaMemcpyAsync(d_uv, host_uv, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    const int nb = (int)std::min<int64_t>(1024, (n + 255) / 256);
    k_uv_max<<<nb, 256, 0, ctx->stream>>>(n, d_uv, d_bm);
    std::vector<double> bm(nb);
    FB_CUDA(cudaMemcpyAsync(bm.data(), d_bm, sizeof(double) * nb, cudaMemcpyDeviceToHost, ctx->stream));
    FB_CUDA(cudaStreamSynchronize(ctx->stream));
    double m = bm[0];
    for (int i = 1; i < nb; i++) m = bm[i] > m ? bm[i] : m;
    *host_max = m;
    cudaFree(d_uv);
    cudaFree(d_bm);
    return 0;
}

int fb_uv_bin(fb_ctx *ctx, int64_t n, const double *host_uv, const double *host_V, int v_is_complex, const double *host_w,
              int w_stride, double bin_width, int nbins, int32_t *host_idx, long long *host_counts, double *host_sums,
              double *host_err)
{
    if (!ctx) return -1;
    if (!host_uv || !host_V || !host_w || n < 1 || nbins < 1 || !(bin_width > 0)) FB_FAIL(-61, "fb_uv_bin: bad arguments");
    FB_CUDA(cudaSetDevice(ctx->device));
    const int64_t nw = w_stride ? n : 1, nv = v_is_complex ? 2 * n : n;
    // staging: inputs | idx | counts | sums | err in one allocation
    // (V first: its complex loads are 16 bytes wide and need that alignment)
    const size_t in_doubles = (size_t)n + nv + nw, idx_doubles = ((size_t)n + 1) / 2, out_doubles = (size_t)nbins * 7;
    double *d_buf = nullptr;
    FB_CUDA(cudaMalloc(&d_buf, sizeof(double) * (in_doubles + idx_doubles + out_doubles)));
    double *d_V = d_buf, *d_uv = d_V + nv, *d_w = d_uv + n;
    int32_t *d_idx = (int32_t *)(d_w + nw);
    long long *d_c64 = (long long *)(d_buf + in_doubles + idx_doubles);
    double *d_sums = (double *)(d_c64 + nbins), *d_err = d_sums + 4 * (size_t)nbins;
    int rc = 0;
    do {
        if (cudaMemcpyAsync(d_uv, host_uv, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
            cudaMemcpyAsync(d_V, host_V, sizeof(double) * nv, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
            cudaMemcpyAsync(d_w, host_w, sizeof(double) * nw, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
            ctx->err = "fb_uv_bin: host-to-device copy failed"; rc = -62; break;
        }
        rc = fb_uv_bin_dev(ctx, n, d_uv, d_V, v_is_complex, d_w, w_stride, bin_width, nbins, d_idx, d_c64, d_sums, d_err);
        if (rc) break;
        cudaError_t e = cudaSuccess;
        if (host_idx) e = cudaMemcpyAsync(host_idx, d_idx, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess && host_counts) e = cudaMemcpyAsync(host_counts, d_c64, sizeof(long long) * nbins, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess && host_sums) e = cudaMemcpyAsync(host_sums, d_sums, sizeof(double) * 4 * nbins, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess && host_err) e = cudaMemcpyAsync(host_err, d_err, sizeof(double) * 2 * nbins, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { ctx->err = "fb_uv_bin: device-to-host copy failed"; rc = -63; }
    } while (0);
    cudaFree(d_buf);
    return rc;
}

}  // extern "C"
