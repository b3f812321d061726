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

__global__ void __launch_bounds__(256)
k_bin_index(int64_t n, const double *__restrict__ uv, double width, double norm, int nbins, int32_t *__restrict__ idx_out,
            uint32_t *__restrict__ counts)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = uv[i];
    int idx = (int)floor(__dmul_rn(x, norm));                                  // utilities.py:338
    if (x < __dmul_rn((double)idx, width)) idx -= 1;                           // :341  fix rounding
    if (idx == nbins) idx -= 1;                                                // :343  point on the outer boundary
    if (x >= __dmul_rn((double)(idx + 1), width) && idx + 1 != nbins) idx += 1;   // :346-347
    idx_out[i] = idx;
    if (idx >= 0 && idx < nbins) atomicAdd(&counts[idx], 1u);                  // integer: order independent
}

// one warp per bin: fixed lane-strided partial sums, fixed shuffle tree
template <bool ERRORS>
__global__ void __launch_bounds__(256)
k_bin_reduce(int nbins, const uint32_t *__restrict__ offs, const uint32_t *__restrict__ counts, const uint64_t *__restrict__ items,
             const double *__restrict__ uv, const double2 *__restrict__ Vc, const double *__restrict__ Vr,
             const double *__restrict__ w, int w_stride, const double *__restrict__ sums_in, double *__restrict__ out)
{
    const int bin = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (bin >= nbins) return;
    const uint32_t start = offs[bin], cnt = counts[bin];
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    double mu_re = 0.0, mu_im = 0.0;
    if (ERRORS && cnt > 0) {
        const double sw = sums_in[4 * (size_t)bin + 1];
        mu_re = sums_in[4 * (size_t)bin + 2] / sw;                             // utilities.py:223-224
        mu_im = sums_in[4 * (size_t)bin + 3] / sw;
    }
    for (uint32_t k = lane; k < cnt; k += 32) {
        const uint32_t i = (uint32_t)(items[start + k] & 0xffffffffull);
        const double wi = w[(size_t)i * w_stride];
        const double re = Vc ? Vc[i].x : Vr[i], im = Vc ? Vc[i].y : 0.0;
        if (ERRORS) {
            const double w2 = wi * wi, dr = re - mu_re, di = im - mu_im;      // utilities.py:243-247
            s0 += w2 * (dr * dr);
            s1 += w2 * (di * di);
        } else {
            s0 += wi * uv[i];
            s1 += wi;
            s2 += wi * re;
            s3 += wi * im;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_down_sync(0xffffffffu, s0, o);
        s1 += __shfl_down_sync(0xffffffffu, s1, o);
        if (!ERRORS) {
            s2 += __shfl_down_sync(0xffffffffu, s2, o);
            s3 += __shfl_down_sync(0xffffffffu, s3, o);
        }
    }
    if (lane == 0) {
        if (ERRORS) { out[2 * (size_t)bin] = s0; out[2 * (size_t)bin + 1] = s1; }
        else { out[4 * (size_t)bin] = s0; out[4 * (size_t)bin + 1] = s1; out[4 * (size_t)bin + 2] = s2; out[4 * (size_t)bin + 3] = s3; }
    }
}

__global__ void k_u32_to_i64(int n, const uint32_t *__restrict__ in, long long *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

// exclusive scan (single block) of uint32 counters into offs
__global__ void __launch_bounds__(1024) k_excl_scan(int n, const uint32_t *__restrict__ in, uint32_t *__restrict__ out)
{
    __shared__ uint32_t part[1024];
    const int per = (n + 1023) / 1024, b0 = threadIdx.x * per, b1 = min(n, b0 + per);
    uint32_t s = 0;
    for (int i = b0; i < b1; i++) s += in[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int t = 0; t < 1024; t++) { uint32_t c = part[t]; part[t] = run; run += c; }
    }
    __syncthreads();
    uint32_t run = part[threadIdx.x];
    for (int i = b0; i < b1; i++) { out[i] = run; run += in[i]; }
}

}  // namespace

extern "C" {

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
    if (!ctx || !host_uv || !host_V || !host_w || n < 1 || nbins < 1 || !(bin_width > 0)) return -1;
    if (n > 0xffffffffLL) FB_FAIL(-60, "fb_uv_bin: more than 2^32 visibilities per call");
    FB_CUDA(cudaSetDevice(ctx->device));
    const int64_t nw = w_stride ? n : 1, nv = v_is_complex ? 2 * n : n;
    double *d_uv = nullptr, *d_V = nullptr, *d_w = nullptr, *d_sums = nullptr, *d_err = nullptr;
    int32_t *d_idx = nullptr;
    uint32_t *d_cnt = nullptr, *d_off = nullptr;
    uint64_t *d_it = nullptr;
    long long *d_c64 = nullptr;
    FB_CUDA(cudaMalloc(&d_uv, sizeof(double) * n));
    FB_CUDA(cudaMalloc(&d_V, sizeof(double) * nv));
    FB_CUDA(cudaMalloc(&d_w, sizeof(double) * nw));
    FB_CUDA(cudaMalloc(&d_idx, sizeof(int32_t) * n));
    FB_CUDA(cudaMalloc(&d_cnt, sizeof(uint32_t) * nbins));
    FB_CUDA(cudaMalloc(&d_off, sizeof(uint32_t) * nbins));
    FB_CUDA(cudaMalloc(&d_c64, sizeof(long long) * nbins));
    FB_CUDA(cudaMalloc(&d_sums, sizeof(double) * 4 * nbins));
    FB_CUDA(cudaMalloc(&d_err, sizeof(double) * 2 * nbins));
    FB_CUDA(cudaMalloc(&d_it, sizeof(uint64_t) * 2 * n));
    FB_CUDA(cudaMemcpyAsync(d_uv, host_uv, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    FB_CUDA(cudaMemcpyAsync(d_V, host_V, sizeof(double) * nv, cudaMemcpyHostToDevice, ctx->stream));
    FB_CUDA(cudaMemcpyAsync(d_w, host_w, sizeof(double) * nw, cudaMemcpyHostToDevice, ctx->stream));
    FB_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(uint32_t) * nbins, ctx->stream));
    const double norm = 1.0 / bin_width;                                        // utilities.py:213
    k_bin_index<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, d_uv, bin_width, norm, nbins, d_idx, d_cnt);
    FB_CUDA(cudaGetLastError());
    k_excl_scan<<<1, 1024, 0, ctx->stream>>>(nbins, d_cnt, d_off);
    int rc = fb_items_from_keys(ctx, n, d_idx, d_it);
    if (rc) return rc;
    int nbits = 8;
    while (nbits < 32 && (1LL << nbits) < nbins) nbits += 8;
    int st = 0;
    uint64_t *sorted = fb_radix_sort_items(ctx, n, d_it, d_it + n, nbits, &st);
    if (st) return st;
    const double2 *Vc = v_is_complex ? (const double2 *)d_V : nullptr;
    const double *Vr = v_is_complex ? nullptr : d_V;
    k_bin_reduce<false><<<(nbins + 7) / 8, 256, 0, ctx->stream>>>(nbins, d_off, d_cnt, sorted, d_uv, Vc, Vr, d_w, w_stride, nullptr, d_sums);
    k_bin_reduce<true><<<(nbins + 7) / 8, 256, 0, ctx->stream>>>(nbins, d_off, d_cnt, sorted, d_uv, Vc, Vr, d_w, w_stride, d_sums, d_err);
    k_u32_to_i64<<<(nbins + 255) / 256, 256, 0, ctx->stream>>>(nbins, d_cnt, d_c64);
    FB_CUDA(cudaGetLastError());
    if (host_idx) FB_CUDA(cudaMemcpyAsync(host_idx, d_idx, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (host_counts) FB_CUDA(cudaMemcpyAsync(host_counts, d_c64, sizeof(long long) * nbins, cudaMemcpyDeviceToHost, ctx->stream));
    if (host_sums) FB_CUDA(cudaMemcpyAsync(host_sums, d_sums, sizeof(double) * 4 * nbins, cudaMemcpyDeviceToHost, ctx->stream));
    if (host_err) FB_CUDA(cudaMemcpyAsync(host_err, d_err, sizeof(double) * 2 * nbins, cudaMemcpyDeviceToHost, ctx->stream));
    FB_CUDA(cudaStreamSynchronize(ctx->stream));
    for (void *p : {(void *)d_uv, (void *)d_V, (void *)d_w, (void *)d_idx, (void *)d_cnt, (void *)d_off, (void *)d_c64, (void *)d_sums,
                    (void *)d_err, (void *)d_it})
        cudaFree(p);
    return 0;
}

}  // extern "C"
