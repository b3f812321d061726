// K1b: deterministic (stable) LSD radix sort of the pre-passed visibilities by baseline bin.
//
// Why sort: the J0 table lookup in the Gram kernel is a gather; with the 32 lanes of a warp holding 32
// visibilities of (nearly) the same baseline, every lane hits the same table row for a given mode and the
// gather collapses to one L1 wavefront.  The same machinery is the "sort by bin, then segmented reduce" of the
// uv-binning pass (frank/utilities.py:300-367 accumulates with np.bincount; here bins become contiguous
// segments).
//
// Key = (channel << 16) | floor(a * key_scale) clipped to 16 bits; payload = original index.  8-bit digits,
// 2 passes (3 with channels).  Each pass: per-block digit histograms -> exclusive scan over (digit, block) ->
// stable scatter.  Stability (and therefore bit-reproducible sums downstream) comes from ranking inside a
// block with __match_any_sync in a fixed element order.
#include "fb_common.cuh"

namespace {

constexpr int SORT_THREADS = 256;
constexpr int SORT_ROUNDS = 16;                              // elements per lane
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_CHUNK = SORT_THREADS * SORT_ROUNDS;       // 4096 elements per block

// element e of block b handled by (warp, round, lane): contiguous per warp, round-major inside the warp
__device__ __forceinline__ int64_t elem_index(int64_t base, int warp, int round, int lane)
{
    return base + (int64_t)warp * (32 * SORT_ROUNDS) + round * 32 + lane;
}

__device__ __forceinline__ uint64_t make_item(const double4 *__restrict__ rec, const int32_t *__restrict__ chan, int64_t i,
                                              double key_scale)
{
    int k = __double2int_rz(rec[i].x * key_scale);
    k = k < 0 ? 0 : (k > 65535 ? 65535 : k);
    uint64_t key = (uint64_t)k;
    if (chan) key |= (uint64_t)(uint32_t)chan[i] << 16;
    return (key << 32) | (uint64_t)(uint32_t)i;
}

template <bool FIRST>
__global__ void __launch_bounds__(SORT_THREADS)
k_sort_hist(int64_t n, const double4 *__restrict__ rec, const int32_t *__restrict__ chan, double key_scale,
            const uint64_t *__restrict__ items, int shift, uint32_t *__restrict__ hist, int nblocks)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * SORT_CHUNK;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; r++) {
        int64_t i = elem_index(base, warp, r, lane);
        if (i < n) {
            uint64_t it = FIRST ? make_item(rec, chan, i, key_scale) : items[i];
            atomicAdd(&h[(it >> shift) & 255], 1u);
        }
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of `total` uint32 counters, single block of 1024 threads, slab per thread
__global__ void __launch_bounds__(1024) k_sort_scan(uint32_t *__restrict__ data, int64_t total)
{
    __shared__ uint32_t warp_sums[32];
    const int64_t per = (total + 1023) / 1024;
    const int64_t b0 = (int64_t)threadIdx.x * per, b1 = b0 + per < total ? b0 + per : total;
    uint32_t s = 0;
    for (int64_t i = b0; i < b1; i++) s += data[i];
    // block exclusive scan of s
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_sums[lane], winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        warp_sums[lane] = winc - w;
    }
    __syncthreads();
    uint32_t run = warp_sums[warp] + inc - s;
    for (int64_t i = b0; i < b1; i++) {
        uint32_t c = data[i];
        data[i] = run;
        run += c;
    }
}

template <bool FIRST>
__global__ void __launch_bounds__(SORT_THREADS)
k_sort_scatter(int64_t n, const double4 *__restrict__ rec, const int32_t *__restrict__ chan, double key_scale,
               const uint64_t *__restrict__ items, int shift, const uint32_t *__restrict__ offs, int nblocks,
               uint64_t *__restrict__ out)
{
    __shared__ uint32_t cnt[SORT_WARPS][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t base = (int64_t)blockIdx.x * SORT_CHUNK;
    for (int w = 0; w < SORT_WARPS; w++) cnt[w][threadIdx.x] = 0;
    __syncthreads();

    uint64_t it[SORT_ROUNDS];
    // walk 1: per-warp digit counts
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; r++) {
        int64_t i = elem_index(base, warp, r, lane);
        const bool ok = i < n;
        it[r] = ok ? (FIRST ? make_item(rec, chan, i, key_scale) : items[i]) : ~0ull;
        const unsigned active = __ballot_sync(0xffffffffu, ok);
        if (ok) {
            const int d = (int)((it[r] >> shift) & 255);
            const unsigned m = __match_any_sync(active, d);
            if (lane == __ffs(m) - 1) cnt[warp][d] += __popc(m);
        }
        __syncwarp();
    }
    __syncthreads();
    // per digit: global base + exclusive prefix over the warps of this block
    {
        const int d = threadIdx.x;
        uint32_t run = offs[(size_t)d * nblocks + blockIdx.x];
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) {
            uint32_t c = cnt[w][d];
            cnt[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
    // walk 2: same order, stable ranks
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; r++) {
        int64_t i = elem_index(base, warp, r, lane);
        const bool ok = i < n;
        const unsigned active = __ballot_sync(0xffffffffu, ok);
        if (ok) {
            const int d = (int)((it[r] >> shift) & 255);
            const unsigned m = __match_any_sync(active, d);
            const uint32_t pos = cnt[warp][d] + __popc(m & ((1u << lane) - 1));
            out[pos] = it[r];
            __syncwarp(m);
            if (lane == __ffs(m) - 1) cnt[warp][d] += __popc(m);
        }
        __syncwarp();
    }
}

// permute the records into the structure-of-arrays layout the Gram kernel reads; zero padding to n_pad
__global__ void __launch_bounds__(256)
k_sort_gather(int64_t n, int64_t n_pad, const uint64_t *__restrict__ items, const double4 *__restrict__ rec,
              double *__restrict__ a, double *__restrict__ sw, double *__restrict__ swV, double *__restrict__ kz,
              uint32_t *__restrict__ perm)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint32_t src = (uint32_t)(items[i] & 0xffffffffull);
        const double4 r = rec[src];
        a[i] = r.x; sw[i] = r.y; swV[i] = r.z; kz[i] = r.w;
        perm[i] = src;
    } else if (i < n_pad) {
        a[i] = 0.0; sw[i] = 0.0; swV[i] = 0.0; kz[i] = 0.0;
    }
}

}  // namespace

// Sort ctx->d_rec[0..n) by key and write ctx->d_a/d_sw/d_swV/d_kz (padded to n_pad) and ctx->d_perm.
int fb_launch_sort(fb_ctx *ctx, int64_t n, int64_t n_pad, double a_max)
{
    if (n <= 0) return 0;
    const int nblocks = (int)((n + SORT_CHUNK - 1) / SORT_CHUNK);
    const size_t hist_need = (size_t)256 * nblocks;
    if (hist_need > ctx->hist_cap) {
        if (ctx->d_hist) FB_CUDA(cudaFree(ctx->d_hist));
        ctx->d_hist = nullptr;
        FB_CUDA(cudaMalloc(&ctx->d_hist, sizeof(uint32_t) * hist_need));
        ctx->hist_cap = hist_need;
    }
    const double key_scale = a_max > 0 ? 65535.5 / a_max : 0.0;
    const double4 *rec = (const double4 *)ctx->d_rec;
    uint64_t *buf0 = ctx->d_items, *buf1 = ctx->d_items + ctx->cap;
    // pass 0: low byte of the baseline bin
    k_sort_hist<true><<<nblocks, SORT_THREADS, 0, ctx->stream>>>(n, rec, nullptr, key_scale, nullptr, 32, ctx->d_hist, nblocks);
    k_sort_scan<<<1, 1024, 0, ctx->stream>>>(ctx->d_hist, (int64_t)hist_need);
    k_sort_scatter<true><<<nblocks, SORT_THREADS, 0, ctx->stream>>>(n, rec, nullptr, key_scale, nullptr, 32, ctx->d_hist, nblocks, buf0);
    // pass 1: high byte
    k_sort_hist<false><<<nblocks, SORT_THREADS, 0, ctx->stream>>>(n, rec, nullptr, key_scale, buf0, 40, ctx->d_hist, nblocks);
    k_sort_scan<<<1, 1024, 0, ctx->stream>>>(ctx->d_hist, (int64_t)hist_need);
    k_sort_scatter<false><<<nblocks, SORT_THREADS, 0, ctx->stream>>>(n, rec, nullptr, key_scale, buf0, 40, ctx->d_hist, nblocks, buf1);
    FB_CUDA(cudaGetLastError());
    k_sort_gather<<<(unsigned)((n_pad + 255) / 256), 256, 0, ctx->stream>>>(n, n_pad, buf1, rec, ctx->d_a, ctx->d_sw, ctx->d_swV,
                                                                         ctx->d_kz, ctx->d_perm);
    FB_CUDA(cudaGetLastError());
    return 0;
}
