// K1b: deterministic (stable) LSD radix sort of the pre-passed visibilities by baseline bin.
//
// Why sort: the J0 table lookup in the Gram kernel is a gather; with the 32 lanes of a warp holding 32
// visibilities of (nearly) the same baseline, every lane hits the same table row for a given mode and the
// gather collapses to one L1 wavefront.  The same machinery is the "sort by bin, then segmented reduce" of the
// uv-binning pass (frank/utilities.py:300-367 accumulates with np.bincount; here bins become contiguous
// segments).
//
// Key = floor(a * key_scale) clipped to 16 or 24 bits (Gram; 24 when the modes reach arguments so large that a
// 16-bit bin would span more than the validity window of a J0 table row) or the uv-bin index (binner); payload =
// original index.  8-bit digits, ceil(bits / 8) passes.  Each pass: per-block digit histograms -> exclusive scan over (digit, block) ->
// stable scatter.  Stability (and therefore bit-reproducible sums downstream) comes from ranking inside a
// block with __match_any_sync in a fixed element order.
#include "fb_common.cuh"

#include <algorithm>

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

// key = (channel << nbits) | floor(a * key_scale) clipped to [0, kmax], payload = index.  The scale follows the chunk's
// largest baseline, read from the chunk's device-resident reduction (red[2] = max q): no host round trip.
__global__ void __launch_bounds__(256)
k_items_from_rec(int64_t n, const double4 *__restrict__ rec, const double *__restrict__ red, double invQmax, int kmax,
                 int nbits, const int32_t *__restrict__ chan, int nchan, uint64_t *__restrict__ items)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const double a_max = red[2] * invQmax;
        const double key_scale = a_max > 0 ? ((double)kmax + 0.5) / a_max : 0.0;
        int k = __double2int_rz(rec[i].x * key_scale);
        k = k < 0 ? 0 : (k > kmax ? kmax : k);
        uint32_t key = (uint32_t)k;
        if (chan) {
            int c = chan[i];
            c = c < 0 ? 0 : (c >= nchan ? nchan - 1 : c);
            key |= (uint32_t)c << nbits;
        }
        items[i] = ((uint64_t)key << 32) | (uint64_t)(uint32_t)i;
    }
}

// key = a non-negative 32-bit integer per element (uv-bin index), payload = index
__global__ void __launch_bounds__(256)
k_items_from_keys(int64_t n, const int32_t *__restrict__ keys, uint64_t *__restrict__ items)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) items[i] = ((uint64_t)(uint32_t)keys[i] << 32) | (uint64_t)(uint32_t)i;
}

// Lanes of the warp (among `active`) whose 8-bit digit equals this lane's: eight ballots, one per digit bit (the
// MATCH.ANY instruction behind __match_any_sync measured ~4x slower than this on sm_100a: 233 -> see DESIGN.md).
// Every lane of the warp must call; lanes outside `active` get 0.
__device__ __forceinline__ unsigned match_digit(int d, unsigned active)
{
    unsigned m = active;
#pragma unroll
    for (int b = 0; b < 8; b++) {
        const bool bit = (d >> b) & 1;
        const unsigned bal = __ballot_sync(0xffffffffu, bit);
        m &= bit ? bal : ~bal;
    }
    return ((active >> (threadIdx.x & 31)) & 1u) ? m : 0u;
}

__global__ void __launch_bounds__(SORT_THREADS)
k_sort_hist(int64_t n, const uint64_t *__restrict__ items, int shift, uint32_t *__restrict__ hist, int nblocks,
            uint32_t *__restrict__ digit_total)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * SORT_CHUNK;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; r++) {
        int64_t i = elem_index(base, warp, r, lane);
        if (i < n) atomicAdd(&h[(items[i] >> shift) & 255], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
    if (h[threadIdx.x]) atomicAdd(&digit_total[threadIdx.x], h[threadIdx.x]);      // integer: order independent
}

// Exclusive scan of the (digit, block) counters in place, one CTA per digit row: the row's base is the sum of the
// totals of the smaller digits (accumulated by k_sort_hist), the row itself is scanned in coalesced chunks of 4096
// (4 per thread) with warp shuffles and a running carry.
__global__ void __launch_bounds__(1024) k_sort_scan(uint32_t *__restrict__ data, int nblocks, const uint32_t *__restrict__ digit_total)
{
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s[2];                // double-buffered: read in one chunk, written for the next
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int digit = blockIdx.x;
    {
        uint32_t b = threadIdx.x < digit ? digit_total[threadIdx.x] : 0u;        // digit < 256 <= blockDim
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
        if (lane == 0) warp_sums[warp] = b;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
            for (int w2 = 0; w2 < 8; w2++) t += warp_sums[w2];
            carry_s[0] = t;
        }
        __syncthreads();
    }
    data += (size_t)digit * nblocks;
    const int64_t total = nblocks;
    int cur = 0;
    for (int64_t base = 0; base < total; base += 4096, cur ^= 1) {
        const int64_t i0 = base + (int64_t)threadIdx.x * 4;
        uint32_t v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = i0 + k < total ? data[i0 + k] : 0u;
        const uint32_t s = v[0] + v[1] + v[2] + v[3];
        uint32_t inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        __syncthreads();                 // warp_sums of the previous chunk (or of the base) have been read
        if (lane == 31) warp_sums[warp] = inc;
        __syncthreads();
        const uint32_t carry = carry_s[cur];
        if (warp == 0) {
            uint32_t w = warp_sums[lane], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            warp_sums[lane] = winc - w;
            if (lane == 31) carry_s[cur ^ 1] = carry + winc;
        }
        __syncthreads();
        uint32_t run = carry + warp_sums[warp] + inc - s;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (i0 + k < total) data[i0 + k] = run;
            run += v[k];
        }
    }
}

// Stable scatter of one digit pass.  The block's 4096 items are first ordered by digit in shared memory (stable:
// warps, rounds and lanes in element order), then written out position by position, so that the items of one digit
// leave as one contiguous run (16 items = 128 B on average) instead of one scattered 8-byte store per item.
__global__ void __launch_bounds__(SORT_THREADS, 3)      // 80 registers: three CTAs per SM (106 registers allowed two)
k_sort_scatter(int64_t n, const uint64_t *__restrict__ items, int shift, const uint32_t *__restrict__ offs, int nblocks,
               uint64_t *__restrict__ out)
{
    __shared__ uint32_t cnt[SORT_WARPS][256];
    __shared__ uint32_t gbase[256];                 // global position of local position 0 of each digit's run
    __shared__ uint32_t wsum[SORT_WARPS];
    __shared__ uint64_t staged[SORT_CHUNK];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t base = (int64_t)blockIdx.x * SORT_CHUNK;
    for (int w = 0; w < SORT_WARPS; w++) cnt[w][threadIdx.x] = 0;
    __syncthreads();

    uint64_t it[SORT_ROUNDS];
    unsigned peers[SORT_ROUNDS];                    // lanes of the warp holding the same digit in this round
    // walk 1: per-warp digit counts
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; r++) {
        int64_t i = elem_index(base, warp, r, lane);
        const bool ok = i < n;
        it[r] = ok ? items[i] : ~0ull;
        const int d = (int)((it[r] >> shift) & 255);
        const unsigned m = match_digit(d, __ballot_sync(0xffffffffu, ok));
        peers[r] = m;
        if (ok && lane == __ffs(m) - 1) cnt[warp][d] += __popc(m);
        __syncwarp();
    }
    __syncthreads();
    // per digit: exclusive prefix over the warps of this block, then over the digits -> local start of every
    // (warp, digit) group; gbase = global offset of the digit in this block - local start of the digit
    {
        const int d = threadIdx.x;
        uint32_t c[SORT_WARPS], tot = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) { c[w] = cnt[w][d]; tot += c[w]; }
        uint32_t inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        uint32_t before = 0;
        for (int w = 0; w < warp; w++) before += wsum[w];
        uint32_t run = before + inc - tot;           // local start of digit d
        gbase[d] = offs[(size_t)d * nblocks + blockIdx.x] - run;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) { cnt[w][d] = run; run += c[w]; }
    }
    __syncthreads();
    // walk 2: same order, stable local ranks
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; r++) {
        const unsigned m = peers[r];
        if (m) {
            const int d = (int)((it[r] >> shift) & 255);
            staged[cnt[warp][d] + __popc(m & ((1u << lane) - 1))] = it[r];
        }
        __syncwarp();
        if (m && lane == __ffs(m) - 1) cnt[warp][(int)((it[r] >> shift) & 255)] += __popc(m);
        __syncwarp();
    }
    __syncthreads();
    const int nvalid = (int)(n - base < SORT_CHUNK ? n - base : SORT_CHUNK);
    for (int k = threadIdx.x; k < nvalid; k += SORT_THREADS) {
        const uint64_t x = staged[k];
        out[gbase[(int)((x >> shift) & 255)] + (uint32_t)k] = x;
    }
}

// Channel segments of the sorted items (multi-frequency calls: the channel is the high part of the key, so every
// channel is one contiguous run): start[c] = first sorted position of a channel >= c, c = 0 .. nchan.
__global__ void __launch_bounds__(256)
k_chan_starts(int64_t n, const uint64_t *__restrict__ sorted, int nbits, int nchan, int *__restrict__ start)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
        const int c = (int)(sorted[p] >> (32 + nbits));
        const int prev = p > 0 ? (int)(sorted[p - 1] >> (32 + nbits)) : -1;
        for (int b = prev + 1; b <= c; b++) start[b] = (int)p;
        if (p == n - 1)
            for (int b = c + 1; b <= nchan; b++) start[b] = (int)n;
    }
}

// seg = start[0 .. nchan] | pad[0 .. nchan]: every channel's run is laid out from a tile boundary, pad[c] = sum over the
// channels before c of their counts rounded up to whole tiles.  The Gram kernel reads its tile range from here.
__global__ void k_seg_finish(int64_t n, int nchan, int *__restrict__ seg)
{
    int *start = seg, *pad = seg + FB_MAX_CHAN + 1;
    if (nchan == 1) { start[0] = 0; start[1] = (int)n; }
    if (n == 0) for (int c = 0; c <= nchan; c++) start[c] = 0;
    int acc = 0;
    for (int c = 0; c < nchan; c++) {
        pad[c] = acc;
        acc += (start[c + 1] - start[c] + FB_TV - 1) / FB_TV * FB_TV;
    }
    pad[nchan] = acc;
}

// permute the records into the structure-of-arrays layout the Gram kernel reads; zero padding at the end of every
// channel's run.  A block covers 4 tiles of FB_TV = 64 slots: it also records each tile's range (min a, max a), from
// which the Gram kernel picks one J0 table row per (mode, tile).
__global__ void __launch_bounds__(256)
k_sort_gather(int64_t n_slots, int nchan, const int *__restrict__ seg, const uint64_t *__restrict__ items,
              const double4 *__restrict__ rec, double *__restrict__ a, double *__restrict__ sw, double *__restrict__ swV,
              double *__restrict__ kz, uint32_t *__restrict__ perm, double *__restrict__ amid)
{
    static_assert(FB_TV == 64, "two warps per tile");
    __shared__ double s_lo[8], s_hi[8];
    __shared__ int s_seg[2 * (FB_MAX_CHAN + 1)];
    for (int k = threadIdx.x; k < 2 * (FB_MAX_CHAN + 1); k += blockDim.x) s_seg[k] = seg[k];
    __syncthreads();
    const int *start = s_seg, *pad = s_seg + FB_MAX_CHAN + 1;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = pad[nchan];
    double lo = INFINITY, hi = -INFINITY;
    if (i < total) {
        int c = 0;
        while (c + 1 < nchan && i >= pad[c + 1]) c++;
        const int64_t p = (int64_t)start[c] + (i - pad[c]);
        if (p < start[c + 1]) {
            const uint32_t src = (uint32_t)(items[p] & 0xffffffffull);
            const double4 r = rec[src];
            a[i] = r.x; sw[i] = r.y; swV[i] = r.z; kz[i] = r.w;
            perm[i] = src;
            lo = hi = r.x;
        } else {
            a[i] = 0.0; sw[i] = 0.0; swV[i] = 0.0; kz[i] = 0.0;
            perm[i] = 0xffffffffu;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { s_lo[warp] = lo; s_hi[warp] = hi; }
    __syncthreads();
    if (threadIdx.x < 4) {
        const int64_t tile = (int64_t)blockIdx.x * 4 + threadIdx.x;
        if (tile * FB_TV < total && tile * FB_TV < n_slots) {
            const double l = fmin(s_lo[2 * threadIdx.x], s_lo[2 * threadIdx.x + 1]);
            const double h = fmax(s_hi[2 * threadIdx.x], s_hi[2 * threadIdx.x + 1]);
            amid[2 * tile] = h >= l ? l : 0.0;
            amid[2 * tile + 1] = h >= l ? h : 0.0;
        }
    }
}

}  // namespace

// Digit-histogram workspace for sorting n items.
int fb_reserve_sort(fb_ctx *ctx, FbLane &ln, int64_t n)
{
    const size_t hist_need = (size_t)256 * ((n + SORT_CHUNK - 1) / SORT_CHUNK);
    if (hist_need + 256 > ln.hist_cap) {                     // + 256 digit totals behind the table
        cudaStreamSynchronize(ln.stream);
        if (ln.d_hist) cudaFree(ln.d_hist);
        ln.d_hist = nullptr;
        const size_t cap = hist_need + hist_need / 4 + 4096;
        if (cudaMalloc(&ln.d_hist, sizeof(uint32_t) * cap) != cudaSuccess) FB_FAIL(-50, "radix sort: out of memory");
        ln.hist_cap = cap;
    }
    return 0;
}

// Stable LSD radix sort of n items (key << 32 | index) on `nbits` key bits, 8 bits per pass, ping-ponging between
// buf0 (input) and buf1, on the lane's stream.  Returns the buffer holding the sorted items.
uint64_t *fb_radix_sort_items(fb_ctx *ctx, FbLane &ln, int64_t n, uint64_t *buf0, uint64_t *buf1, int nbits, int *status)
{
    *status = 0;
    const int nblocks = (int)((n + SORT_CHUNK - 1) / SORT_CHUNK);
    const size_t hist_need = (size_t)256 * nblocks;
    if (fb_reserve_sort(ctx, ln, n)) { *status = -50; return nullptr; }
    uint64_t *src = buf0, *dst = buf1;
    for (int shift = 32; shift < 32 + nbits; shift += 8) {
        uint32_t *digit_total = ln.d_hist + hist_need;            // 256 counters behind the (digit, block) table
        if (cudaMemsetAsync(digit_total, 0, sizeof(uint32_t) * 256, ln.stream) != cudaSuccess) { *status = -51; ctx->err = "radix sort: memset failed"; return nullptr; }
        k_sort_hist<<<nblocks, SORT_THREADS, 0, ln.stream>>>(n, src, shift, ln.d_hist, nblocks, digit_total);
        k_sort_scan<<<256, 1024, 0, ln.stream>>>(ln.d_hist, nblocks, digit_total);
        k_sort_scatter<<<nblocks, SORT_THREADS, 0, ln.stream>>>(n, src, shift, ln.d_hist, nblocks, dst);
        uint64_t *t = src; src = dst; dst = t;
    }
    if (cudaGetLastError() != cudaSuccess) { *status = -51; ctx->err = "radix sort: launch failed"; return nullptr; }
    return src;
}

int fb_items_from_keys(fb_ctx *ctx, FbLane &ln, int64_t n, const int32_t *dev_keys, uint64_t *items)
{
    k_items_from_keys<<<(unsigned)((n + 255) / 256), 256, 0, ln.stream>>>(n, dev_keys, items);
    FB_CUDA(cudaGetLastError());
    return 0;
}

// Sort the lane's records [0, n) by (channel, baseline bin) and write the padded SoA arrays, the permutation, the
// per-tile ranges and the channel segments.  Enqueue only.
int fb_enqueue_sort(fb_ctx *ctx, FbLane &ln, int chunk, int64_t n, const int32_t *chan, int nchan)
{
    const int nbits = ctx->sort_bits;
    const int kmax = (1 << nbits) - 1;
    int chanbits = 0;
    while ((1 << chanbits) < nchan) chanbits++;
    const double4 *rec = (const double4 *)ln.d_rec;
    uint64_t *buf0 = ln.d_items, *buf1 = ln.d_items + ln.cap;
    const uint64_t *sorted = buf0;
    if (n > 0) {
        k_items_from_rec<<<(unsigned)((n + 255) / 256), 256, 0, ln.stream>>>(n, rec, ctx->d_chunkred + 4 * chunk, ctx->invQmax, kmax,
                                                                            nbits, nchan > 1 ? chan : nullptr, nchan, buf0);
        int st = 0;
        sorted = fb_radix_sort_items(ctx, ln, n, buf0, buf1, nbits + chanbits, &st);
        if (st) return st;
        if (nchan > 1) {
            const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->num_sms * 16);
            k_chan_starts<<<grid, 256, 0, ln.stream>>>(n, sorted, nbits, nchan, ln.d_seg);
        }
    }
    k_seg_finish<<<1, 1, 0, ln.stream>>>(n, nchan, ln.d_seg);
    const int64_t n_slots = (n + FB_TV - 1) / FB_TV * FB_TV + (int64_t)FB_TV * (nchan - 1);
    if (n_slots > 0)
        k_sort_gather<<<(unsigned)((n_slots + 255) / 256), 256, 0, ln.stream>>>(n_slots, nchan, ln.d_seg, sorted, rec, ln.d_a, ln.d_sw,
                                                                              ln.d_swV, ln.d_kz, ln.d_perm, ln.d_amid);
    FB_CUDA(cudaGetLastError());
    return 0;
}
