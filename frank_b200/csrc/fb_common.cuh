// Shared declarations for libfrankb200 (sm_100a).  Internal; the public surface is include/frankb200.h.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/frankb200.h"
#include "fb_j0_table.h"

// ---- geometry of the fused J0 + Gram kernel -------------------------------------------------
// One tile = FB_TV visibilities.  The design-matrix tile G[mode][vis] lives in shared memory with a leading
// dimension FB_LDV = FB_TV + 4 doubles so that the m8n8k4 fragment pattern (8 modes x 4 vis) touches all 32 banks
// exactly once per half-warp.
//
// The symmetric (N+1)x(N+1) Gram matrix is cut into 8x8 tiles, the tiles into panels of <= FB_PT tiles, and
// the upper triangle of the panel grid into work-item types:
//   OFF  : tile rows of panel A (all of them, or one of two halves)  x  all tile columns of panel B   (A < B)
//   DIAG : the upper triangle of one panel, executed as the skewed strip (row r, offset d) -> column (r+d) mod n
// so that every warp holds <= FB_ACC accumulator tiles (a 4 x 4 warp grid covers the block) and J0 is evaluated
// only for the <= FB_GCOLS columns the block touches.
constexpr int FB_TV = 64;
constexpr int FB_LDV = FB_TV + 4;
constexpr int FB_PT = 20;              // max 8-column tiles per panel (160 modes)
constexpr int FB_GCOLS = 256;          // columns of G held per tile
constexpr int FB_ACC = 16;             // accumulator tiles per warp
constexpr int FB_PSZ = 16 * FB_ACC * 64;         // doubles per work-item partial (16 warps x FB_ACC tiles)
constexpr int FB_GRAM_THREADS = 512;

constexpr int FB_KIND_OFF = 0;
constexpr int FB_KIND_DIAG = 1;

struct FbGramType {
    int kind;
    int a_t0, a_nt;   // OFF: tile rows [a_t0, a_t0 + a_nt)   DIAG: the panel
    int b_t0, b_nt;   // OFF: tile columns                     DIAG: unused (b_nt = 0)
    int ld;           // tiles per row of the partial block: b_nt (OFF) | a_nt / 2 + 1 (DIAG)
};

constexpr int FB_MAX_CHAN = 64;        // channels of one multi-frequency mapping call (np.unique(frequencies))
constexpr int FB_MAX_CHUNKS = 64;      // chunks of the host entry point's copy / compute pipeline

// status bits a mapping call raises on the device (no host read in the middle of a call)
constexpr int FB_ST_QRANGE = 1;        // data beyond the last collocation point (check_qbounds)
constexpr int FB_ST_TABLE = 2;         // the J0 table does not reach a_max * j_{N-1}: the call is redone with a larger table

// One in-flight chunk of a mapping call: a stream and every per-visibility workspace.  The device entry points use
// lane 0 for the whole call; the host entry point alternates between the two lanes, so that the copy and the kernels of
// consecutive chunks overlap and the tail of one chunk's Gram kernel is filled by the next chunk's kernels.
struct FbLane {
    cudaStream_t stream = nullptr;
    int64_t cap = 0;                                                   // padded visibilities the workspaces hold
    double *d_a = nullptr, *d_sw = nullptr, *d_swV = nullptr, *d_kz = nullptr;   // sorted, padded, SoA
    double *d_amid = nullptr;      // per tile of FB_TV sorted visibilities: (min a, max a)
    double *d_rec = nullptr;       // unsorted records (a, sqrt w, sqrt w * Re V, kz), 32 B each
    uint64_t *d_items = nullptr;   // 2 x cap sort buffers of (key << 32 | index)
    uint32_t *d_perm = nullptr;    // sorted position -> original index
    uint32_t *d_hist = nullptr;
    size_t hist_cap = 0;
    double *d_red = nullptr;       // pre-pass block reductions
    int red_cap = 0;
    int *d_seg = nullptr;          // [2 * (FB_MAX_CHAN + 1)]: sorted start | padded start of every channel
    double *d_partial = nullptr;   // partial blocks of this lane's Gram launches (one set per channel)
    size_t partial_cap = 0;
    double *d_in = nullptr;        // device staging of the host entry point: u | v | V | w | chan
    int64_t in_cap = 0;
    double *h_pin = nullptr;       // pinned host staging for pageable inputs (same layout as d_in)
    int64_t pin_cap = 0;
    cudaEvent_t ev_copied = nullptr, ev_done = nullptr, ev_acc = nullptr, ev_g0 = nullptr, ev_g1 = nullptr, ev_p0 = nullptr;
};

struct fb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    // DHT
    int N = 0, NT = 0, NC = 0;
    double Qmax = 0, invQmax = 0;
    std::vector<double> h_jk, h_ck;
    double *d_jk = nullptr, *d_ck = nullptr, *d_Y = nullptr;
    // J0 table
    double2 *d_tab = nullptr;
    int tab_rows = 0;
    // panel decomposition and the work table of the Gram kernel (both depend on N only: built by fb_dht_setup)
    int P = 0, ntypes = 0;
    std::vector<FbGramType> h_types;
    FbGramType *d_types = nullptr;
    int *d_tile_panel = nullptr, *d_panel_t0 = nullptr, *d_panel_nt = nullptr;
    int *d_pair_code = nullptr;    // [P * P * 3]: (first OFF type, second OFF type | -1, rows in the first) ; DIAG type on the diagonal
    int *d_work = nullptr;         // per-CTA item ranges [grid + 1], items (type, chunk, partial slot) [3 n_items], per type (chunks, first slot)
    int n_items = 0;
    int sort_bits = 16;            // key bits of the baseline sort (16 or 24)
    // mapping lanes
    FbLane lane[2];
    cudaStream_t stream_copy = nullptr;
    double *d_S = nullptr;         // [nchan][npairs][64] unscaled Gram accumulated over the chunks of a call
    size_t S_cap = 0;
    double *d_chunkred = nullptr;  // [FB_MAX_CHUNKS][4]: H0 sum, min q, max q, (unused) of every chunk
    double *d_result = nullptr;    // [4]: H0, min q, max q, status
    int *d_status = nullptr;       // status bits of the call in flight
    double *h_result = nullptr;    // pinned mirror of d_result
    uint32_t *d_binstart = nullptr;   // uv binner: segment starts of the sorted items [nbins + 1]
    size_t bin_cap = 0;
    double *d_H2 = nullptr;
    double *d_predI = nullptr;     // prediction: brightness profile [N] and behind it the table-overflow flag (2 words)
    int predI_cap = 0;
    double *d_out = nullptr;       // host entry point: M | j | H0 on the device
    size_t out_cap = 0;
    int map_chunks = 1;            // chunks of the most recent call (timing)
    std::vector<cudaEvent_t> mev;  // per chunk: (start, sorted, gram done, accumulated); two more for the copy stream
    std::vector<double> h_H2;      // debris H2 currently on the device
    int stage_threads = 8;         // host threads gathering pageable inputs into the pinned staging ring
    int64_t map_chunk = 250000;    // smallest first chunk of the host entry point's pipeline (visibilities)
    double map_growth = 1.5;       // ratio of consecutive chunk sizes (fixed: a given n always gives the same chunks, hence the same
                                   // bits); 0 = adapt to the copy / kernel rates measured on the previous call
    double rate_copy_ns = 0.0, rate_gram_ns = 0.0;      // ns per visibility measured on the previous host call, for N = rate_N
    int rate_N = 0;
    int map_kmax = 8;              // most chunks per call
    bool force_staging = false;    // treat every host input as pageable (tests)
    // multi-GPU: NCCL communicator attached by fb_comm_init (library loaded with dlopen)
    void *nccl_lib = nullptr;
    void *nccl_comm = nullptr;
    int comm_rank = 0, comm_size = 1;
    double *d_pack = nullptr;      // packed all-reduce buffer
    size_t pack_cap = 0;
    // solver workspaces (fb_solve.cu)
    int sv_B = 0, sv_N = 0;
    double *sv_D = nullptr, *sv_p = nullptr, *sv_mu = nullptr, *sv_tr2 = nullptr, *sv_alpha = nullptr, *sv_p0 = nullptr;
    double *sv_Tinv = nullptr, *sv_M = nullptr, *sv_j = nullptr, *sv_Z = nullptr, *sv_rdiag = nullptr;
    double *sv_diag = nullptr;         // factorised diagonal blocks of the Cholesky panels, parked until the last panel is done
    int *sv_flags = nullptr;
    double *sv_rhs = nullptr;          // power-spectrum update: right-hand side beta + log p
    int *sv_notconv = nullptr;         // ... and its per-problem 'some entry moved by more than tol' flag
    double *sv_hist = nullptr;         // iteration history (p, mu) of fb_frank_normal_loop
    size_t sv_hist_cap = 0;
    // LogNormal model state
    int ln_N = 0;
    double *ln_S = nullptr, *ln_vec = nullptr;
    double *ln_pin = nullptr;          // pinned (mapped) scalars of the device-resident Newton iteration
    double *ln_ws = nullptr;           // its vectors
    int ln_ws_N = 0;
    double ln_s0 = 0.0, ln_full_hess = 1.0;
    cudaEvent_t ev[8] = {};
    cudaEvent_t tev[2] = {};
    cudaStream_t stream2 = nullptr;     // fork / join branch of the solver graph
    cudaStream_t stream3 = nullptr;     // side stream of the Cholesky (rest of the trailing updates)
    std::vector<cudaEvent_t> cev;       // per block column: (panel done, rest of the update done)
    cudaEvent_t fev[2] = {};
    double timing[4] = {0, 0, 0, 0};
    int num_sms = 148;
    int64_t last_n = 0;
};

#define FB_CUDA(call)                                                                        \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            char buf__[512];                                                                 \
            snprintf(buf__, sizeof(buf__), "%s failed: %s (%s:%d)", #call,                    \
                     cudaGetErrorString(e__), __FILE__, __LINE__);                           \
            ctx->err = buf__;                                                                \
            return -(int)e__ - 1000;                                                         \
        }                                                                                    \
    } while (0)

#define FB_FAIL(code, msg)      \
    do {                        \
        ctx->err = (msg);       \
        return (code);          \
    } while (0)

// np.hypot = glibc's non-FMA kernel (sysdeps/ieee754/dbl-64/e_hypot.c, glibc >= 2.35), one correctly rounded operation at
// a time: q, and with it every J0 argument, is bit-equal to NumPy's (tests/test_gpu_mapping.py::test_prepass_bits).
#ifdef __CUDACC__
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

#endif

// kernels / launchers implemented in the .cu files
struct FbMapJob {              // arguments of one mapping call, shared by its chunks
    fb_geometry geom;
    int vis_model = 0;
    int nchan = 1;
    int check_qbounds = 0;
    double q_last = 0.0;
};
int fb_reserve_lane(fb_ctx *ctx, FbLane &ln, int64_t n, int nchan);
int fb_reserve_sort(fb_ctx *ctx, FbLane &ln, int64_t n);
int fb_enqueue_prep(fb_ctx *ctx, FbLane &ln, int chunk, int64_t n, const double *u, const double *v, const double *V,
                    const double *w, int w_stride, const int32_t *chan, const FbMapJob &job);
int fb_enqueue_sort(fb_ctx *ctx, FbLane &ln, int chunk, int64_t n, const int32_t *chan, int nchan);
int fb_enqueue_gram(fb_ctx *ctx, FbLane &ln, int chan, int vis_model);
int fb_enqueue_accumulate(fb_ctx *ctx, FbLane &ln, int nchan, int first);
int fb_enqueue_scale(fb_ctx *ctx, cudaStream_t st, int nchan, double model_scale, double *dev_M, double *dev_j);
int fb_enqueue_result(fb_ctx *ctx, cudaStream_t st, int nchunks, double *dev_H0);
int fb_build_j0_table(fb_ctx *ctx, double x_max);
int fb_build_gram_plan(fb_ctx *ctx);
int fb_items_from_keys(fb_ctx *ctx, FbLane &ln, int64_t n, const int32_t *dev_keys, uint64_t *items);
uint64_t *fb_radix_sort_items(fb_ctx *ctx, FbLane &ln, int64_t n, uint64_t *buf0, uint64_t *buf1, int nbits, int *status);
int fb_comm_allreduce_map(fb_ctx *ctx, cudaStream_t st, int nchan, double *dev_M, double *dev_j, double *dev_H0);
