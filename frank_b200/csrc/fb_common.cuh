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
    // panel decomposition
    int P = 0, ntypes = 0;
    std::vector<FbGramType> h_types;
    FbGramType *d_types = nullptr;
    int *d_tile_panel = nullptr, *d_panel_t0 = nullptr, *d_panel_nt = nullptr;
    int *d_pair_code = nullptr;    // [P * P * 3]: (first OFF type, second OFF type | -1, rows in the first) ; DIAG type on the diagonal
    // a mapping call may run as two parts (host entry point: the copy of the second half overlaps the first half's
    // kernels); each part has its own work tables and its own range of partial slots
    int *d_work2 = nullptr;
    int work2_cap = 0;
    long long part_tiles[2] = {0, 0};
    const int *part_typetab[2] = {nullptr, nullptr};
    int part_slot0[2] = {0, 0};
    cudaEvent_t pev[4] = {};
    int *d_work = nullptr;         // per launch: per-CTA item ranges, items (type, chunk, partial slot), per type (chunks, first slot)
    int work_cap = 0;
    // workspaces
    int64_t cap = 0;
    double *d_a = nullptr, *d_sw = nullptr, *d_swV = nullptr, *d_kz = nullptr;   // sorted, padded, SoA
    double *d_amid = nullptr;      // per tile of FB_TV sorted visibilities: (min a, max a)
    int sort_bits = 16;            // key bits of the baseline sort (16 or 24)
    double *d_rec = nullptr;       // unsorted records (a, sqrt w, sqrt w * Re V, kz), 32 B each
    uint64_t *d_items = nullptr;   // 2 x cap sort buffers of (key << 32 | index)
    uint32_t *d_perm = nullptr;    // sorted position -> original index
    uint32_t *d_hist = nullptr;
    size_t hist_cap = 0;
    uint32_t *d_binstart = nullptr;   // uv binner: segment starts of the sorted items [nbins + 1]
    size_t bin_cap = 0;
    double *d_red = nullptr;     // pre-pass block reductions
    int red_cap = 0;
    double *d_partial = nullptr;
    size_t partial_cap = 0;
    double *d_H2 = nullptr;
    // staging for the host entry point
    double *d_in = nullptr;
    int64_t in_cap = 0;
    double *d_out = nullptr;
    size_t out_cap = 0;
    // solver workspaces (fb_solve.cu)
    int sv_B = 0, sv_N = 0;
    double *sv_D = nullptr, *sv_p = nullptr, *sv_mu = nullptr, *sv_tr2 = nullptr, *sv_alpha = nullptr, *sv_p0 = nullptr;
    double *sv_Tinv = nullptr, *sv_M = nullptr, *sv_j = nullptr, *sv_Z = nullptr, *sv_rdiag = nullptr;
    int *sv_flags = nullptr;
    double *sv_rhs = nullptr;          // power-spectrum update: right-hand side beta + log p
    int *sv_notconv = nullptr;         // ... and its per-problem 'some entry moved by more than tol' flag
    double *sv_hist = nullptr;         // iteration history (p, mu) of fb_frank_normal_loop
    size_t sv_hist_cap = 0;
    // LogNormal model state
    int ln_N = 0;
    double *ln_S = nullptr, *ln_vec = nullptr;
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

// kernels / launchers implemented in the .cu files
int fb_launch_prep(fb_ctx *ctx, int64_t n, const double *u, const double *v, const double *V, const double *w,
                   int w_stride, const fb_geometry *g, double *dev_H0, double *host_qminmax, double *host_H0);
int fb_reserve_prep(fb_ctx *ctx, int64_t n_pad);
int fb_reserve_sort(fb_ctx *ctx, int64_t n);
int fb_launch_gram_part(fb_ctx *ctx, int part, int nparts, int64_t n, int vis_model);
int fb_launch_gram_finalize(fb_ctx *ctx, int nparts, double model_scale, double *dev_M, double *dev_j);
int fb_build_j0_table(fb_ctx *ctx, double x_max);
int fb_launch_sort(fb_ctx *ctx, int64_t n, int64_t n_pad, double a_max);
int fb_build_gram_plan(fb_ctx *ctx);
uint64_t *fb_radix_sort_items(fb_ctx *ctx, int64_t n, uint64_t *buf0, uint64_t *buf1, int nbits, int *status);
int fb_items_from_keys(fb_ctx *ctx, int64_t n, const int32_t *dev_keys, uint64_t *items);
