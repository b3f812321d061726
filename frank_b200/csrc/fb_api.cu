// C ABI of libfrankb200 (see include/frankb200.h): context, DHT setup, map_visibilities entry points.
#include "fb_common.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>

extern "C" {

int fb_version(void) { return 100; }

int fb_ctx_create(fb_ctx **out, int device)
{
    if (!out) return -1;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return -2;   // no CUDA device: fail loudly
    if (device < 0 || device >= count) return -3;
    fb_ctx *ctx = new fb_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return -4; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return -5; }
    if (prop.major < 10) { delete ctx; return -6; }                            // sm_100a only
    ctx->num_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return -7; }
    ctx->lane[0].stream = ctx->stream;
    if (cudaStreamCreateWithFlags(&ctx->lane[1].stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return -7; }
    if (cudaStreamCreateWithFlags(&ctx->stream_copy, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return -7; }
    for (auto &e : ctx->ev)
        if (cudaEventCreate(&e) != cudaSuccess) { delete ctx; return -8; }
    for (FbLane &ln : ctx->lane)
        for (cudaEvent_t *e : {&ln.ev_copied, &ln.ev_done, &ln.ev_acc})
            if (cudaEventCreateWithFlags(e, cudaEventDisableTiming) != cudaSuccess) { delete ctx; return -8; }
    if (cudaMalloc(&ctx->d_chunkred, sizeof(double) * 4 * FB_MAX_CHUNKS) != cudaSuccess ||
        cudaMalloc(&ctx->d_result, sizeof(double) * 8) != cudaSuccess || cudaMalloc(&ctx->d_status, sizeof(int) * 4) != cudaSuccess ||
        cudaHostAlloc(&ctx->h_result, sizeof(double) * 8, cudaHostAllocDefault) != cudaSuccess) { delete ctx; return -9; }
    cudaMemset(ctx->d_status, 0, sizeof(int) * 4);
    {   // host threads for the staging gather: at most 8, and a fair share of the cores when several ranks share the host
        int share = (int)std::thread::hardware_concurrency();
        if (const char *e = getenv("LOCAL_WORLD_SIZE")) share /= std::max(1, atoi(e));
        ctx->stage_threads = std::max(1, std::min(8, share));
    }
    if (const char *e = getenv("FB_STAGE_THREADS")) ctx->stage_threads = std::max(1, atoi(e));
    if (const char *e = getenv("FB_MAP_CHUNK")) ctx->map_chunk = std::max<int64_t>(FB_TV, (int64_t)atof(e));
    if (const char *e = getenv("FB_MAP_GROWTH")) ctx->map_growth = atof(e) < 1.0 ? 0.0 : atof(e);
    if (const char *e = getenv("FB_MAP_KMAX")) ctx->map_kmax = std::max(1, atoi(e));
    *out = ctx;
    return 0;
}

int fb_ctx_destroy(fb_ctx *ctx)
{
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    fb_comm_destroy(ctx);
    for (FbLane &ln : ctx->lane) {
        for (void *p : {(void *)ln.d_a, (void *)ln.d_sw, (void *)ln.d_swV, (void *)ln.d_kz, (void *)ln.d_amid, (void *)ln.d_rec,
                        (void *)ln.d_items, (void *)ln.d_perm, (void *)ln.d_hist, (void *)ln.d_red, (void *)ln.d_seg,
                        (void *)ln.d_partial, (void *)ln.d_in})
            if (p) cudaFree(p);
        if (ln.h_pin) cudaFreeHost(ln.h_pin);
        for (cudaEvent_t e : {ln.ev_copied, ln.ev_done, ln.ev_acc})
            if (e) cudaEventDestroy(e);
    }
    for (void *p : {(void *)ctx->d_jk, (void *)ctx->d_ck, (void *)ctx->d_Y, (void *)ctx->d_tab, (void *)ctx->d_types,
                    (void *)ctx->d_tile_panel, (void *)ctx->d_panel_t0, (void *)ctx->d_panel_nt, (void *)ctx->d_pair_code,
                    (void *)ctx->d_work, (void *)ctx->d_S, (void *)ctx->d_chunkred, (void *)ctx->d_result, (void *)ctx->d_status,
                    (void *)ctx->d_H2, (void *)ctx->d_predI, (void *)ctx->d_out, (void *)ctx->d_binstart, (void *)ctx->d_pack,
                    (void *)ctx->sv_D, (void *)ctx->sv_p, (void *)ctx->sv_mu, (void *)ctx->sv_tr2, (void *)ctx->sv_alpha,
                    (void *)ctx->sv_p0, (void *)ctx->sv_Tinv, (void *)ctx->sv_M, (void *)ctx->sv_j, (void *)ctx->sv_Z,
                    (void *)ctx->sv_flags, (void *)ctx->sv_hist, (void *)ctx->sv_rhs, (void *)ctx->sv_notconv, (void *)ctx->sv_rdiag, (void *)ctx->sv_diag, (void *)ctx->ln_S, (void *)ctx->ln_vec, (void *)ctx->ln_ws})
        if (p) cudaFree(p);
    if (ctx->h_result) cudaFreeHost(ctx->h_result);
    if (ctx->ln_pin) cudaFreeHost(ctx->ln_pin);
    for (auto &e : ctx->ev)
        if (e) cudaEventDestroy(e);
    for (auto &e : ctx->cev)
        if (e) cudaEventDestroy(e);
    for (auto &e : ctx->mev)
        if (e) cudaEventDestroy(e);
    if (ctx->stream3) cudaStreamDestroy(ctx->stream3);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    if (ctx->stream_copy) cudaStreamDestroy(ctx->stream_copy);
    if (ctx->lane[1].stream) cudaStreamDestroy(ctx->lane[1].stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return 0;
}

const char *fb_last_error(fb_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int fb_set_option(fb_ctx *ctx, const char *name, double value)
{
    if (!ctx || !name) return -1;
    const std::string k(name);
    if (k == "map_chunk") { ctx->map_chunk = std::max<int64_t>(FB_TV, (int64_t)value); return 0; }
    if (k == "map_growth") { ctx->map_growth = value < 1.0 ? 0.0 : value; return 0; }      // 0: adaptive
    if (k == "map_kmax") { ctx->map_kmax = std::max(1, (int)value); return 0; }
    if (k == "force_staging") { ctx->force_staging = value != 0.0; return 0; }
    if (k == "stage_threads") { ctx->stage_threads = std::max(1, (int)value); return 0; }
    FB_FAIL(-21, "fb_set_option: unknown option");
}

int fb_dht_setup(fb_ctx *ctx, int N, double Qmax, const double *host_j_nk, const double *host_coef,
                 const double *host_Ycoef, double x_max)
{
    if (!ctx) return -1;
    if (N < 1 || !host_j_nk || !host_coef || !(Qmax > 0)) FB_FAIL(-10, "fb_dht_setup: bad arguments");
    FB_CUDA(cudaSetDevice(ctx->device));
    FB_CUDA(cudaDeviceSynchronize());
    ctx->N = N;
    ctx->NT = (N + 1 + 7) / 8;
    ctx->NC = ctx->NT * 8;
    ctx->Qmax = Qmax;
    ctx->invQmax = 1.0 / Qmax;                    // hankel.py:189  k = 1. / self._Qmax
    ctx->h_jk.assign(ctx->NC, 0.0);
    ctx->h_ck.assign(ctx->NC, 0.0);
    ctx->h_H2.clear();
    for (int k = 0; k < N; k++) { ctx->h_jk[k] = host_j_nk[k]; ctx->h_ck[k] = host_coef[k]; }
    for (double **p : {&ctx->d_jk, &ctx->d_ck, &ctx->d_Y, &ctx->d_H2}) {
        if (*p) FB_CUDA(cudaFree(*p));
        *p = nullptr;
    }
    FB_CUDA(cudaMalloc(&ctx->d_jk, sizeof(double) * ctx->NC));
    FB_CUDA(cudaMalloc(&ctx->d_ck, sizeof(double) * ctx->NC));
    FB_CUDA(cudaMalloc(&ctx->d_H2, sizeof(double) * ctx->NC));
    FB_CUDA(cudaMemset(ctx->d_H2, 0, sizeof(double) * ctx->NC));
    FB_CUDA(cudaMemcpy(ctx->d_jk, ctx->h_jk.data(), sizeof(double) * ctx->NC, cudaMemcpyHostToDevice));
    FB_CUDA(cudaMemcpy(ctx->d_ck, ctx->h_ck.data(), sizeof(double) * ctx->NC, cudaMemcpyHostToDevice));
    if (host_Ycoef) {
        FB_CUDA(cudaMalloc(&ctx->d_Y, sizeof(double) * (size_t)N * N));
        FB_CUDA(cudaMemcpy(ctx->d_Y, host_Ycoef, sizeof(double) * (size_t)N * N, cudaMemcpyHostToDevice));
    }

    // block decomposition of the NT x NT tile grid and the Gram kernel's work table (fb_gram.cu)
    {
        int rc = fb_build_gram_plan(ctx);
        if (rc) return rc;
    }
    // Baseline sort resolution: the arguments a * j_k of one stage must fit the validity window of one J0 table
    // row (slack 1/32 either side), so a sort bin may span at most a fraction of it at the largest mode.
    ctx->sort_bits = host_j_nk[N - 1] * 2.0 / 65536.0 > 0.03 ? 24 : 16;

    if (!(x_max > 0)) x_max = host_j_nk[N - 1];
    return fb_build_j0_table(ctx, x_max);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// The mapping engine.  A call is a sequence of chunks; a chunk is enqueued on one lane (stream + workspaces) as
//     k_prep -> k_prep_reduce -> sort by (channel, baseline) -> per channel k_gram -> k_gram_accumulate
// without any host read: the range check, the J0-table check, the sort scale and the channels' tile ranges all live in
// device memory (status bits, chunk reductions, segments).  The end of the call combines the chunks (k_map_result),
// scales S into M and j, optionally all-reduces over the attached communicator, and only then is anything read back.
// ---------------------------------------------------------------------------------------------------------------------
namespace {

struct MapCall {
    int64_t n;
    const double *u, *v, *V, *w;
    int w_stride;
    const int32_t *chan;
    FbMapJob job;
    double model_scale;
};

int map_check_args(fb_ctx *ctx, int64_t n, const fb_geometry *geom, const void *M, const void *j, const void *H0,
                   const int32_t *chan, int nchan, int vis_model, const double *host_H2)
{
    if (ctx->N == 0) FB_FAIL(-11, "fb_map_visibilities: fb_dht_setup has not been called");
    if (n < 0 || !geom || !M || !j || !H0) FB_FAIL(-12, "fb_map_visibilities: bad arguments");
    if (nchan < 1 || nchan > FB_MAX_CHAN) FB_FAIL(-13, "fb_map_visibilities: nchan must be in [1, 64]");
    if (nchan > 1 && !chan) FB_FAIL(-13, "fb_map_visibilities: nchan > 1 needs the channel index array");
    if (n > 0x7fffff00LL) FB_FAIL(-19, "fb_map_visibilities: more than 2^31 visibilities per call (shard the call)");
    if (vis_model < 0 || vis_model > 2) FB_FAIL(-14, "fb_map_visibilities: vis_model must be 0, 1 or 2");
    if (vis_model == FB_MODEL_DEBRIS) {
        if (!host_H2) FB_FAIL(-15, "fb_map_visibilities: debris model needs H2");
        std::vector<double> h2(ctx->NC, 0.0);
        for (int k = 0; k < ctx->N; k++) h2[k] = host_H2[k];
        if (h2 != ctx->h_H2) {                               // uploaded once per scale-height profile
            FB_CUDA(cudaDeviceSynchronize());
            FB_CUDA(cudaMemcpy(ctx->d_H2, h2.data(), sizeof(double) * ctx->NC, cudaMemcpyHostToDevice));
            ctx->h_H2 = h2;
        }
    }
    // unscaled Gram accumulator of the call
    const size_t need = (size_t)nchan * (ctx->NT * (ctx->NT + 1) / 2) * 64;
    if (need > ctx->S_cap) {
        FB_CUDA(cudaDeviceSynchronize());
        if (ctx->d_S) FB_CUDA(cudaFree(ctx->d_S));
        ctx->d_S = nullptr;
        FB_CUDA(cudaMalloc(&ctx->d_S, sizeof(double) * need));
        ctx->S_cap = need;
    }
    while (ctx->mev.size() < 4 * (size_t)FB_MAX_CHUNKS + 4) {
        cudaEvent_t e;
        FB_CUDA(cudaEventCreate(&e));
        ctx->mev.push_back(e);
    }
    return 0;
}

int reserve_partials(fb_ctx *ctx, FbLane &ln, int nchan)
{
    const size_t need = (size_t)nchan * ctx->n_items * FB_PSZ;
    if (need > ln.partial_cap) {
        FB_CUDA(cudaStreamSynchronize(ln.stream));
        if (ln.d_partial) FB_CUDA(cudaFree(ln.d_partial));
        ln.d_partial = nullptr;
        FB_CUDA(cudaMalloc(&ln.d_partial, need * sizeof(double)));
        ln.partial_cap = need;
    }
    return 0;
}

// One chunk on one lane.  `prev` = the lane of the previous chunk (its accumulate precedes this one's), or null.
int enqueue_chunk(fb_ctx *ctx, FbLane &ln, int chunk, int64_t n, const double *u, const double *v, const double *V,
                  const double *w, int w_stride, const int32_t *chan, const FbMapJob &job, FbLane *prev)
{
    cudaEvent_t *ev = &ctx->mev[4 * (size_t)chunk];
    FB_CUDA(cudaEventRecord(ev[0], ln.stream));
    int rc = fb_enqueue_prep(ctx, ln, chunk, n, u, v, V, w, w_stride, chan, job);
    if (rc) return rc;
    rc = fb_enqueue_sort(ctx, ln, chunk, n, chan, job.nchan);
    if (rc) return rc;
    FB_CUDA(cudaEventRecord(ev[1], ln.stream));
    for (int c = 0; c < job.nchan; c++) {
        rc = fb_enqueue_gram(ctx, ln, c, job.vis_model);
        if (rc) return rc;
    }
    FB_CUDA(cudaEventRecord(ev[2], ln.stream));
    if (prev && prev != &ln) FB_CUDA(cudaStreamWaitEvent(ln.stream, prev->ev_acc, 0));
    rc = fb_enqueue_accumulate(ctx, ln, job.nchan, chunk == 0);
    if (rc) return rc;
    FB_CUDA(cudaEventRecord(ln.ev_acc, ln.stream));
    FB_CUDA(cudaEventRecord(ev[3], ln.stream));
    FB_CUDA(cudaEventRecord(ln.ev_done, ln.stream));
    return 0;
}

// End of a call on the main stream: chunk reductions -> result, S -> M, j, the collective, the result read-back.
int enqueue_finish(fb_ctx *ctx, int nchunks, int nchan, double model_scale, double *dev_M, double *dev_j, double *dev_H0)
{
    int rc = fb_enqueue_result(ctx, ctx->stream, nchunks, dev_H0);
    if (rc) return rc;
    rc = fb_enqueue_scale(ctx, ctx->stream, nchan, model_scale, dev_M, dev_j);
    if (rc) return rc;
    if (ctx->nccl_comm) {
        rc = fb_comm_allreduce_map(ctx, ctx->stream, nchan, dev_M, dev_j, dev_H0);
        if (rc) return rc;
    }
    FB_CUDA(cudaMemcpyAsync(ctx->h_result, ctx->d_result, sizeof(double) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    return 0;
}

void collect_timing(fb_ctx *ctx, int nchunks, float copy_ms)
{
    float prep = 0, gram = 0, fin = 0;
    for (int c = 0; c < nchunks; c++) {
        float t = 0;
        cudaEvent_t *ev = &ctx->mev[4 * (size_t)c];
        if (cudaEventElapsedTime(&t, ev[0], ev[1]) == cudaSuccess) prep += t;
        if (cudaEventElapsedTime(&t, ev[1], ev[2]) == cudaSuccess) gram += t;
        if (cudaEventElapsedTime(&t, ev[2], ev[3]) == cudaSuccess) fin += t;
    }
    ctx->timing[0] = prep; ctx->timing[1] = gram; ctx->timing[2] = fin; ctx->timing[3] = copy_ms;
}

// Status of the finished call from the pinned result block; grows the J0 table when the data outran it.
// Returns 0, FB_E_QRANGE, or FB_E_RETRY (table rebuilt: the caller resubmits).
int map_status(fb_ctx *ctx, double *host_qminmax)
{
    const int st = (int)ctx->h_result[3];
    if (host_qminmax) { host_qminmax[0] = ctx->h_result[1]; host_qminmax[1] = ctx->h_result[2]; }
    if (st & FB_ST_QRANGE)
        FB_FAIL(FB_E_QRANGE, "last collocation point is at a shorter baseline than the longest deprojected baseline");
    if (st & FB_ST_TABLE) {
        const double xneed = ctx->h_result[2] * ctx->invQmax * ctx->h_jk[ctx->N - 1];
        int rc = fb_build_j0_table(ctx, xneed * 1.05);
        if (rc) return rc;
        return FB_E_RETRY;
    }
    return 0;
}

int map_begin(fb_ctx *ctx)
{
    FB_CUDA(cudaMemsetAsync(ctx->d_status, 0, sizeof(int) * 4, ctx->stream));
    FB_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    return 0;
}

// Gather the pieces of one chunk (u, v, V, w, chan) from pageable user memory into the pinned staging slot with T host
// threads: ONE parallel region per chunk over 1 MB blocks of all pieces (the OpenMP runtime keeps its workers alive
// between chunks, so there is no thread start-up on the critical path).
struct CopyPiece { void *dst; const void *src; size_t bytes; };

void parallel_gather(const CopyPiece *pieces, int npieces, int T)
{
    constexpr size_t BLK = (size_t)1 << 20;
    size_t nblk[8], total = 0;
    for (int p = 0; p < npieces; p++) { nblk[p] = (pieces[p].bytes + BLK - 1) / BLK; total += nblk[p]; }
    if (total <= 4 || T <= 1) {
        for (int p = 0; p < npieces; p++) memcpy(pieces[p].dst, pieces[p].src, pieces[p].bytes);
        return;
    }
#pragma omp parallel for num_threads(T) schedule(static)
    for (long long b = 0; b < (long long)total; b++) {
        size_t r = (size_t)b;
        int p = 0;
        while (r >= nblk[p]) { r -= nblk[p]; p++; }
        const size_t off = r * BLK, len = std::min(BLK, pieces[p].bytes - off);
        memcpy((char *)pieces[p].dst + off, (const char *)pieces[p].src + off, len);
    }
}

bool is_pinned(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

}  // namespace

extern "C" {

int fb_map_visibilities_dev_async(fb_ctx *ctx, int64_t n, const double *dev_u, const double *dev_v, const double *dev_V_reim,
                                  const double *dev_w, int w_stride, const int32_t *dev_chan, int nchan, const fb_geometry *geom,
                                  int vis_model, double model_scale, const double *host_H2, int check_qbounds, double q_last,
                                  double *dev_M, double *dev_j, double *dev_H0)
{
    if (!ctx) return -1;
    FB_CUDA(cudaSetDevice(ctx->device));
    int rc = map_check_args(ctx, n, geom, dev_M, dev_j, dev_H0, dev_chan, nchan, vis_model, host_H2);
    if (rc) return rc;
    FbLane &ln = ctx->lane[0];
    rc = fb_reserve_lane(ctx, ln, n, nchan);
    if (rc) return rc;
    rc = reserve_partials(ctx, ln, nchan);
    if (rc) return rc;
    FbMapJob job;
    job.geom = *geom; job.vis_model = vis_model; job.nchan = nchan; job.check_qbounds = check_qbounds; job.q_last = q_last;
    rc = map_begin(ctx);
    if (rc) return rc;
    rc = enqueue_chunk(ctx, ln, 0, n, dev_u, dev_v, dev_V_reim, dev_w, w_stride, dev_chan, job, nullptr);
    if (rc) return rc;
    rc = enqueue_finish(ctx, 1, nchan, model_scale, dev_M, dev_j, dev_H0);
    if (rc) return rc;
    ctx->map_chunks = 1;
    return 0;
}

int fb_map_sync(fb_ctx *ctx, double *host_qminmax)
{
    if (!ctx) return -1;
    FB_CUDA(cudaSetDevice(ctx->device));
    FB_CUDA(cudaStreamSynchronize(ctx->stream));
    collect_timing(ctx, ctx->map_chunks, 0.f);
    return map_status(ctx, host_qminmax);
}

int fb_map_visibilities_dev(fb_ctx *ctx, int64_t n, const double *dev_u, const double *dev_v, const double *dev_V_reim,
                            const double *dev_w, int w_stride, const int32_t *dev_chan, int nchan, const fb_geometry *geom,
                            int vis_model, double model_scale, const double *host_H2, int check_qbounds, double q_last,
                            double *dev_M, double *dev_j, double *dev_H0, double *host_qminmax)
{
    if (!ctx) return -1;
    if (!host_qminmax) FB_FAIL(-12, "fb_map_visibilities: bad arguments");
    for (int attempt = 0; attempt < 2; attempt++) {
        int rc = fb_map_visibilities_dev_async(ctx, n, dev_u, dev_v, dev_V_reim, dev_w, w_stride, dev_chan, nchan, geom, vis_model,
                                               model_scale, host_H2, check_qbounds, q_last, dev_M, dev_j, dev_H0);
        if (rc) return rc;
        rc = fb_map_sync(ctx, host_qminmax);
        if (rc != FB_E_RETRY) return rc;
    }
    FB_FAIL(-17, "fb_map_visibilities: J0 table could not be grown to cover the data");
}

// Host entry point: a K-deep pipeline over two lanes.  Chunk k + 1 is copied (pageable inputs: first gathered into a
// pinned staging slot by a few host threads) while chunk k is in the pre-pass / sort / Gram kernels, and the kernels of
// consecutive chunks run on alternating streams, so that the tail of one Gram launch is filled by the next chunk's
// kernels.  Only the first chunk's copy is exposed, and the copy engine needs to sustain just (bytes of the call) /
// (kernel time of the call) -- about 10 GB/s at N = 300 -- instead of the burst a two-part split asks for, which is what
// keeps eight ranks sharing one host memory system in step.  Chunks are folded into S in chunk order: deterministic.
int fb_map_visibilities_host(fb_ctx *ctx, int64_t n, const double *host_u, const double *host_v, const double *host_V_reim,
                             const double *host_w, int w_stride, const int32_t *host_chan, int nchan, const fb_geometry *geom,
                             int vis_model, double model_scale, const double *host_H2, int check_qbounds, double q_last,
                             double *host_M, double *host_j, double *host_H0, double *host_qminmax)
{
    if (!ctx) return -1;
    if (!host_qminmax) FB_FAIL(-12, "fb_map_visibilities: bad arguments");
    FB_CUDA(cudaSetDevice(ctx->device));
    int rc = map_check_args(ctx, n, geom, host_M, host_j, host_H0, host_chan, nchan, vis_model, host_H2);
    if (rc) return rc;
    const size_t N = ctx->N, nout = (size_t)nchan * (N * N + N) + 1;
    if (nout > ctx->out_cap) {
        FB_CUDA(cudaDeviceSynchronize());
        if (ctx->d_out) FB_CUDA(cudaFree(ctx->d_out));
        ctx->d_out = nullptr;
        FB_CUDA(cudaMalloc(&ctx->d_out, sizeof(double) * nout));
        ctx->out_cap = nout;
    }
    // Chunking: sizes grow geometrically (factor map_growth) from a first chunk of at least map_chunk visibilities, at most
    // map_kmax chunks.  The first chunk's copy is the exposed one, so it is small; each later copy (and, for pageable
    // inputs, the host-side gather into the staging ring) has the previous chunk's kernels to hide under.
    // The floor of the first chunk scales with the work per visibility and the channels: a chunk must give every one of the
    // Gram kernel's ~900 work items of every channel several tiles, and at large N the copy is a vanishing share of the call
    // anyway (N = 2000, 4 channels: 220 ns of kernels against 0.8 ns of copy per visibility -> one or two chunks).
    const double nn = (double)ctx->N / 300.0;
    const int64_t chunk_min = (int64_t)((double)ctx->map_chunk * nchan * std::max(1.0, nn * nn));
    int K = 1;
    std::vector<int64_t> csize(1, n);
    if (n >= 2 * chunk_min) {
        // growth: the configured factor, or (map_growth = 0, the default) what the rates measured on the previous call
        // allow -- a chunk's copy has to fit under its predecessor's kernels: ratio <= 0.85 x (kernel time per visibility) /
        // (copy or staging time per visibility), between 1.25 and 3; 2 before anything has been measured
        double r = ctx->map_growth;
        if (!(r >= 1.0)) {
            r = 2.0;
            if (ctx->rate_copy_ns > 0.0 && ctx->rate_gram_ns > 0.0 && ctx->rate_N == ctx->N)
                r = std::min(3.0, std::max(1.25, 0.85 * ctx->rate_gram_ns / ctx->rate_copy_ns));
        }
        K = std::max(1, std::min(ctx->map_kmax, FB_MAX_CHUNKS));
        auto first = [&](int k) { return r == 1.0 ? (double)n / k : (double)n * (r - 1.0) / (std::pow(r, k) - 1.0); };
        while (K > 1 && first(K) < (double)chunk_min) K--;
        csize.assign(K, 0);
        double c = first(K);
        int64_t used = 0;
        for (int k = 0; k < K; k++, c *= r) {
            csize[k] = k == K - 1 ? n - used : std::min<int64_t>(n - used, (int64_t)(c / FB_TV + 0.5) * FB_TV);
            used += csize[k];
        }
    }
    int64_t cs = 0, cs2 = 0;               // largest chunk of lane 0 (even chunks) / lane 1 (odd chunks)
    for (int k = 0; k < K; k++) {
        if (k & 1) cs2 = std::max(cs2, csize[k]);
        else cs = std::max(cs, csize[k]);
    }
    const bool pinned = ctx->force_staging ? false : n == 0 || (is_pinned(host_u) && is_pinned(host_v) && is_pinned(host_V_reim) && (!w_stride || is_pinned(host_w)) &&
                                   (!host_chan || is_pinned(host_chan)));
    // Device staging holds the WHOLE call: u [n] | v [n] | V [2 n] | w [n or 1] | chan [n / 2 + 1], so the copy stream never waits
    // for a slot to be consumed -- copies run back to back from the first byte, whatever the kernels are doing.  Pageable
    // inputs pass through a ring of two pinned slots (one per lane, sized for the lane's largest chunk).
    const int64_t D_v = n, D_V = 2 * n, D_w = 4 * n, D_c = 5 * n, d_need = 5 * n + n / 2 + 8;
    {
        FbLane &l0 = ctx->lane[0];
        if (d_need > l0.in_cap) {
            FB_CUDA(cudaDeviceSynchronize());
            if (l0.d_in) FB_CUDA(cudaFree(l0.d_in));
            l0.d_in = nullptr;
            FB_CUDA(cudaMalloc(&l0.d_in, sizeof(double) * d_need));
            l0.in_cap = d_need;
        }
    }
    for (int l = 0; l < (K > 1 ? 2 : 1); l++) {
        FbLane &ln = ctx->lane[l];
        const int64_t L = l ? cs2 : cs, slot = 5 * L + L / 2 + 8;
        rc = fb_reserve_lane(ctx, ln, L, nchan);
        if (rc) return rc;
        rc = reserve_partials(ctx, ln, nchan);
        if (rc) return rc;
        if (!pinned && slot > ln.pin_cap) {
            FB_CUDA(cudaStreamSynchronize(ctx->stream_copy));
            if (ln.h_pin) FB_CUDA(cudaFreeHost(ln.h_pin));
            ln.h_pin = nullptr;
            FB_CUDA(cudaHostAlloc(&ln.h_pin, sizeof(double) * slot, cudaHostAllocDefault));
            ln.pin_cap = slot;
        }
    }
    FbMapJob job;
    job.geom = *geom; job.vis_model = vis_model; job.nchan = nchan; job.check_qbounds = check_qbounds; job.q_last = q_last;
    double *dM = ctx->d_out, *dj = dM + (size_t)nchan * N * N, *dH0 = dj + (size_t)nchan * N;

    for (int attempt = 0; attempt < 2; attempt++) {
        rc = map_begin(ctx);
        if (rc) return rc;
        if (K > 1) FB_CUDA(cudaStreamWaitEvent(ctx->lane[1].stream, ctx->ev[0], 0));      // status reset precedes every chunk
        FB_CUDA(cudaStreamWaitEvent(ctx->stream_copy, ctx->ev[0], 0));
        cudaEvent_t ev_c0 = ctx->mev[4 * (size_t)FB_MAX_CHUNKS], ev_c1 = ctx->mev[4 * (size_t)FB_MAX_CHUNKS + 1];
        FB_CUDA(cudaEventRecord(ev_c0, ctx->stream_copy));
        FbLane *prev = nullptr;
        int64_t off = 0;
        for (int k = 0; k < K; off += csize[k], k++) {
            FbLane &ln = ctx->lane[k & 1];
            const int64_t cnt = csize[k], L = (k & 1) ? cs2 : cs;
            const int64_t off_v = L, off_V = 2 * L, off_w = 4 * L, off_c = 5 * L;      // layout of the lane's pinned slot
            double *dbase = ctx->lane[0].d_in;
            double *d_u = dbase + off, *d_v = dbase + D_v + off, *d_V = dbase + D_V + 2 * off, *d_w = dbase + D_w + (w_stride ? off : 0);
            int32_t *d_c = (int32_t *)(dbase + D_c) + off;
            cudaStream_t sc = ctx->stream_copy;
            if (pinned) {
                FB_CUDA(cudaMemcpyAsync(d_u, host_u + off, sizeof(double) * cnt, cudaMemcpyHostToDevice, sc));
                FB_CUDA(cudaMemcpyAsync(d_v, host_v + off, sizeof(double) * cnt, cudaMemcpyHostToDevice, sc));
                FB_CUDA(cudaMemcpyAsync(d_V, host_V_reim + 2 * off, sizeof(double) * 2 * cnt, cudaMemcpyHostToDevice, sc));
                if (w_stride) FB_CUDA(cudaMemcpyAsync(d_w, host_w + off, sizeof(double) * cnt, cudaMemcpyHostToDevice, sc));
                else if (k == 0) FB_CUDA(cudaMemcpyAsync(d_w, host_w, sizeof(double), cudaMemcpyHostToDevice, sc));
                if (host_chan) FB_CUDA(cudaMemcpyAsync(d_c, host_chan + off, sizeof(int32_t) * cnt, cudaMemcpyHostToDevice, sc));
            } else {
                if (k >= 2) FB_CUDA(cudaEventSynchronize(ln.ev_copied));                   // the pinned slot has left for the device
                double *h = ln.h_pin;
                CopyPiece pc[5];
                int np = 0;
                pc[np++] = {h, host_u + off, sizeof(double) * (size_t)cnt};
                pc[np++] = {h + off_v, host_v + off, sizeof(double) * (size_t)cnt};
                pc[np++] = {h + off_V, host_V_reim + 2 * off, sizeof(double) * 2 * (size_t)cnt};
                if (w_stride) pc[np++] = {h + off_w, host_w + off, sizeof(double) * (size_t)cnt};
                else h[off_w] = host_w[0];
                if (host_chan) pc[np++] = {h + off_c, host_chan + off, sizeof(int32_t) * (size_t)cnt};
                parallel_gather(pc, np, ctx->stage_threads);
                FB_CUDA(cudaMemcpyAsync(d_u, h, sizeof(double) * cnt, cudaMemcpyHostToDevice, sc));
                FB_CUDA(cudaMemcpyAsync(d_v, h + off_v, sizeof(double) * cnt, cudaMemcpyHostToDevice, sc));
                FB_CUDA(cudaMemcpyAsync(d_V, h + off_V, sizeof(double) * 2 * cnt, cudaMemcpyHostToDevice, sc));
                if (w_stride || k == 0) FB_CUDA(cudaMemcpyAsync(d_w, h + off_w, sizeof(double) * (w_stride ? cnt : 1), cudaMemcpyHostToDevice, sc));
                if (host_chan) FB_CUDA(cudaMemcpyAsync(d_c, h + off_c, sizeof(int32_t) * cnt, cudaMemcpyHostToDevice, sc));
            }
            FB_CUDA(cudaEventRecord(ln.ev_copied, sc));
            FB_CUDA(cudaStreamWaitEvent(ln.stream, ln.ev_copied, 0));
            rc = enqueue_chunk(ctx, ln, k, cnt, d_u, d_v, d_V, d_w, w_stride, host_chan ? d_c : nullptr, job, prev);
            if (rc) { cudaDeviceSynchronize(); return rc; }
            prev = &ln;
        }
        FB_CUDA(cudaEventRecord(ev_c1, ctx->stream_copy));
        if (K > 1) {
            FB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->lane[1].ev_done, 0));
            FB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->lane[0].ev_done, 0));
        }
        rc = enqueue_finish(ctx, K, nchan, model_scale, dM, dj, dH0);
        if (rc) { cudaDeviceSynchronize(); return rc; }
        FB_CUDA(cudaMemcpyAsync(host_M, dM, sizeof(double) * nchan * N * N, cudaMemcpyDeviceToHost, ctx->stream));
        FB_CUDA(cudaMemcpyAsync(host_j, dj, sizeof(double) * nchan * N, cudaMemcpyDeviceToHost, ctx->stream));
        FB_CUDA(cudaStreamSynchronize(ctx->stream));
        FB_CUDA(cudaStreamSynchronize(ctx->stream_copy));
        float copy_ms = 0;
        cudaEventElapsedTime(&copy_ms, ev_c0, ev_c1);
        collect_timing(ctx, K, copy_ms);
        ctx->map_chunks = K;
        if (K > 1 && copy_ms > 0.f) {           // arrival and kernel rates of this call steer the next call's chunking
            ctx->rate_copy_ns = 1e6 * (double)copy_ms / (double)n;
            ctx->rate_gram_ns = 1e6 * (ctx->timing[0] + ctx->timing[1]) / (double)n;
            ctx->rate_N = ctx->N;
        }
        rc = map_status(ctx, host_qminmax);
        if (rc == 0) *host_H0 = ctx->h_result[0];
        if (rc != FB_E_RETRY) return rc;
    }
    FB_FAIL(-17, "fb_map_visibilities: J0 table could not be grown to cover the data");
}

int fb_timer_start(fb_ctx *ctx)
{
    if (!ctx) return -1;
    FB_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->tev[0]) { FB_CUDA(cudaEventCreate(&ctx->tev[0])); FB_CUDA(cudaEventCreate(&ctx->tev[1])); }
    FB_CUDA(cudaEventRecord(ctx->tev[0], ctx->stream));
    return 0;
}

int fb_timer_stop(fb_ctx *ctx, double *elapsed_ms)
{
    if (!ctx || !elapsed_ms || !ctx->tev[0]) return -1;
    FB_CUDA(cudaEventRecord(ctx->tev[1], ctx->stream));
    FB_CUDA(cudaEventSynchronize(ctx->tev[1]));
    float ms = 0;
    FB_CUDA(cudaEventElapsedTime(&ms, ctx->tev[0], ctx->tev[1]));
    *elapsed_ms = ms;
    return 0;
}

int fb_last_map_timing(fb_ctx *ctx, double *out4)
{
    if (!ctx || !out4) return -1;
    for (int i = 0; i < 4; i++) out4[i] = ctx->timing[i];
    return 0;
}

int fb_debug_prepped(fb_ctx *ctx, int64_t n, double *host_q, double *host_kz, double *host_Vre, uint32_t *host_perm)
{
    if (!ctx) return -1;
    if (n > ctx->last_n) FB_FAIL(-20, "fb_debug_prepped: n exceeds the last mapped size");
    FB_CUDA(cudaSetDevice(ctx->device));
    std::vector<double> a(n), sw(n), swV(n);
    FB_CUDA(cudaMemcpy(a.data(), ctx->lane[0].d_a, sizeof(double) * n, cudaMemcpyDeviceToHost));
    FB_CUDA(cudaMemcpy(sw.data(), ctx->lane[0].d_sw, sizeof(double) * n, cudaMemcpyDeviceToHost));
    FB_CUDA(cudaMemcpy(swV.data(), ctx->lane[0].d_swV, sizeof(double) * n, cudaMemcpyDeviceToHost));
    FB_CUDA(cudaMemcpy(host_kz, ctx->lane[0].d_kz, sizeof(double) * n, cudaMemcpyDeviceToHost));
    if (host_perm) FB_CUDA(cudaMemcpy(host_perm, ctx->lane[0].d_perm, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
    // the pre-pass stores a = q * (1/Qmax) and sqrt(w) * Re V; hand back a (callers compare a against
    // np.hypot(u', v') * (1/Qmax)) and Re V recovered by the division (exact only to rounding)
    for (int64_t i = 0; i < n; i++) { host_q[i] = a[i]; host_Vre[i] = sw[i] != 0.0 ? swV[i] / sw[i] : 0.0; }
    return 0;
}

}  // extern "C"
