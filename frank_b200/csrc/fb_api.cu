// C ABI of libfrankb200 (see include/frankb200.h): context, DHT setup, map_visibilities entry points.
#include "fb_common.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>

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
    for (auto &e : ctx->ev)
        if (cudaEventCreate(&e) != cudaSuccess) { delete ctx; return -8; }
    *out = ctx;
    return 0;
}

int fb_ctx_destroy(fb_ctx *ctx)
{
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (void *p : {(void *)ctx->d_jk, (void *)ctx->d_ck, (void *)ctx->d_Y, (void *)ctx->d_tab, (void *)ctx->d_types,
                    (void *)ctx->d_tile_panel, (void *)ctx->d_panel_t0, (void *)ctx->d_panel_nt, (void *)ctx->d_pair_code,
                    (void *)ctx->d_a, (void *)ctx->d_sw, (void *)ctx->d_swV, (void *)ctx->d_kz, (void *)ctx->d_amid, (void *)ctx->d_red,
                    (void *)ctx->d_partial, (void *)ctx->d_H2, (void *)ctx->d_in, (void *)ctx->d_out,
                    (void *)ctx->d_rec, (void *)ctx->d_items, (void *)ctx->d_perm, (void *)ctx->d_hist, (void *)ctx->d_binstart, (void *)ctx->d_work, (void *)ctx->d_work2,
                    (void *)ctx->sv_D, (void *)ctx->sv_p, (void *)ctx->sv_mu, (void *)ctx->sv_tr2, (void *)ctx->sv_alpha,
                    (void *)ctx->sv_p0, (void *)ctx->sv_Tinv, (void *)ctx->sv_M, (void *)ctx->sv_j, (void *)ctx->sv_Z,
                    (void *)ctx->sv_flags, (void *)ctx->sv_hist, (void *)ctx->sv_rhs, (void *)ctx->sv_notconv, (void *)ctx->sv_rdiag, (void *)ctx->ln_S, (void *)ctx->ln_vec})
        if (p) cudaFree(p);
    for (auto &e : ctx->ev)
        if (e) cudaEventDestroy(e);
    for (auto &e : ctx->cev)
        if (e) cudaEventDestroy(e);
    if (ctx->stream3) cudaStreamDestroy(ctx->stream3);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return 0;
}

const char *fb_last_error(fb_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int fb_dht_setup(fb_ctx *ctx, int N, double Qmax, const double *host_j_nk, const double *host_coef,
                 const double *host_Ycoef, double x_max)
{
    if (!ctx) return -1;
    if (N < 1 || !host_j_nk || !host_coef || !(Qmax > 0)) FB_FAIL(-10, "fb_dht_setup: bad arguments");
    FB_CUDA(cudaSetDevice(ctx->device));
    FB_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->N = N;
    ctx->NT = (N + 1 + 7) / 8;
    ctx->NC = ctx->NT * 8;
    ctx->Qmax = Qmax;
    ctx->invQmax = 1.0 / Qmax;                    // hankel.py:189  k = 1. / self._Qmax
    ctx->h_jk.assign(ctx->NC, 0.0);
    ctx->h_ck.assign(ctx->NC, 0.0);
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

    // block decomposition of the NT x NT tile grid (fb_gram.cu)
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

// One part of a mapping call: pre-pass + sort + Gram kernel over n visibilities already on the device.
// Accumulates H0 and the range of q on the host (fixed order over the parts -> deterministic).
static int map_part(fb_ctx *ctx, int part, int nparts, int64_t n, const double *u, const double *v, const double *V,
                    const double *w, int w_stride, const fb_geometry *geom, int vis_model, int check_qbounds, double q_last,
                    double *host_H0, double *host_qminmax)
{
    double qmm[2], h0 = 0.0;
    int rc = fb_launch_prep(ctx, n, u, v, V, w, w_stride, geom, nullptr, qmm, &h0);
    if (rc) return rc;
    FB_CUDA(cudaEventRecord(part == 0 ? ctx->ev[1] : ctx->pev[0], ctx->stream));
    if (part == 0) { *host_H0 = h0; host_qminmax[0] = qmm[0]; host_qminmax[1] = qmm[1]; }
    else {
        *host_H0 += h0;
        if (n > 0) { host_qminmax[0] = fmin(host_qminmax[0], qmm[0]); host_qminmax[1] = fmax(host_qminmax[1], qmm[1]); }
    }
    if (n > 0) {
        // statistical_models.py:526: raise when the last collocation point is inside the data
        if (check_qbounds && q_last < qmm[1]) FB_FAIL(FB_E_QRANGE, "last collocation point is at a shorter baseline than the longest deprojected baseline");
        // make sure the J0 table reaches the largest argument a_max * j_{N-1}
        const double xneed = qmm[1] * ctx->invQmax * ctx->h_jk[ctx->N - 1];
        if (fb_j0_rows_for(xneed) > ctx->tab_rows) {
            rc = fb_build_j0_table(ctx, xneed * 1.05);
            if (rc) return rc;
        }
    }
    return fb_launch_gram_part(ctx, part, nparts, n, vis_model);
}

static int map_check_args(fb_ctx *ctx, int64_t n, const fb_geometry *geom, const void *M, const void *j, const void *H0,
                          const double *host_qminmax, const int32_t *chan, int nchan, int vis_model, const double *host_H2)
{
    if (ctx->N == 0) FB_FAIL(-11, "fb_map_visibilities: fb_dht_setup has not been called");
    if (n < 0 || !geom || !M || !j || !H0 || !host_qminmax) FB_FAIL(-12, "fb_map_visibilities: bad arguments");
    if (nchan != 1 || chan != nullptr) FB_FAIL(-13, "fb_map_visibilities: multi-channel input must be split by the caller");
    if (vis_model < 0 || vis_model > 2) FB_FAIL(-14, "fb_map_visibilities: vis_model must be 0, 1 or 2");
    if (vis_model == FB_MODEL_DEBRIS) {
        if (!host_H2) FB_FAIL(-15, "fb_map_visibilities: debris model needs H2");
        std::vector<double> h2(ctx->NC, 0.0);
        for (int k = 0; k < ctx->N; k++) h2[k] = host_H2[k];
        FB_CUDA(cudaMemcpyAsync(ctx->d_H2, h2.data(), sizeof(double) * ctx->NC, cudaMemcpyHostToDevice, ctx->stream));
        FB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    if (!ctx->pev[0])
        for (auto &e : ctx->pev) FB_CUDA(cudaEventCreate(&e));
    return 0;
}

int fb_map_visibilities_dev(fb_ctx *ctx, int64_t n, const double *dev_u, const double *dev_v, const double *dev_V_reim,
                            const double *dev_w, int w_stride, const int32_t *dev_chan, int nchan, const fb_geometry *geom,
                            int vis_model, double model_scale, const double *host_H2, int check_qbounds, double q_last,
                            double *dev_M, double *dev_j, double *dev_H0, double *host_qminmax)
{
    if (!ctx) return -1;
    FB_CUDA(cudaSetDevice(ctx->device));
    int rc = map_check_args(ctx, n, geom, dev_M, dev_j, dev_H0, host_qminmax, dev_chan, nchan, vis_model, host_H2);
    if (rc) return rc;
    FB_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    double h0 = 0.0;
    rc = map_part(ctx, 0, 1, n, dev_u, dev_v, dev_V_reim, dev_w, w_stride, geom, vis_model, check_qbounds, q_last, &h0, host_qminmax);
    if (rc) return rc;
    rc = fb_launch_gram_finalize(ctx, 1, model_scale, dev_M, dev_j);
    if (rc) return rc;
    FB_CUDA(cudaMemcpyAsync(dev_H0, &h0, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    FB_CUDA(cudaEventRecord(ctx->ev[3], ctx->stream));
    FB_CUDA(cudaStreamSynchronize(ctx->stream));
    float t01 = 0, t12 = 0, t23 = 0;
    cudaEventElapsedTime(&t01, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&t12, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&t23, ctx->ev[2], ctx->ev[3]);
    ctx->timing[0] = t01; ctx->timing[1] = t12; ctx->timing[2] = t23; ctx->timing[3] = 0;
    return 0;
}

// Host entry point.  From FB_SPLIT_MIN visibilities on, the call runs as two parts so that the host-to-device
// copy of the second part (on a second stream) overlaps the kernels of the first; the partial blocks of both parts
// are summed in a fixed order, so the result is deterministic (it differs from the one-pass result of the device
// entry point in the last bits, like any other change of the summation order).
// The first part is as small as the overlap allows, because its own copy is the exposed one: the copy of the second
// part (c = 0.75 ns per visibility at ~53 GB/s from pinned memory) has to fit under the first part's kernels
// (p = 0.16 + 3.85 (N/300)^2 ns per visibility, measured: pre-pass + sort + Gram), i.e. f >= c / (c + p); 25 % margin,
// at most one half.  N = 300: f = 0.2 (copy exposed: 1.4 ms of a 1e7-visibility call instead of 3.6 ms).
constexpr int64_t FB_SPLIT_MIN = 4000000;

int fb_map_visibilities_host(fb_ctx *ctx, int64_t n, const double *host_u, const double *host_v, const double *host_V_reim,
                             const double *host_w, int w_stride, const int32_t *host_chan, int nchan, const fb_geometry *geom,
                             int vis_model, double model_scale, const double *host_H2, int check_qbounds, double q_last,
                             double *host_M, double *host_j, double *host_H0, double *host_qminmax)
{
    if (!ctx) return -1;
    FB_CUDA(cudaSetDevice(ctx->device));
    int rc = map_check_args(ctx, n, geom, host_M, host_j, host_H0, host_qminmax, host_chan, nchan, vis_model, host_H2);
    if (rc) return rc;
    const int64_t nw = w_stride ? n : 1;
    const int64_t need = 4 * n + nw + 8;
    if (need > ctx->in_cap) {
        if (ctx->d_in) FB_CUDA(cudaFree(ctx->d_in));
        ctx->d_in = nullptr;
        FB_CUDA(cudaMalloc(&ctx->d_in, sizeof(double) * need));
        ctx->in_cap = need;
    }
    const size_t N = ctx->N, nout = N * N + N + 1;
    if (nout > ctx->out_cap) {
        if (ctx->d_out) FB_CUDA(cudaFree(ctx->d_out));
        ctx->d_out = nullptr;
        FB_CUDA(cudaMalloc(&ctx->d_out, sizeof(double) * nout));
        ctx->out_cap = nout;
    }
    if (!ctx->stream2) FB_CUDA(cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
    double *du = ctx->d_in, *dv = du + n, *dV = dv + n, *dw = dV + 2 * n;
    const int nparts = n >= FB_SPLIT_MIN ? 2 : 1;
    int64_t n0 = n;
    if (nparts == 2) {
        const double c = 0.75, p = 0.16 + 3.85 * ((double)ctx->N / 300.0) * ((double)ctx->N / 300.0);
        const double f = std::min(0.5, 1.25 * c / (c + p));
        n0 = std::max<int64_t>(FB_TV, (int64_t)(f * (double)n) / FB_TV * FB_TV);
    }
    const int64_t n1 = n - n0;
    rc = fb_reserve_prep(ctx, (std::max(n0, n1) + FB_TV - 1) / FB_TV * FB_TV);
    if (rc) return rc;
    auto copy_part = [&](int64_t off, int64_t cnt, cudaStream_t st) -> int {
        FB_CUDA(cudaMemcpyAsync(du + off, host_u + off, sizeof(double) * cnt, cudaMemcpyHostToDevice, st));
        FB_CUDA(cudaMemcpyAsync(dv + off, host_v + off, sizeof(double) * cnt, cudaMemcpyHostToDevice, st));
        FB_CUDA(cudaMemcpyAsync(dV + 2 * off, host_V_reim + 2 * off, sizeof(double) * 2 * cnt, cudaMemcpyHostToDevice, st));
        if (w_stride) FB_CUDA(cudaMemcpyAsync(dw + off, host_w + off, sizeof(double) * cnt, cudaMemcpyHostToDevice, st));
        return 0;
    };
    FB_CUDA(cudaEventRecord(ctx->ev[4], ctx->stream));
    if (!w_stride) FB_CUDA(cudaMemcpyAsync(dw, host_w, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    rc = copy_part(0, n0, ctx->stream);
    if (rc) return rc;
    FB_CUDA(cudaEventRecord(ctx->ev[5], ctx->stream));
    if (nparts == 2) {
        FB_CUDA(cudaStreamWaitEvent(ctx->stream2, ctx->ev[4], 0));           // not before this call's start (buffer reuse)
        rc = copy_part(n0, n1, ctx->stream2);
        if (rc) return rc;
        FB_CUDA(cudaEventRecord(ctx->pev[3], ctx->stream2));
    }
    FB_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    double *dM = ctx->d_out, *dj = dM + N * N, *dH0 = dj + N;
    double h0 = 0.0;
    rc = map_part(ctx, 0, nparts, n0, du, dv, dV, dw, w_stride, geom, vis_model, check_qbounds, q_last, &h0, host_qminmax);
    if (rc) { if (nparts == 2) cudaStreamSynchronize(ctx->stream2); return rc; }
    if (nparts == 2) {
        FB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->pev[3], 0));
        rc = map_part(ctx, 1, 2, n1, du + n0, dv + n0, dV + 2 * n0, w_stride ? dw + n0 : dw, w_stride, geom, vis_model,
                      check_qbounds, q_last, &h0, host_qminmax);
        if (rc) return rc;
    }
    rc = fb_launch_gram_finalize(ctx, nparts, model_scale, dM, dj);
    if (rc) return rc;
    FB_CUDA(cudaEventRecord(ctx->ev[6], ctx->stream));
    FB_CUDA(cudaMemcpyAsync(host_M, dM, sizeof(double) * N * N, cudaMemcpyDeviceToHost, ctx->stream));
    FB_CUDA(cudaMemcpyAsync(host_j, dj, sizeof(double) * N, cudaMemcpyDeviceToHost, ctx->stream));
    FB_CUDA(cudaEventRecord(ctx->ev[7], ctx->stream));
    FB_CUDA(cudaStreamSynchronize(ctx->stream));
    *host_H0 = h0;
    (void)dH0;
    float h2d = 0, d2h = 0, t01 = 0, t12 = 0, t23 = 0, p12 = 0;
    cudaEventElapsedTime(&h2d, ctx->ev[4], ctx->ev[5]);
    cudaEventElapsedTime(&d2h, ctx->ev[6], ctx->ev[7]);
    cudaEventElapsedTime(&t01, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&t12, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&t23, ctx->ev[2], ctx->ev[6]);
    if (nparts == 2) cudaEventElapsedTime(&p12, ctx->pev[1], ctx->pev[2]);
    ctx->timing[0] = t01; ctx->timing[1] = t12 + p12; ctx->timing[2] = t23 - p12; ctx->timing[3] = h2d + d2h;
    return 0;
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
    FB_CUDA(cudaMemcpy(a.data(), ctx->d_a, sizeof(double) * n, cudaMemcpyDeviceToHost));
    FB_CUDA(cudaMemcpy(sw.data(), ctx->d_sw, sizeof(double) * n, cudaMemcpyDeviceToHost));
    FB_CUDA(cudaMemcpy(swV.data(), ctx->d_swV, sizeof(double) * n, cudaMemcpyDeviceToHost));
    FB_CUDA(cudaMemcpy(host_kz, ctx->d_kz, sizeof(double) * n, cudaMemcpyDeviceToHost));
    if (host_perm) FB_CUDA(cudaMemcpy(host_perm, ctx->d_perm, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
    // the pre-pass stores a = q * (1/Qmax) and sqrt(w) * Re V; hand back a (callers compare a against
    // np.hypot(u', v') * (1/Qmax)) and Re V recovered by the division (exact only to rounding)
    for (int64_t i = 0; i < n; i++) { host_q[i] = a[i]; host_Vre[i] = sw[i] != 0.0 ? swV[i] / sw[i] : 0.0; }
    return 0;
}

}  // extern "C"
